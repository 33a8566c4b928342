"""Batch data parallelism: one process per GPU, NCCL all-reduce of the gradient buckets in
the order they become final, overlapped with the rest of the backward pass.

The reference has no parallelism at all; samples are independent on
this path, so the batch is sharded by sample with every parameter replicated and the only
exchange step is the gradient sum.  Each rank scales its gradients by
B_local/B_global inside the kernels (grad_scale), so the all-reduce is a plain SUM.
"""
import torch
import torch.distributed as dist


class GradReducer:
    """comm_dtype=torch.bfloat16 rounds every bucket to bf16 for the exchange (mmqg_pack_bf16 / mmqg_unpack_bf16 around
    the all-reduce, on the communication stream): half the bytes on NVLink and half the time NCCL's CTAs hold SMs that
    the cooperative persistent kernels are waiting for.  The sum is then formed in bf16 (relative error per element
    ~2^-8 * sqrt(world)); fp32 is exact up to summation order.  CUDA buckets only."""

    def __init__(self, engine, world_size, group=None, comm_dtype=torch.float32, schedule="overlap"):
        """schedule: "overlap" = one all-reduce per gradient group, started where the group becomes final, under the rest of
        the backward; "late" = ONE all-reduce of the whole flat gradient buffer after the last group is final (nothing
        competes with the cooperative persistent kernels for SMs; the wire time is exposed instead); "multimem" = per group like
        "overlap", but with our own reduce kernel over NVSwitch multicast memory instead of NCCL (mmqg_allreduce_multimem: the
        engine's flat gradient buffer moves into a symmetric allocation; fp32 sums; needs multicast support)."""
        assert schedule != "multimem" or comm_dtype == torch.float32, "the multimem kernel sums in fp32"
        assert schedule in ("overlap", "late", "multimem")
        self.schedule = schedule
        self.mm = None
        if schedule == "multimem":
            self.mm = _MultimemBuckets(engine, world_size, group)      # moves engine.flat_grads into symmetric memory
            schedule = "overlap"
        self.flat = getattr(engine, "flat_grads", None)
        assert schedule == "overlap" or self.flat is not None, "the late schedule reduces engine.flat_grads"
        self.buckets = engine.grad_buckets
        self.world = world_size
        self.group = group
        self.works = []
        self.events, self.side = None, None
        self.comm_dtype = comm_dtype
        self.stage = None
        if comm_dtype == torch.bfloat16:
            assert self.buckets[0].is_cuda, "bf16 gradient exchange needs CUDA buckets"
            from . import _cabi
            self._lib = _cabi.lib()
            self._check = _cabi.check
            self.stage = [torch.empty(b.numel(), dtype=torch.bfloat16, device=b.device) for b in self.buckets]
            self.stage_flat = torch.empty(self.flat.numel(), dtype=torch.bfloat16, device=self.flat.device) if schedule == "late" else None
        if self.buckets[0].is_cuda:
            # ready events the library records where each gradient group becomes final, and the
            # stream the all-reduces are issued from (so they are ordered after the event only,
            # not after the rest of the backward on the compute stream)
            self.events = [torch.cuda.Event() for _ in range(len(self.buckets))]        # one per gradient group
            for e in self.events:
                e.record()                     # torch creates the cudaEvent_t lazily, at first record
            self.side = torch.cuda.Stream()

    def on_phase(self, i):
        """Called right after backward PHASE i (0 loss head, 1 decoder, 2 video, 3 text layers + embedding) has
        been enqueued on the compute stream.  ProcessGroupNCCL orders the collective after that work on its
        own stream, so it runs under the kernels of the next backward phase (NVLink5/NVSwitch, one flat ring/NVLS)."""
        groups = [i] if i < 3 else range(3, len(self.buckets))
        for g in groups:
            self.works.append(dist.all_reduce(self.buckets[g], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def after_backward(self):
        """After engine.backward_events(): start the all-reduce of every group, each behind its ready event on
        the side stream (loss head, decoder, video, text layers top to bottom, shared embedding)."""
        if self.mm is not None:      # our own reduce kernel over NVSwitch multicast memory, one launch per group behind its event
            for i in range(len(self.buckets)):
                self.side.wait_event(self.events[i])
                with torch.cuda.stream(self.side):
                    self.mm.all_reduce(i)
            return
        if self.side is None:        # CPU buckets (gloo tests): no streams or events, reduce in place
            for t in ([self.flat] if self.schedule == "late" else self.buckets):
                self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        if self.schedule == "late":
            for e in self.events:
                self.side.wait_event(e)
            with torch.cuda.stream(self.side):
                if self.stage is None:
                    self.works.append(dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                else:
                    b, s = self.flat, self.stage_flat
                    sp = torch.cuda.current_stream().cuda_stream
                    self._check(self._lib.mmqg_pack_bf16(b.data_ptr(), s.data_ptr(), b.numel(), sp))
                    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.group, async_op=True).wait()   # stream-level wait
                    self._check(self._lib.mmqg_unpack_bf16(s.data_ptr(), b.data_ptr(), b.numel(), sp))
            return
        for i in range(len(self.buckets)):
            self.side.wait_event(self.events[i])
            with torch.cuda.stream(self.side):
                if self.stage is None:
                    self.works.append(dist.all_reduce(self.buckets[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                    continue
                b, s = self.buckets[i], self.stage[i]
                sp = torch.cuda.current_stream().cuda_stream
                self._check(self._lib.mmqg_pack_bf16(b.data_ptr(), s.data_ptr(), b.numel(), sp))
                dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.group, async_op=True).wait()   # stream-level wait
                self._check(self._lib.mmqg_unpack_bf16(s.data_ptr(), b.data_ptr(), b.numel(), sp))

    def finish(self):
        """Make the compute stream wait for the outstanding all-reduces."""
        for w in self.works:
            w.wait()
        self.works.clear()
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)


class _MultimemBuckets:
    """The engine's flat gradient buffer in symmetric memory (torch.distributed._symmetric_memory: every rank's copy mapped
    on every rank + one NVSwitch multicast address for all of them) and the launcher of mmqg_allreduce_multimem per bucket."""

    def __init__(self, engine, world_size, group=None):
        import os
        import torch.distributed._symmetric_memory as symm_mem
        from . import _cabi
        pg = group if group is not None else dist.group.WORLD
        name = pg.group_name
        try:
            symm_mem.enable_symm_mem_for_group(name)
        except Exception:
            pass
        flat = symm_mem.empty(engine.flat_grads.numel(), dtype=torch.float32, device=engine.flat_grads.device)
        self.hdl = symm_mem.rendezvous(flat, name)
        if not self.hdl.multicast_ptr:
            raise _cabi.MmqgError("multimem all-reduce: no multicast mapping for the gradient buffer (needs NVSwitch + driver support)")
        assert self.hdl.signal_pad_size >= 64 * world_size * 4
        engine.use_flat_grads(flat)
        self.engine, self.rank, self.world = engine, self.hdl.rank, self.hdl.world_size
        # CTAs per launch: few, so that the kernel fits beside the two cooperative 64-CTA recurrent kernels (B200, N=8, cfg-2:
        # 8 -> 5.63 ms per step, 16 -> 5.67, 32 -> 5.72; NCCL 5.83)
        self.ctas = int(os.environ.get("MMQG_MM_CTAS", "8"))
        self._lib, self._check = _cabi.lib(), _cabi.check
        dist.barrier(group=pg)

    def all_reduce(self, i):
        lo, hi = self.engine.bucket_range[i]
        self._check(self._lib.mmqg_allreduce_multimem(self.hdl.multicast_ptr + 4 * lo, hi - lo, self.hdl.signal_pad_ptrs_dev,
                                                      self.rank, self.world, self.ctas, torch.cuda.current_stream().cuda_stream))


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Contiguous per-rank slice of a global batch (samples are independent)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        assert n % world == 0, f"global batch {n} not divisible by {world} ranks"
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per].contiguous()
    return out

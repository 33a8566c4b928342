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
        competes with the cooperative persistent kernels for SMs; the wire time is exposed instead)."""
        assert schedule in ("overlap", "late")
        self.schedule = schedule
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


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Contiguous per-rank slice of a global batch (samples are independent)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        assert n % world == 0, f"global batch {n} not divisible by {world} ranks"
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per].contiguous()
    return out

"""Sequence-level host API over libmmqg.so: one call per train step / greedy decode.

PyTorch is plumbing here (device memory, streams); all arithmetic runs in libmmqg.so.
`TrainEngine.step()` is the batched equivalent of one iteration of the reference's hot
loop (train.py:149-177: zero_grad, encoder, teacher-forced decoder, loss.backward()),
`TrainEngine.greedy()` of train.py:81-110 / evaluate.py:45-104.
"""
import ctypes as C

import torch

from . import _cabi
from .dims import Dims, param_shapes


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class TrainEngine:
    """Holds the parameters, gradients and workspace of one replica on one GPU."""

    def __init__(self, dims: Dims, params: dict, device="cuda", mode="fp32", dropout_p=0.0):
        self.lib = _cabi.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.MmqgError("TrainEngine needs a CUDA device; there is no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _cabi.check(self.lib.mmqg_device_ok(idx))
        self.d = dims
        self.mode = {"fp32": _cabi.MODE_FP32, "bf16": _cabi.MODE_BF16, "fp32_tc": _cabi.MODE_FP32_TC}[mode]
        self.dropout_p = float(dropout_p)
        shapes = param_shapes(dims)
        for name, shape in shapes.items():
            if tuple(params[name].shape) != tuple(shape):
                raise _cabi.MmqgError(f"{name}: shape {tuple(params[name].shape)} != {shape}")
        # Parameters and gradients live in ONE flat fp32 buffer each with identical layout, cut into
        # 4 + L buckets, one per gradient readiness group (grad_group()): the data-parallel
        # all-reduce is one NCCL call per bucket and the fused Adam is one launch over everything;
        # self.params / self.grads are views with the reference's names and shapes.
        self.offsets, total, bucket_range = {}, 0, []
        self.n_groups = 4 + dims.L
        for g in range(self.n_groups):
            lo = total
            for n in shapes:
                if grad_group(n, dims.L) == g:
                    self.offsets[n] = total
                    total += (int(torch.Size(shapes[n]).numel()) + 63) // 64 * 64
            bucket_range.append((lo, total))
        self.flat_params = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_grads = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.params, self.grads = {}, {}
        for n, shape in shapes.items():
            o, k = self.offsets[n], int(torch.Size(shape).numel())
            self.params[n] = self.flat_params[o:o + k].view(shape)
            self.params[n].copy_(params[n].detach().to(self.device, torch.float32))
            self.grads[n] = self.flat_grads[o:o + k].view(shape)
        self.grad_buckets = [self.flat_grads[lo:hi] for lo, hi in bucket_range]
        self.bucket_range = bucket_range
        self._shapes = shapes
        self._adam = None
        self._cd = _cabi.c_dims(dims)
        self._cp = _cabi.c_tensors(self.params, dims.L)
        self._cg = _cabi.c_tensors(self.grads, dims.L)
        nbytes = self.lib.mmqg_train_workspace_bytes(C.byref(self._cd), self.mode)
        if nbytes == 0:
            raise _cabi.MmqgError(self.lib.mmqg_last_error().decode())
        # zeroed once: its first 8 bytes are the dropout call counter (see mmqg.h), the rest is scratch
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._greedy_ws = None
        self.seed = 0          # base dropout seed; the library adds the number of forward calls made so far

    def use_flat_grads(self, flat):
        """Move the gradients into a caller-provided flat fp32 buffer of the same length (e.g. a symmetric-memory
        allocation the multimem all-reduce works on, mmqg.dp.GradReducer(schedule="multimem")): self.grads,
        self.grad_buckets and the C-ABI struct are rebuilt as views of it."""
        assert flat.numel() == self.flat_grads.numel() and flat.dtype == torch.float32 and flat.is_cuda and flat.is_contiguous()
        flat.zero_()
        self.flat_grads = flat
        for n, shape in self._shapes.items():
            o, k = self.offsets[n], int(torch.Size(shape).numel())
            self.grads[n] = flat[o:o + k].view(shape)
        self.grad_buckets = [flat[lo:hi] for lo, hi in self.bucket_range]
        self._cg = _cabi.c_tensors(self.grads, self.d.L)

    # -- batches -------------------------------------------------------------------------
    LENGTH_KEYS = ("ctx_len", "tgt_len", "n_frames")

    def to_device(self, batch: dict, non_blocking=False) -> dict:
        out = {}
        for k, v in batch.items():
            want = torch.int64 if k in ("context", "target") else (torch.int32 if k in self.LENGTH_KEYS else torch.float32)
            out[k] = v.to(self.device, want, non_blocking=non_blocking).contiguous()
        return out

    def _cbatch(self, b):
        d = self.d
        assert tuple(b["context"].shape) == (d.B, d.T_t) and tuple(b["frames"].shape) == (d.B, d.T_v, d.F_v)
        assert tuple(b["audio"].shape) == (d.B, d.T_v, d.H_a)
        for k, v in b.items():
            assert v.is_cuda and v.is_contiguous(), k
        tgt = b.get("target")
        lens = []
        for k in self.LENGTH_KEYS:                       # optional per-sample lengths, int32 (B)
            v = b.get(k)
            assert v is None or (v.dtype == torch.int32 and tuple(v.shape) == (d.B,)), k
            lens.append(0 if v is None else v.data_ptr())
        return _cabi.MmqgBatch(b["context"].data_ptr(), 0 if tgt is None else tgt.data_ptr(),
                               b["frames"].data_ptr(), b["audio"].data_ptr(), *lens)

    # -- train step ----------------------------------------------------------------------
    def forward(self, batch, want_grads=True, grad_scale=1.0):
        cb = self._cbatch(batch)
        assert tuple(batch["target"].shape) == (self.d.B, self.d.T_q)
        _cabi.check(self.lib.mmqg_train_forward(
            C.byref(self._cd), C.byref(self._cp), C.byref(cb), self.ws.data_ptr(), self.ws.numel(),
            self.loss.data_ptr(), int(want_grads), C.byref(self._cg), float(grad_scale), self.dropout_p,
            self.seed, self.mode, _stream_ptr()))
        return self.loss

    def backward(self, batch, phase):
        cb = self._cbatch(batch)
        _cabi.check(self.lib.mmqg_train_backward(
            C.byref(self._cd), C.byref(self._cp), C.byref(cb), self.ws.data_ptr(), self.ws.numel(),
            C.byref(self._cg), int(phase), self.dropout_p, self.seed, self.mode, _stream_ptr()))

    def backward_events(self, batch, events):
        """Whole backward with its internal overlap; events[i] (torch.cuda.Event, already created)
        is recorded where gradient group i+1 becomes final (mmqg_train_backward_events)."""
        cb = self._cbatch(batch)
        assert len(events) == 4 + self.d.L, "one ready event per gradient group"
        arr = (C.c_void_p * len(events))(*[e.cuda_event for e in events])
        _cabi.check(self.lib.mmqg_train_backward_events(
            C.byref(self._cd), C.byref(self._cp), C.byref(cb), self.ws.data_ptr(), self.ws.numel(),
            C.byref(self._cg), arr, self.dropout_p, self.seed, self.mode, _stream_ptr()))

    def step_dp(self, batch, reducer, grad_scale):
        """Data-parallel step: forward, then the overlapped backward with one all-reduce per gradient group
        (loss head included: its gradients may only be final inside the backward call, see mmqg.h) started from
        the group's ready event."""
        loss = self.forward(batch, True, grad_scale)
        self.backward_events(batch, reducer.events)
        reducer.after_backward()
        return loss

    def step(self, batch, grad_scale=1.0, on_phase=None):
        """forward + full backward.  Gradients land in self.grads (overwritten).  on_phase(i)
        is called after backward phase i (0 loss head, 1 decoder, 2 video, 3 text layers + embedding)
        has been enqueued -- the hook the phase-wise data-parallel all-reduce uses."""
        loss = self.forward(batch, True, grad_scale)
        if on_phase is None:
            self.backward(batch, 0)          # whole backward, hoisted products overlapped internally
        else:
            for ph in (1, 2, 3):
                self.backward(batch, ph)
                if ph == 1:
                    on_phase(0)              # the loss-head gradients are final at the latest after phase 1 (mmqg.h)
                on_phase(ph)
        return loss

    # -- optimiser -------------------------------------------------
    def adam_init(self):
        """Allocate the optimiser state (zero moments, step count 0).  Call before capturing a CUDA
        graph that contains adam_step(): allocations and their zero-fill must not be captured."""
        if self._adam is None:
            z = torch.zeros_like(self.flat_params)
            self._adam = {"m": z, "v": z.clone(), "state": torch.zeros(4, dtype=torch.float32, device=self.device)}
        return self._adam

    def adam_step(self, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        """The reference's three torch.optim.Adam(lr=1e-4) steps (train.py:179-181, 265-267) as
        one fused launch over the flat parameter buffer; the shared embedding, registered with two
        of those optimisers, is stepped twice.  Uses self.grads as they are (after the
        all-reduce in data-parallel runs).  CUDA-graph capturable (step count on the device)."""
        a = self.adam_init()
        lo = self.offsets["emb.weight"]
        hi = lo + self.params["emb.weight"].numel()
        _cabi.check(self.lib.mmqg_adam_step(
            self.flat_params.data_ptr(), self.flat_grads.data_ptr(), a["m"].data_ptr(), a["v"].data_ptr(),
            self.flat_params.numel(), lo, hi, float(lr), float(betas[0]), float(betas[1]), float(eps),
            a["state"].data_ptr(), _stream_ptr()))

    def dropout_calls(self) -> int:
        """Forward calls with dropout made on this workspace (the device counter the masks depend on)."""
        return int(self.ws[:8].view(torch.int64).item())

    def reset_dropout_calls(self, value=0):
        self.ws[:8].view(torch.int64).fill_(int(value))

    def dropout_masks(self, seed=None):
        """The multiplicative inter-layer dropout masks (0 or 1/(1-p)) of the NEXT step (or of an
        explicit effective `seed`): {"text": (L-1,T_t,B,H), "dec": (L-1,T_q,B,H)} -- what tests feed
        the oracle.  Effective seed of a step = self.seed + number of forward calls up to and
        including it, counted on the device so that a replayed CUDA graph draws fresh masks too."""
        d = self.d
        seed = self.seed + self.dropout_calls() + 1 if seed is None else seed
        out = {}
        for key, sid0, T in (("text", 10, d.T_t), ("dec", 20, d.T_q)):
            m = torch.empty(max(d.L - 1, 0), T, d.B, d.H, dtype=torch.float32, device=self.device)
            for l in range(d.L - 1):
                _cabi.check(self.lib.mmqg_dropout_mask(m[l].data_ptr(), m[l].numel(), seed, sid0 + l, self.dropout_p,
                                                      _stream_ptr()))
            out[key] = m
        return out

    # -- greedy decode ---------------------------------------------------------------------
    def greedy(self, batch, max_len, strategy="greedy", seed=0):
        """Decode max_len tokens per sample (no early exit; callers cut at <end>).  strategy: "greedy"
        (evaluate.py:71-80), "topk" (the reference's topk(1): the same tokens) or "sampling"
        (evaluate.py:84-90: draw from softmax(logits), deterministic in `seed`)."""
        if strategy not in ("greedy", "topk", "sampling"):
            raise ValueError(f"unknown decode strategy {strategy!r}")
        cb = self._cbatch(batch)
        key = int(max_len)
        if self._greedy_ws is None or self._greedy_ws[0] != key:
            n = self.lib.mmqg_greedy_workspace_bytes(C.byref(self._cd), key, self.mode)
            if n == 0:
                raise _cabi.MmqgError(self.lib.mmqg_last_error().decode())
            ws = self.ws if n <= self.ws.numel() else torch.empty(n, dtype=torch.uint8, device=self.device)
            self._greedy_ws = (key, ws)
        ws = self._greedy_ws[1]
        toks = torch.empty(self.d.B, key, dtype=torch.int64, device=self.device)
        if strategy == "sampling":
            _cabi.check(self.lib.mmqg_sample_decode(
                C.byref(self._cd), C.byref(self._cp), C.byref(cb), ws.data_ptr(), ws.numel(), toks.data_ptr(), key,
                int(seed), self.mode, _stream_ptr()))
        else:
            _cabi.check(self.lib.mmqg_greedy_decode(
                C.byref(self._cd), C.byref(self._cp), C.byref(cb), ws.data_ptr(), ws.numel(), toks.data_ptr(), key,
                self.mode, _stream_ptr()))
        return toks


def grad_group(name: str, L: int = 3) -> int:
    """Readiness group of a gradient tensor, in the order the groups become final during the backward:
    0 loss head (final after the forward's fused loss head), 1 attention Linears + decoder LSTM, 2 video LSTM,
    3 + k text LSTM layer L-1-k (the top layer's BPTT finishes first), 3 + L the shared embedding (final last:
    decoder- and encoder-side scatter-adds both land in it)."""
    if name.startswith("dec.out_layer."):
        return 0
    if name.startswith("dec."):
        return 1
    if name.startswith("video."):
        return 2
    if name.startswith("text.lstm."):
        return 3 + (L - 1 - int(name.rsplit("_l", 1)[1]))
    return 3 + L


def launch_count():
    return int(_cabi.lib().mmqg_launch_count())


class HostFeed:
    """Feeds train steps from HOST batches (pinned memory): two device batch buffers, the
    host->device copy of step i+1 runs on a copy stream under the compute of step i.

        feed = HostFeed(engine, example_batch, make_step)    # make_step(device_batch) -> callable() -> loss
        feed.prefetch(host_batch_0)
        for i in range(n):
            loss = feed.step()                 # step i on the buffer prefetched for it
            feed.prefetch(host_batch_i_plus_1) # enqueued behind the step that last read that buffer
            value = float(loss)                # device->host read of step i's result

    make_step is called once per buffer (a CUDA graph is bound to the buffers it was captured on)."""

    def __init__(self, engine, example_batch, make_step):
        self.bufs = [engine.to_device(example_batch) for _ in range(2)]
        self.steps = [make_step(b) for b in self.bufs]
        self.copy_stream = torch.cuda.Stream()
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.n_prefetched = 0
        self.n_stepped = 0

    def prefetch(self, host_batch):
        k = self.n_prefetched % 2
        assert self.n_prefetched - self.n_stepped < 2, "both buffers hold batches that have not been consumed"
        cs = self.copy_stream
        cs.wait_event(self.done[k])            # the step that last read buffer k (no-op before its first use)
        with torch.cuda.stream(cs):
            for name, dst in self.bufs[k].items():
                dst.copy_(host_batch[name], non_blocking=True)
            self.copied[k].record(cs)
        self.n_prefetched += 1

    def step(self):
        assert self.n_stepped < self.n_prefetched, "prefetch() the batch first"
        k = self.n_stepped % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(self.copied[k])
        loss = self.steps[k]()
        self.done[k].record(cur)
        self.n_stepped += 1
        return loss


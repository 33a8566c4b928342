"""torch.autograd glue for the step-wise drop-in modules (model/encoder.py, model/decoder.py).

The reference's train.py drives its modules one token / one decoder step at a time and relies
on autograd to stitch the steps together (train.py:164-177).  To run that loop unchanged, each
building block is an autograd.Function whose forward and backward are libmmqg.so kernels
(fp32 building blocks of include/mmqg.h through mmqg.ops); PyTorch only owns the tensors and
the graph.  The sequence-level engine (mmqg.engine) is the fast path; this file is the
compatibility path and shares its kernels.
"""
import torch
from torch.autograd import Function

from . import ops


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class Embedding(Function):
    """out(n,:) = weight(idx(n),:)   (reference encoder.py:96, decoder.py:75)."""

    @staticmethod
    def forward(ctx, weight, idx):
        idx = idx.reshape(-1)
        ctx.save_for_backward(idx)
        ctx.shape = weight.shape
        return ops.embedding_gather(_c(weight), idx)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dw = torch.zeros(ctx.shape, device=dout.device, dtype=torch.float32)     # dense, like the reference
        ops.embedding_scatter_add(dw, idx, _c(dout))
        return dw, None


class Linear(Function):
    """y = x W^T + b   (decoder.py:78,84,92,106)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = _c(x), _c(w)
        ctx.save_for_backward(x, w)
        return ops.gemm(x, w, False, True, bias=b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _c(dy)
        dx = ops.gemm(dy, w, False, False)
        dw = ops.gemm(dy, x, True, False)
        db = ops.colsum(dy)
        return dx, dw, db


class Attention(Function):
    """Three location-attention heads of one decoder step (decoder.py:78-97) given the
    pre-softmax scores (B, TM+2AM) packed [text|audio|video]; returns the softmax weights
    and the packed context [c_txt|c_aud|c_vid].  No length mask."""

    @staticmethod
    def forward(ctx, scores, M_txt, M_aud, M_vid, T_t, T_v):
        attn = scores.clone()
        M_txt, M_aud, M_vid = _c(M_txt), _c(M_aud), _c(M_vid)
        c = ops.attn_fwd(attn, M_txt, M_aud, M_vid, T_t, T_v)
        ctx.save_for_backward(attn, M_txt, M_aud, M_vid)
        ctx.T = (T_t, T_v)
        return attn, c

    @staticmethod
    def backward(ctx, dattn, dc):
        attn, M_txt, M_aud, M_vid = ctx.saved_tensors
        if ctx.needs_input_grad[2]:
            raise NotImplementedError("gradient w.r.t. the audio features (VGGish) is outside the hot path")
        T_t, T_v = ctx.T
        ds = attn.clone()
        dM_txt = torch.zeros_like(M_txt)
        dM_vid = torch.zeros_like(M_vid)
        ops.attn_bwd(ds, _c(dc), M_txt, M_aud, M_vid, dM_txt, dM_vid, T_t, T_v)
        if dattn is not None and bool(dattn.any()):
            # the reference only returns the weights for inspection; if a caller differentiates
            # through them, add the softmax Jacobian of that path with plain tensor ops
            TM, AM = M_txt.shape[1], M_aud.shape[1]
            for lo, hi in ((0, TM), (TM, TM + AM), (TM + AM, TM + 2 * AM)):
                a, g = attn[:, lo:hi], dattn[:, lo:hi]
                ds[:, lo:hi] += a * (g - (a * g).sum(1, keepdim=True))
        return ds, dM_txt, None, dM_vid, None, None


class LSTMStack(Function):
    """torch.nn.LSTM(input, hidden, num_layers) over a (T,B,I) sequence, gate order i,f,g,o, both
    biases (reference encoder.py:91,54, decoder.py:69).  Weight gradients are hoisted into one
    product per layer over the whole sequence; per step only h W_hh^T and the cell update run.
    `masks`: optional (L-1,T,B,H) inter-layer dropout masks (already scaled by 1/(1-p))."""

    @staticmethod
    def forward(ctx, x, h0, c0, masks, *weights):
        T, B, _ = x.shape
        L = len(weights) // 4
        H = h0.shape[2]
        inp = _c(x)
        saved = []
        hn, cn = [], []
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh = [_c(w) for w in weights[4 * l:4 * l + 4]]
            acts = ops.gemm(inp.reshape(T * B, -1), w_ih, False, True, bias=b_ih + b_hh).view(T, B, 4 * H)
            hs = torch.empty(T + 1, B, H, device=x.device, dtype=torch.float32)
            cs = torch.empty(T + 1, B, H, device=x.device, dtype=torch.float32)
            hs[0].copy_(h0[l])
            cs[0].copy_(c0[l])
            for t in range(T):
                ops.gemm(hs[t], w_hh, False, True, out=acts[t], Cin=acts[t])
                h, c = ops.lstm_pointwise_fwd(acts[t], cs[t])
                hs[t + 1].copy_(h)
                cs[t + 1].copy_(c)
            saved += [inp, acts, hs, cs]
            hn.append(hs[T])
            cn.append(cs[T])
            inp = hs[1:]
            if masks is not None and l < L - 1:
                inp = inp * masks[l]
        ctx.save_for_backward(*saved, *weights, *([masks] if masks is not None else []))
        ctx.dims = (T, B, H, L, masks is not None)
        return inp.clone(), torch.stack(hn), torch.stack(cn)

    @staticmethod
    def backward(ctx, dy, dhn, dcn):
        T, B, H, L, has_mask = ctx.dims
        sv = ctx.saved_tensors
        saved, weights = sv[:4 * L], sv[4 * L:8 * L]
        masks = sv[8 * L] if has_mask else None
        dev = dy.device
        dext = _c(dy).clone() if dy is not None else torch.zeros(T, B, H, device=dev)
        dh0, dc0 = [None] * L, [None] * L
        grads = [None] * (4 * L)
        dx = None
        for l in range(L - 1, -1, -1):
            inp, acts, hs, cs = saved[4 * l:4 * l + 4]
            w_ih, w_hh = _c(weights[4 * l]), _c(weights[4 * l + 1])
            dg_all = acts.clone()
            dc = dcn[l].clone() if dcn is not None else torch.zeros(B, H, device=dev)
            dh_rec = _c(dhn[l]).clone() if dhn is not None else torch.zeros(B, H, device=dev)
            for t in range(T - 1, -1, -1):
                ops.lstm_pointwise_bwd(dg_all[t], cs[t], cs[t + 1], None, dh_rec, dext[t], dc)
                dh_rec = ops.gemm(dg_all[t], w_hh, False, False)
            dh0[l], dc0[l] = dh_rec, dc
            flat = dg_all.view(T * B, 4 * H)
            grads[4 * l] = ops.gemm(flat, inp.reshape(T * B, -1), True, False)
            grads[4 * l + 1] = ops.gemm(flat, hs[:T].reshape(T * B, H), True, False)
            db = ops.colsum(flat)
            grads[4 * l + 2], grads[4 * l + 3] = db, db.clone()
            dx = ops.gemm(flat, w_ih, False, False).view(T, B, -1)
            if l > 0:
                dext = dx * masks[l - 1] if masks is not None else dx
        return (dx, torch.stack(dh0), torch.stack(dc0), None, *grads)


class LSTMLayerSeq(Function):
    """ONE torch.nn.LSTM layer over a whole (T,B,I) sequence on the tensor-core path
    (mmqg_lstm_seq_fwd / _bwd: hoisted input projection + persistent recurrent kernel, persistent BPTT
    kernel + hoisted weight gradients; bf16 operands, fp32 accumulation).  The call TextEncoder.forward
    makes for a whole context (encoder.py:95-100), VideoConvLstmEncoder for the frame features
    (encoder.py:69) and the non-attention Decoder for a whole question (decoder.py:25-34)."""

    @staticmethod
    def forward(ctx, x, h0, c0, w_ih, w_hh, b_ih, b_hh):
        T, B, I = x.shape
        H = h0.shape[1]
        y, hn, cn, ws = ops.lstm_seq_fwd(_c(x), _c(w_ih), _c(w_hh), _c(b_ih), _c(b_hh), _c(h0), _c(c0))
        ctx.save_for_backward(w_hh, ws)
        ctx.dims = (T, B, I, H)
        return y, hn, cn

    @staticmethod
    def backward(ctx, dy, dhn, dcn):
        w_hh, ws = ctx.saved_tensors
        T, B, I, H = ctx.dims
        g = ops.lstm_seq_bwd(None if dy is None else _c(dy), None if dhn is None else _c(dhn),
                             None if dcn is None else _c(dcn), _c(w_hh), ws, T, B, I, H, want_dx=ctx.needs_input_grad[0])
        return g["dx"], g["dh0"], g["dc0"], g["dw_ih"], g["dw_hh"], g["db_ih"], g["db_hh"]


_mask_calls = 0


class ConvStack(Function):
    """f2: the conv stack of VideoConvLstmEncoder (reference model/encoder.py:62-64): four Conv2d -> ReLU -> BatchNorm2d
    blocks with a MaxPool2d(k, k) behind the second and the fourth, on the CUDA kernels of csrc/convstack.cu.
    Train mode uses batch statistics and updates the running buffers like torch.nn.BatchNorm2d; eval mode uses the
    running statistics.  Inputs: x (N,C,H,W); per layer conv weight, conv bias, bn weight, bn bias; `state`: per layer
    (running_mean, running_var, eps, momentum) (buffers, updated in place) and (K, stride, training)."""

    POOL_AFTER = (1, 3)

    @staticmethod
    def forward(ctx, x, state, *params):
        bn_state, K, stride, training = state
        x = _c(x)
        n_layers = len(params) // 4
        saved, meta = [], []
        inp, sc, sh = x, None, None
        for l in range(n_layers):
            w, b, gamma, beta = [_c(p) for p in params[4 * l:4 * l + 4]]
            rm, rv, eps, mom = bn_state[l]
            y, stats = ops.conv_relu_fwd(inp, w, b, sc, sh, stride, want_stats=training)
            if training:
                count = y.shape[0] * y.shape[2] * y.shape[3]
                scale, shift, mean, invstd = ops.bn_finalize(stats, count, gamma, beta, eps, mom, rm, rv)
            else:
                invstd = torch.rsqrt(rv + eps)
                mean = rm
                scale = gamma * invstd
                shift = beta - rm * scale
            saved += [inp, sc, sh, y, mean, invstd]
            if l in ConvStack.POOL_AFTER:
                pooled, idx = ops.bn_maxpool_fwd(y, scale, shift, K)
                saved.append(idx)
                meta.append((True, y.shape[2], y.shape[3]))
                inp, sc, sh = pooled, None, None
            else:
                saved.append(None)
                meta.append((False, y.shape[2], y.shape[3]))
                inp, sc, sh = y, scale, shift
        if sc is not None:      # stack that does not end in a pool: materialise the last BatchNorm
            inp = inp * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        ctx.meta = (meta, K, stride, training, n_layers)
        ctx.save_for_backward(*[t for t in saved if t is not None], *params)
        ctx.mask = [t is not None for t in saved]
        return inp

    @staticmethod
    def backward(ctx, dout):
        meta, K, stride, training, n_layers = ctx.meta
        it = iter(ctx.saved_tensors)
        saved = [next(it) if m else None for m in ctx.mask]
        params = list(it)
        grads = [None] * (4 * n_layers)
        d = _c(dout)
        for l in range(n_layers - 1, -1, -1):
            inp, sc, sh, y, mean, invstd, idx = saved[7 * l:7 * l + 7]
            w, gamma = _c(params[4 * l]), _c(params[4 * l + 2])
            pooled, H, W = meta[l]
            C_ = y.shape[1]
            if pooled and training:      # pooled gradient -> d z in one pass, no dense gradient w.r.t. the BatchNorm output
                dz, sums = ops.bn_relu_pool_bwd(y, mean, invstd, gamma, _c(d), idx, K, train=True)
            else:
                if pooled:
                    dbn = ops.maxpool_bwd(d, idx, H, W, K)
                else:      # d is ours (conv_bwd_x of the layer above) unless it is autograd's own dout
                    dbn = d.clone() if l == n_layers - 1 else d
                if not training:      # eval-mode BatchNorm: d beta = sum d, d gamma = sum d * xhat over the fixed statistics
                    xhat = (y - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
                    grads[4 * l + 3], grads[4 * l + 2] = dbn.sum((0, 2, 3)), (dbn * xhat).sum((0, 2, 3))
                dz, sums = ops.bn_relu_bwd(y, mean, invstd, gamma, dbn, train=training)      # in place: dz aliases dbn
            if training:
                grads[4 * l + 3], grads[4 * l + 2] = sums[:C_].clone(), sums[C_:].clone()
            grads[4 * l], grads[4 * l + 1] = ops.conv_bwd_w(inp, dz, K, stride, sc, sh)
            if l > 0:
                d = ops.conv_bwd_x(dz, w, inp.shape[2], inp.shape[3], stride)
        return (None, None, *grads)


def conv_stack(x, module):
    """Run module.conv1..4 / bn1..4 / maxpool1..2 of a VideoConvLstmEncoder on (N,C,H,W) frames -> (N, 10, h, w)."""
    convs = [module.conv1, module.conv2, module.conv3, module.conv4]
    bns = [module.bn1, module.bn2, module.bn3, module.bn4]
    K = convs[0].kernel_size[0]
    stride = convs[0].stride[0]
    for c in convs:
        assert c.kernel_size == (K, K) and c.stride == (stride, stride) and c.padding == (0, 0) and c.dilation == (1, 1) and c.groups == 1
    assert module.maxpool1.kernel_size == K and module.maxpool1.stride == K and module.maxpool2.kernel_size == K
    training = module.training
    bn_state = []
    params = []
    for c, bn in zip(convs, bns):
        mom = bn.momentum if bn.momentum is not None else 0.1
        bn_state.append((bn.running_mean, bn.running_var, bn.eps, mom))
        params += [c.weight, c.bias, bn.weight, bn.bias]
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return ConvStack.apply(x, (bn_state, K, stride, training), *params)


def _dropout_masks(L, T, B, H, p, device):
    """(L-1,T,B,H) inter-layer dropout masks (0 or 1/(1-p)) drawn by the library's counter-based generator
    (mmqg_dropout_mask), a fresh stream per call; torch.manual_seed() reseeds it through torch.initial_seed()."""
    global _mask_calls
    _mask_calls += 1
    seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * _mask_calls) & 0xFFFFFFFFFFFFFFFF
    m = torch.empty(L - 1, T, B, H, device=device, dtype=torch.float32)
    for l in range(L - 1):
        ops.dropout_mask(m[l], seed, 30 + l, p)
    return m


def drop_in_mode():
    """Arithmetic of the step-wise drop-in modules: "fp32" (default: fp32 storage and accumulation, the mode that
    carries the 1e-3 parity bar) or "bf16" (MMQG_DROPIN_MODE=bf16: whole-sequence LSTM calls run on the
    tensor-core sequence kernels).  An explicit choice, not a fallback: shapes the sequence kernels do not
    support raise in bf16 mode only when the call cannot be served by the per-step blocks either."""
    import os
    return os.environ.get("MMQG_DROPIN_MODE", "fp32")


def lstm_stack(x, hidden, lstm_module, training):
    """Run `lstm_module`'s parameters (a torch.nn.LSTM used purely as the parameter container
    that keeps the reference's state_dict keys) through the CUDA kernels."""
    h0, c0 = hidden
    L = lstm_module.num_layers
    weights = []
    for l in range(L):
        weights += [getattr(lstm_module, f"weight_ih_l{l}"), getattr(lstm_module, f"weight_hh_l{l}"),
                    getattr(lstm_module, f"bias_ih_l{l}"), getattr(lstm_module, f"bias_hh_l{l}")]
    masks = None
    p = float(lstm_module.dropout)
    T, B, _ = x.shape
    H = lstm_module.hidden_size
    if training and p > 0 and L > 1:
        masks = _dropout_masks(L, T, B, H, p, x.device)
    if drop_in_mode() == "bf16" and T > 1 and ops.lstm_seq_ok(B, H):
        inp, hn, cn = x, [], []
        for l in range(L):
            y, h, c = LSTMLayerSeq.apply(inp, h0[l], c0[l], *weights[4 * l:4 * l + 4])
            hn.append(h)
            cn.append(c)
            inp = y * masks[l] if (masks is not None and l < L - 1) else y
        return inp, (torch.stack(hn), torch.stack(cn))
    y, hn, cn = LSTMStack.apply(x, _c(h0), _c(c0), masks, *weights)
    return y, (hn, cn)


_cat_cache = {}


def cat_cached(key, tensors, dim=0):
    """torch.cat of parameters, redone only when one of them changed (optimizer step, load_state_dict) or when
    autograd needs the graph through it; saves the three attention Linears' re-concatenation on every decoder
    step in inference loops (decoder.py:78,84,92 as one product)."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        return torch.cat(tensors, dim)
    ver = tuple((t.data_ptr(), t._version) for t in tensors)
    hit = _cat_cache.get(key)
    if hit is None or hit[0] != ver:
        hit = (ver, torch.cat([t.detach() for t in tensors], dim))
        _cat_cache[key] = hit
    return hit[1]


def require_cuda(*tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor) and not t.is_cuda:
            raise RuntimeError("mmqg modules run on CUDA only (B200, sm_100a): move the module and its inputs to a CUDA "
                               "device; there is no CPU fallback")

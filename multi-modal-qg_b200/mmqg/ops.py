"""Tensor-level wrappers of the libmmqg.so building blocks (include/mmqg.h, second half).

Each function takes CUDA float32 tensors, passes raw pointers / leading dimensions to the
C ABI on torch's current stream, and returns torch tensors.  They raise on CPU tensors:
there is no fallback path.
"""
import ctypes as C

import torch

from . import _cabi
from ._cabi import MmqgGemmArgs, check, lib


def _st():
    return torch.cuda.current_stream().cuda_stream


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _cabi.MmqgError("mmqg ops need CUDA tensors (no CPU fallback)")
        if t.dtype not in (torch.float32, torch.int64, torch.bfloat16):
            raise _cabi.MmqgError(f"unsupported dtype {t.dtype}")
        if t.dim() >= 1 and t.stride(-1) != 1 and t.numel() > 1:
            raise _cabi.MmqgError("innermost dimension must be contiguous")


def _ld(t):
    return t.stride(0) if t.dim() == 2 and t.shape[0] > 1 else t.shape[-1]


def gemm(A, B, transA=False, transB=False, out=None, A2=None, B2=None, Cin=None, beta=1.0, bias=None,
         alpha=1.0, split_k=1):
    """out = alpha*(op(A)op(B) [+ op(A2)op(B2)]) + beta*Cin + bias.  2-D row-major views
    with arbitrary row stride are accepted.  split_k>1 returns the (split_k,M,N) partials."""
    _chk(A, B, A2, B2, Cin, bias, out)
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == K
    if out is None:
        out = torch.empty((split_k, M, N) if split_k > 1 else (M, N), device=A.device, dtype=torch.float32)
    a = MmqgGemmArgs()
    a.A, a.B, a.lda, a.ldb, a.K = A.data_ptr(), B.data_ptr(), _ld(A), _ld(B), K
    if A2 is not None:
        K2 = A2.shape[0] if transA else A2.shape[1]
        a.A2, a.B2, a.lda2, a.ldb2, a.K2 = A2.data_ptr(), B2.data_ptr(), _ld(A2), _ld(B2), K2
    a.C = out.data_ptr()
    a.ldc = _ld(out[0]) if split_k > 1 else _ld(out)
    if Cin is not None:
        a.Cin, a.ldcin = Cin.data_ptr(), _ld(Cin)
    if bias is not None:
        a.bias = bias.data_ptr()
    a.M, a.N, a.alpha, a.beta = M, N, alpha, beta
    a.transA, a.transB, a.split_k = int(transA), int(transB), split_k
    a.c_split_stride = out.stride(0) if split_k > 1 else 0
    check(lib().mmqg_gemm_f32(C.byref(a), _st()))
    return out


def gemm_bf16(A, B, a_mn_major=False, b_mn_major=False, out=None, out_dtype=torch.float32, A2=None, B2=None,
              Cin=None, beta=1.0, bias=None, alpha=1.0, split_k=1):
    """tcgen05 GEMM on bf16 operands.  A: (M,K) or, if a_mn_major, (K,M); B: (N,K) or, if
    b_mn_major, (K,N).  Returns fp32 or bf16 (out_dtype); split_k>1 returns fp32 partials."""
    _chk(A, B, A2, B2, Cin, bias, out)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    M = A.shape[1] if a_mn_major else A.shape[0]
    K = A.shape[0] if a_mn_major else A.shape[1]
    N = B.shape[1] if b_mn_major else B.shape[0]
    assert (B.shape[0] if b_mn_major else B.shape[1]) == K
    if out is None:
        out = torch.empty((split_k, M, N) if split_k > 1 else (M, N), device=A.device, dtype=out_dtype)
    a = _cabi.MmqgGemmBf16Args()
    a.A, a.B, a.lda, a.ldb, a.K = A.data_ptr(), B.data_ptr(), _ld(A), _ld(B), K
    if A2 is not None:
        a.A2, a.B2, a.lda2, a.ldb2 = A2.data_ptr(), B2.data_ptr(), _ld(A2), _ld(B2)
        a.K2 = A2.shape[0] if a_mn_major else A2.shape[1]
    a.a_mn_major, a.b_mn_major = int(a_mn_major), int(b_mn_major)
    a.C = out.data_ptr()
    a.ldc = _ld(out[0]) if split_k > 1 else _ld(out)
    a.c_bf16 = int(out.dtype == torch.bfloat16)
    if Cin is not None:
        a.Cin, a.ldcin = Cin.data_ptr(), _ld(Cin)
    if bias is not None:
        a.bias = bias.data_ptr()
    a.M, a.N, a.alpha, a.beta, a.split_k = M, N, alpha, beta, split_k
    a.c_split_stride = out.stride(0) if split_k > 1 else 0
    check(lib().mmqg_gemm_bf16(C.byref(a), _st()))
    return out


def embedding_gather(emb, idx):
    _chk(emb, idx)
    idx = idx.reshape(-1).contiguous()
    out = torch.empty(idx.numel(), emb.shape[1], device=emb.device, dtype=torch.float32)
    check(lib().mmqg_embedding_gather(emb.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), emb.shape[1],
                                      emb.shape[0], _st()))
    return out


def embedding_scatter_add(demb, idx, dx):
    _chk(demb, idx, dx)
    idx = idx.reshape(-1).contiguous()
    dx = dx.contiguous()
    check(lib().mmqg_embedding_scatter_add(demb.data_ptr(), idx.data_ptr(), dx.data_ptr(), idx.numel(),
                                           demb.shape[1], demb.shape[0], _st()))
    return demb


def lstm_pointwise_fwd(gates, c_prev, h2=None):
    """gates (B,4H) pre-activations, overwritten with the activated gates.  Returns (h, c)."""
    _chk(gates, c_prev, h2)
    Bn, G = gates.shape
    H = G // 4
    c = torch.empty(Bn, H, device=gates.device, dtype=torch.float32)
    h = torch.empty_like(c)
    check(lib().mmqg_lstm_pointwise_fwd(gates.data_ptr(), _ld(gates), _cabi.ptr(c_prev),
                                        0 if c_prev is None else _ld(c_prev), c.data_ptr(), H, h.data_ptr(), H,
                                        _cabi.ptr(h2), 0 if h2 is None else _ld(h2), Bn, H, _st()))
    return h, c


def lstm_pointwise_bwd(acts, c_prev, c_new, dh_parts, dh1, dh2, dc):
    """acts (B,4H) activated gates -> overwritten with d/d pre-activations; dc (B,H) d/dc' ->
    overwritten with d/dc_prev (dc None = zeros, a fresh buffer is returned).
    dh_parts: (n,B,H) partial sums or None; dh1, dh2: (B,H) or None."""
    _chk(acts, c_prev, c_new, dh_parts, dh1, dh2, dc)
    Bn, G = acts.shape
    H = G // 4
    zero = dc is None
    if zero:
        dc = torch.empty(Bn, H, device=acts.device, dtype=torch.float32)
    n0 = 0 if dh_parts is None else dh_parts.shape[0]
    s0 = 0 if dh_parts is None else dh_parts.stride(0)
    check(lib().mmqg_lstm_pointwise_bwd(
        acts.data_ptr(), _ld(acts), _cabi.ptr(c_prev), 0 if c_prev is None else _ld(c_prev), c_new.data_ptr(),
        _ld(c_new), _cabi.ptr(dh_parts), H, n0, s0, _cabi.ptr(dh1), 0 if dh1 is None else _ld(dh1), 1, 0,
        _cabi.ptr(dh2), 0 if dh2 is None else _ld(dh2), dc.data_ptr(), _ld(dc), int(zero), Bn, H, _st()))
    return acts, dc


def attn_fwd(scores, M_txt, M_aud, M_vid, T_t, T_v, ctx=None):
    """scores (B,S[+pad]) -> softmaxes in place; returns ctx (B,H+H_a+H_v)."""
    _chk(scores, M_txt, M_aud, M_vid, ctx)
    Bn, TM, H = M_txt.shape
    AM, H_a = M_aud.shape[1:]
    H_v = M_vid.shape[2]
    if ctx is None:
        ctx = torch.empty(Bn, H + H_a + H_v, device=scores.device, dtype=torch.float32)
    check(lib().mmqg_attn_fwd(scores.data_ptr(), _ld(scores), M_txt.data_ptr(), M_aud.data_ptr(), M_vid.data_ptr(),
                              ctx.data_ptr(), _ld(ctx), Bn, TM, AM, H, H_a, H_v, T_t, T_v, _st()))
    return ctx


def attn_bwd(attn, dctx, M_txt, M_aud, M_vid, dM_txt, dM_vid, T_t, T_v):
    """attn (B,S) softmax outputs -> overwritten with d/d scores; dM_* accumulated (+=)."""
    _chk(attn, dctx, M_txt, M_aud, M_vid, dM_txt, dM_vid)
    Bn, TM, H = M_txt.shape
    AM, H_a = M_aud.shape[1:]
    H_v = M_vid.shape[2]
    check(lib().mmqg_attn_bwd(attn.data_ptr(), _ld(attn), dctx.data_ptr(), _ld(dctx), M_txt.data_ptr(),
                              M_aud.data_ptr(), M_vid.data_ptr(), _cabi.ptr(dM_txt), _cabi.ptr(dM_vid), Bn, TM, AM, H,
                              H_a, H_v, T_t, T_v, _st()))
    return attn


def nll_rows(logits, targets, dlogits_scale=0.0):
    _chk(logits, targets)
    R, V = logits.shape
    targets = targets.reshape(-1).contiguous()
    nll = torch.empty(R, device=logits.device, dtype=torch.float32)
    check(lib().mmqg_nll_rows(logits.data_ptr(), _ld(logits), targets.data_ptr(), 1, nll.data_ptr(), R, V,
                              float(dlogits_scale), _st()))
    return nll


def argmax_rows(logits):
    _chk(logits)
    R, V = logits.shape
    tok = torch.empty(R, device=logits.device, dtype=torch.int64)
    check(lib().mmqg_argmax_rows(logits.data_ptr(), _ld(logits), tok.data_ptr(), 1, R, V, _st()))
    return tok


def sample_rows(logits, seed=0, step=0):
    """tokens(r) ~ softmax(logits(r,:)) -- the reference's 'sampling' strategy (evaluate.py:84-90)."""
    _chk(logits)
    R, V = logits.shape
    tok = torch.empty(R, device=logits.device, dtype=torch.int64)
    check(lib().mmqg_sample_rows(logits.data_ptr(), _ld(logits), tok.data_ptr(), 1, R, V, int(seed), int(step), _st()))
    return tok


def sample_uniform(n, seed=0, step=0, device="cuda"):
    """The uniforms sample_rows(seed, step) uses for rows 0..n-1 (for tests)."""
    out = torch.empty(n, device=device, dtype=torch.float32)
    check(lib().mmqg_sample_uniform(out.data_ptr(), n, int(seed), int(step), _st()))
    return out


def colsum(X, out=None, beta=0.0):
    _chk(X, out)
    M, N = X.shape
    if out is None:
        out = torch.empty(N, device=X.device, dtype=torch.float32)
    check(lib().mmqg_colsum(X.data_ptr(), _ld(X), out.data_ptr(), M, N, float(beta), _st()))
    return out


def dropout_mask(out, seed, sid, p):
    """out <- the multiplicative mask (0 or 1/(1-p)) of the library's counter-based generator (mmqg_dropout_mask)."""
    _chk(out)
    check(lib().mmqg_dropout_mask(out.data_ptr(), out.numel(), int(seed), int(sid), float(p), _st()))
    return out


_sm_count = {}


def lstm_seq_ok(B, H):
    """Shapes the persistent sequence kernels take (see mmqg_lstm_seq_fwd in mmqg.h)."""
    dev = torch.cuda.current_device()
    if dev not in _sm_count:
        _sm_count[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return H % 64 == 0 and 64 <= H <= 512 and (H // 16) * ((B + 127) // 128) <= _sm_count[dev]


def lstm_seq_fwd(x, w_ih, w_hh, b_ih, b_hh, h0=None, c0=None):
    """One LSTM layer over x (T,B,I) on the tensor-core sequence kernels.  Returns y (T,B,H), hn, cn (B,H) and the
    workspace tensor that lstm_seq_bwd needs (saved activations)."""
    _chk(x, w_ih, w_hh, b_ih, b_hh, h0, c0)
    T, B, I = x.shape
    H = w_hh.shape[1]
    n = lib().mmqg_lstm_seq_workspace_bytes(T, B, I, H)
    ws = torch.empty(n, dtype=torch.uint8, device=x.device)
    y = torch.empty(T, B, H, device=x.device, dtype=torch.float32)
    hn = torch.empty(B, H, device=x.device, dtype=torch.float32)
    cn = torch.empty_like(hn)
    check(lib().mmqg_lstm_seq_fwd(x.data_ptr(), w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(),
                                  _cabi.ptr(h0), _cabi.ptr(c0), T, B, I, H, ws.data_ptr(), n, y.data_ptr(), hn.data_ptr(),
                                  cn.data_ptr(), _st()))
    return y, hn, cn, ws


def lstm_seq_bwd(dy, dhn, dcn, w_hh, ws, T, B, I, H, want_dx=True):
    """Backward of lstm_seq_fwd (same workspace).  Returns a dict of fp32 gradients."""
    _chk(dy, dhn, dcn, w_hh)
    dev = w_hh.device
    g = {"dx": torch.empty(T, B, I, device=dev, dtype=torch.float32) if want_dx else None,
         "dw_ih": torch.empty(4 * H, I, device=dev, dtype=torch.float32),
         "dw_hh": torch.empty(4 * H, H, device=dev, dtype=torch.float32),
         "db_ih": torch.zeros(4 * H, device=dev, dtype=torch.float32),
         "db_hh": torch.zeros(4 * H, device=dev, dtype=torch.float32),
         "dh0": torch.empty(B, H, device=dev, dtype=torch.float32),
         "dc0": torch.empty(B, H, device=dev, dtype=torch.float32)}
    check(lib().mmqg_lstm_seq_bwd(_cabi.ptr(dy), _cabi.ptr(dhn), _cabi.ptr(dcn), w_hh.data_ptr(), T, B, I, H, ws.data_ptr(),
                                  ws.numel(), _cabi.ptr(g["dx"]), g["dw_ih"].data_ptr(), g["dw_hh"].data_ptr(),
                                  g["db_ih"].data_ptr(), g["db_hh"].data_ptr(), g["dh0"].data_ptr(), g["dc0"].data_ptr(), _st()))
    return g


# ---- f2: conv stack building blocks (csrc/convstack.cu) --------------------------------------------------------------
def _p(t):
    return t.data_ptr() if t is not None else None


def conv_relu_fwd(x, w, b, in_scale=None, in_shift=None, stride=1, want_stats=True):
    """y = relu(conv2d(x * in_scale + in_shift, w) + b) on (N,Cin,H,W) fp32; returns (y, stats) with stats = per-block partial
    (sum | sum of squares) of y per channel, (nparts, 2*Cout) floats (None when want_stats is False)."""
    _chk(x, w, b, in_scale, in_shift)
    N, Cin, H, W = x.shape
    Cout, _, K, _ = w.shape
    Ho, Wo = (H - K) // stride + 1, (W - K) // stride + 1
    y = torch.empty(N, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
    nparts = lib().mmqg_conv_stats_parts(N, H, W, K, stride)
    stats = torch.empty(nparts, 2 * Cout, device=x.device, dtype=torch.float32) if want_stats else None
    check(lib().mmqg_conv_relu_fwd(x.data_ptr(), _p(in_scale), _p(in_shift), w.data_ptr(), _p(b), y.data_ptr(), _p(stats),
                                   N, Cin, H, W, Cout, K, stride, _st()))
    return y, stats


def bn_finalize(stats, count, gamma, beta, eps, momentum, running_mean, running_var):
    """Train-mode BatchNorm2d statistics -> (scale, shift, mean, invstd); running stats updated in place (None = skip)."""
    stats = stats.view(-1, stats.shape[-1]) if stats.dim() > 1 else stats.view(1, -1)
    C_ = stats.shape[1] // 2
    buf = torch.empty(4, C_, device=stats.device, dtype=torch.float32)      # scale | shift contiguous: they double as the reduction scratch
    out = [buf[i] for i in range(4)]
    check(lib().mmqg_bn_finalize(stats.data_ptr(), stats.shape[0], int(count), _p(gamma), _p(beta), float(eps), float(momentum), _p(running_mean),
                                 _p(running_var), *[o.data_ptr() for o in out], C_, _st()))
    return out


def bn_maxpool_fwd(y, scale, shift, K):
    N, C_, H, W = y.shape
    Hp, Wp = (H - K) // K + 1, (W - K) // K + 1
    out = torch.empty(N, C_, Hp, Wp, device=y.device, dtype=torch.float32)
    idx = torch.empty(N, C_, Hp, Wp, device=y.device, dtype=torch.uint8)
    check(lib().mmqg_bn_maxpool_fwd(y.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), idx.data_ptr(), N, C_, H, W, K, _st()))
    return out, idx


def maxpool_bwd(dpool, idx, H, W, K):
    N, C_ = dpool.shape[:2]
    dbn = torch.empty(N, C_, H, W, device=dpool.device, dtype=torch.float32)
    check(lib().mmqg_maxpool_bwd(dpool.data_ptr(), idx.data_ptr(), dbn.data_ptr(), N, C_, H, W, K, _st()))
    return dbn


def bn_relu_bwd(y, mean, invstd, gamma, dbn, train=True):
    """Gradient w.r.t. the conv output (in place of dbn) and, in train mode, sums = (d beta | d gamma)."""
    N, C_, H, W = y.shape
    sums = torch.empty(2 * C_, device=y.device, dtype=torch.float32) if train else None
    check(lib().mmqg_bn_relu_bwd(y.data_ptr(), _p(mean), invstd.data_ptr(), _p(gamma), dbn.data_ptr(), dbn.data_ptr(), _p(sums),
                                 N, C_, H, W, _st()))
    return dbn, sums


def bn_relu_pool_bwd(y, mean, invstd, gamma, dpool, idx, K, train=True):
    """maxpool_bwd + bn_relu_bwd in one pass (no dense gradient w.r.t. the BatchNorm output): returns (dz, sums)."""
    N, C_, H, W = y.shape
    _chk(y, dpool)
    dz = torch.empty_like(y)
    sums = torch.empty(2 * C_, device=y.device, dtype=torch.float32) if train else None
    check(lib().mmqg_bn_relu_pool_bwd(y.data_ptr(), _p(mean), invstd.data_ptr(), _p(gamma), dpool.data_ptr(), idx.data_ptr(), K,
                                      dz.data_ptr(), _p(sums), N, C_, H, W, _st()))
    return dz, sums


def conv_bwd_w(x, dz, K, stride=1, in_scale=None, in_shift=None):
    N, Cin, H, W = x.shape
    Cout = dz.shape[1]
    dw = torch.empty(Cout, Cin, K, K, device=x.device, dtype=torch.float32)
    db = torch.empty(Cout, device=x.device, dtype=torch.float32)
    check(lib().mmqg_conv_bwd_w(x.data_ptr(), _p(in_scale), _p(in_shift), dz.data_ptr(), dw.data_ptr(), db.data_ptr(), N, Cin, H, W,
                                Cout, K, stride, _st()))
    return dw, db


def conv_bwd_x(dz, w, H, W, stride=1):
    N, Cout = dz.shape[:2]
    _, Cin, K, _ = w.shape
    dxn = torch.empty(N, Cin, H, W, device=dz.device, dtype=torch.float32)
    check(lib().mmqg_conv_bwd_x(dz.data_ptr(), w.data_ptr(), dxn.data_ptr(), N, Cin, H, W, Cout, K, stride, _st()))
    return dxn

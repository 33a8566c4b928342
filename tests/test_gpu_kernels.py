"""Building-block kernels vs plain PyTorch fp32/fp64 on the same inputs (GPU only).
Every call goes through the C ABI (mmqg.ops -> ctypes -> libmmqg.so)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import ops as o
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


GEMM_SHAPES = [
    # M, N, K, transA, transB
    (256, 2048, 512, False, True),     # recurrent step h W_hh^T
    (256, 512, 2048, False, False),    # dG W_hh
    (2048, 512, 2560, True, False),    # dW = dG^T H
    (1536, 1280, 300, False, True),    # hoisted projection, big tiles
    (37, 29, 23, False, True),         # ragged scalar path
    (37, 29, 23, True, False),
    (37, 29, 23, False, False),
    (37, 29, 23, True, True),
    (64, 64, 16, True, True),
    (3, 128, 12, False, True),
    (130, 260, 1030, False, True),
]


@pytest.mark.parametrize("M,N,K,tA,tB", GEMM_SHAPES)
def test_gemm(ops, M, N, K, tA, tB):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g)
    ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double()
    out = ops.gemm(A, B, tA, tB)
    assert rel(out, ref) < 2e-6
    # bias + addend + second operand pair + alpha
    K2 = max(4, K // 2)
    A2 = torch.randn((K2, M) if tA else (M, K2), device="cuda", generator=g)
    B2 = torch.randn((N, K2) if tB else (K2, N), device="cuda", generator=g)
    Cin = torch.randn(M, N, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    ref2 = 0.5 * (ref + (A2.t() if tA else A2).double() @ (B2.t() if tB else B2).double()) + 2.0 * Cin.double() + bias.double()
    out2 = ops.gemm(A, B, tA, tB, A2=A2, B2=B2, Cin=Cin, beta=2.0, bias=bias, alpha=0.5)
    assert rel(out2, ref2) < 2e-6
    # split-K partials sum to the product
    parts = ops.gemm(A, B, tA, tB, split_k=4)
    assert rel(parts.sum(0), ref) < 2e-6


def test_gemm_strided_views(ops):
    """Leading dimensions larger than the logical width, offset base pointers (column slices
    of W_ih_l0, per-timestep slabs of batch-major frames)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    W = torch.randn(2048, 1452, device="cuda", generator=g)
    x = torch.randn(64, 1152, device="cuda", generator=g)
    out = ops.gemm(x, W[:, 300:], False, True)
    assert rel(out, x.double() @ W[:, 300:].double().t()) < 2e-6
    frames = torch.randn(8, 10, 2048, device="cuda", generator=g)
    Wv = torch.randn(512, 2048, device="cuda", generator=g)
    out = ops.gemm(frames[:, 3, :], Wv, False, True)
    assert rel(out, frames[:, 3, :].double() @ Wv.double().t()) < 2e-6
    dst = torch.zeros(2048, 1452, device="cuda")
    dG = torch.randn(96, 2048, device="cuda", generator=g)
    ops.gemm(dG, x[:, :300].contiguous().repeat(2, 1)[:96], True, False, out=dst[:, :300])
    assert rel(dst[:, :300], dG.double().t() @ x[:, :300].repeat(2, 1)[:96].double()) < 2e-6
    assert float(dst[:, 300:].abs().max()) == 0.0


@pytest.mark.parametrize("B,H", [(256, 512), (3, 32), (5, 24)])
def test_lstm_pointwise(ops, B, H):
    g = torch.Generator(device="cuda").manual_seed(B + H)
    gates = torch.randn(B, 4 * H, device="cuda", generator=g) * 2
    c_prev = torch.randn(B, H, device="cuda", generator=g)
    x = gates.double().clone().requires_grad_(True)
    cp = c_prev.double().clone().requires_grad_(True)
    i, f, gg, o = x[:, :H].sigmoid(), x[:, H:2 * H].sigmoid(), x[:, 2 * H:3 * H].tanh(), x[:, 3 * H:].sigmoid()
    c_ref = f * cp + i * gg
    h_ref = o * c_ref.tanh()
    acts = gates.clone()
    h2 = torch.zeros(B, H + 8, device="cuda")
    h, c = ops.lstm_pointwise_fwd(acts, c_prev, h2[:, :H])
    assert rel(h, h_ref) < 1e-6 and rel(c, c_ref) < 1e-6
    assert torch.equal(h2[:, :H], h) and float(h2[:, H:].abs().max()) == 0.0
    assert rel(acts, torch.cat([i, f, gg, o], 1)) < 1e-6
    # c_prev = None means zeros
    acts0 = gates.clone()
    h0, c0 = ops.lstm_pointwise_fwd(acts0, None)
    assert rel(c0, (i * gg)) < 1e-6
    # backward
    dh = torch.randn(B, H, device="cuda", generator=g)
    dcn = torch.randn(B, H, device="cuda", generator=g)
    parts = torch.randn(3, B, H, device="cuda", generator=g)
    dh_tot = dh.double() + parts.double().sum(0)
    (h_ref * dh_tot + c_ref * dcn.double()).sum().backward()
    dc = dcn.clone()
    dg, dc_prev = ops.lstm_pointwise_bwd(acts, c_prev, c, parts, dh, None, dc)
    assert rel(dg, x.grad) < 2e-6 and rel(dc_prev, cp.grad) < 2e-6


def _attn_ref(scores, M_txt, M_aud, M_vid, TM, AM):
    a_t = scores[:, :TM].softmax(1)
    a_a = scores[:, TM:TM + AM].softmax(1)
    a_v = scores[:, TM + AM:TM + 2 * AM].softmax(1)
    ctx = torch.cat([torch.bmm(a_t.unsqueeze(1), M_txt).squeeze(1), torch.bmm(a_a.unsqueeze(1), M_aud).squeeze(1),
                     torch.bmm(a_v.unsqueeze(1), M_vid).squeeze(1)], 1)
    return torch.cat([a_t, a_a, a_v], 1), ctx


@pytest.mark.parametrize("B,TM,AM,H,H_a,H_v,T_t,T_v", [(16, 283, 101, 512, 128, 512, 100, 10),
                                                       (3, 9, 5, 32, 8, 32, 6, 3), (2, 7, 4, 16, 6, 24, 5, 2),
                                                       (4, 400, 101, 512, 128, 512, 400, 64)])
def test_attention(ops, B, TM, AM, H, H_a, H_v, T_t, T_v):
    g = torch.Generator(device="cuda").manual_seed(TM + AM)
    S = TM + 2 * AM
    Sp = (S + 3) // 4 * 4
    scores = torch.randn(B, Sp, device="cuda", generator=g)
    M_txt = torch.randn(B, TM, H, device="cuda", generator=g); M_txt[:, T_t:] = 0
    M_aud = torch.randn(B, AM, H_a, device="cuda", generator=g); M_aud[:, T_v:] = 0
    M_vid = torch.randn(B, AM, H_v, device="cuda", generator=g); M_vid[:, T_v:] = 0
    sd = scores[:, :S].double().clone().requires_grad_(True)
    mt = M_txt.double().clone().requires_grad_(True)
    mv = M_vid.double().clone().requires_grad_(True)
    a_ref, ctx_ref = _attn_ref(sd, mt, M_aud.double(), mv, TM, AM)
    attn = scores.clone()
    ctx = ops.attn_fwd(attn, M_txt, M_aud, M_vid, T_t, T_v)
    assert rel(attn[:, :S], a_ref) < 2e-6 and rel(ctx, ctx_ref) < 2e-6
    # padded slots keep probability mass (reference mask is a no-op)
    if TM > T_t:
        assert float(attn[:, T_t:TM].sum()) > 0
    dctx = torch.randn_like(ctx)
    (ctx_ref * dctx.double()).sum().backward()
    dM_txt = torch.zeros_like(M_txt)
    dM_vid = torch.zeros_like(M_vid)
    ds = ops.attn_bwd(attn, dctx, M_txt, M_aud, M_vid, dM_txt, dM_vid, T_t, T_v)
    assert rel(ds[:, :S], sd.grad) < 5e-6
    assert rel(dM_txt[:, :T_t], mt.grad[:, :T_t]) < 2e-6
    assert rel(dM_vid[:, :T_v], mv.grad[:, :T_v]) < 2e-6


@pytest.mark.parametrize("R,V", [(64, 10000), (7, 37), (5, 50000)])
def test_nll_argmax_colsum(ops, R, V):
    g = torch.Generator(device="cuda").manual_seed(R + V)
    logits = torch.randn(R, V, device="cuda", generator=g) * 3
    tgt = torch.randint(0, V, (R,), device="cuda", generator=g)
    x = logits.double().clone().requires_grad_(True)
    nll_ref = torch.nn.functional.cross_entropy(x, tgt, reduction="none")
    (nll_ref.sum() * 0.25).backward()
    work = logits.clone()
    nll = ops.nll_rows(work, tgt, 0.0)
    assert torch.equal(work, logits) and rel(nll, nll_ref) < 1e-6
    nll = ops.nll_rows(work, tgt, 0.25)
    assert rel(work, x.grad) < 2e-6
    # argmax: lowest index on ties
    logits[:, 5] = 100.0
    logits[:, 3] = 100.0
    tok = ops.argmax_rows(logits)
    assert torch.equal(tok, torch.full((R,), 3, device="cuda"))
    logits2 = torch.randn(R, V, device="cuda", generator=g)
    assert torch.equal(ops.argmax_rows(logits2), logits2.argmax(1))
    assert rel(ops.colsum(logits2), logits2.double().sum(0)) < 1e-6


def test_embedding(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    emb = torch.randn(1000, 300, device="cuda", generator=g)
    idx = torch.randint(0, 1000, (4, 77), device="cuda", generator=g)
    out = ops.embedding_gather(emb, idx)
    assert torch.equal(out, emb[idx.reshape(-1)])
    dx = torch.randn(4 * 77, 300, device="cuda", generator=g)
    demb = torch.zeros_like(emb)
    ops.embedding_scatter_add(demb, idx, dx)
    ref = torch.zeros_like(emb).double().index_add_(0, idx.reshape(-1), dx.double())
    assert rel(demb, ref) < 1e-6


def test_pack_unpack_bf16_round_trip(ops):
    """mmqg_pack_bf16 / mmqg_unpack_bf16 (bf16 gradient exchange): round-to-nearest-even like torch, tail elements included."""
    from mmqg import _cabi
    lib = _cabi.lib()
    for n in (1, 3, 4, 1027, 1 << 20):
        x = torch.randn(n, device="cuda") * 3
        y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
        z = torch.empty(n, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        _cabi.check(lib.mmqg_pack_bf16(x.data_ptr(), y.data_ptr(), n, st))
        _cabi.check(lib.mmqg_unpack_bf16(y.data_ptr(), z.data_ptr(), n, st))
        torch.cuda.synchronize()
        assert torch.equal(y, x.to(torch.bfloat16))
        assert torch.equal(z, x.to(torch.bfloat16).float())

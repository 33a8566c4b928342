"""Fused loss head (K4) and greedy arg-max (K6) through the C ABI against a plain fp64 PyTorch statement of
decoder.py:106 + train.py:174 / train.py:107-108 on the SAME bf16-rounded operands.

Tolerances: nll / lse 1e-4 relative (fp32 accumulation of bf16 products, __expf); dH, dW, db 2e-2 relative per
tensor (the d-logits operand is rounded to bf16 before the two products: 2^-9 per element)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import _cabi
    return _cabi


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def stream():
    return torch.cuda.current_stream().cuda_stream


def make_case(R, V, H, seed, scale_w=1.0):
    g = torch.Generator().manual_seed(seed)
    h = (torch.randn(R, H, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(V, H, generator=g) * (scale_w / H ** 0.5)).to(torch.bfloat16)
    b = torch.randn(V, generator=g)
    t = torch.randint(0, V, (R,), generator=g, dtype=torch.int64)
    rw = (torch.rand(R, generator=g) > 0.2).float()
    return h, w, b, t, rw


def reference(h, w, b, t, rw, scale):
    hd, wd = h.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    logits = hd @ wd.t() + bd
    lse = torch.logsumexp(logits, 1)
    nll = rw.double() * (lse - logits.gather(1, t[:, None]).squeeze(1))
    (scale * nll.sum()).backward()
    return nll.detach(), lse.detach(), hd.grad, wd.grad, bd.grad, logits.detach()


@pytest.mark.parametrize("R,V,H", [(5, 37, 64), (130, 1003, 64), (300, 10000, 512), (256, 2049, 128)])
def test_vocab_nll_fwd_bwd_match_fp64(lib, R, V, H):
    L = lib.lib()
    h, w, b, t, rw = make_case(R, V, H, seed=R + V)
    scale = 1.0 / 7.0
    nll_ref, lse_ref, dh_ref, dw_ref, db_ref, _ = reference(h, w, b, t, rw, scale)
    dev = "cuda"
    hd, wd, bd, td, rwd = h.to(dev), w.to(dev), b.to(dev), t.to(dev), rw.to(dev)
    n = L.mmqg_vocab_workspace_bytes(R, V, H)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    nll = torch.empty(R, device=dev); lse = torch.empty(R, device=dev); rs = torch.empty(R, device=dev)
    lib.check(L.mmqg_vocab_nll_fwd(hd.data_ptr(), wd.data_ptr(), bd.data_ptr(), td.data_ptr(), rwd.data_ptr(), R, V, H, scale,
                                   ws.data_ptr(), n, nll.data_ptr(), lse.data_ptr(), rs.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert rel(lse, lse_ref) < 1e-4
    assert rel(nll, nll_ref) < 1e-4
    assert torch.allclose(rs.cpu(), scale * rw)
    dH = torch.empty(R, H, device=dev); dW = torch.empty(V, H, device=dev); db = torch.empty(V, device=dev)
    lib.check(L.mmqg_vocab_nll_bwd(hd.data_ptr(), wd.data_ptr(), bd.data_ptr(), td.data_ptr(), lse.data_ptr(), rs.data_ptr(), R, V, H,
                                   ws.data_ptr(), n, dH.data_ptr(), dW.data_ptr(), db.data_ptr(), 0, stream()))
    torch.cuda.synchronize()
    errs = (rel(dH, dh_ref), rel(dW, dw_ref), rel(db, db_ref))
    print(f"vocab R={R} V={V} H={H}: nll {rel(nll, nll_ref):.2e} dH {errs[0]:.2e} dW {errs[1]:.2e} db {errs[2]:.2e}")
    assert max(errs) < 2e-2, errs
    # accumulate = 1 adds a second copy into dW / db
    lib.check(L.mmqg_vocab_nll_bwd(hd.data_ptr(), wd.data_ptr(), bd.data_ptr(), td.data_ptr(), lse.data_ptr(), rs.data_ptr(), R, V, H,
                                   ws.data_ptr(), n, dH.data_ptr(), dW.data_ptr(), db.data_ptr(), 1, stream()))
    torch.cuda.synchronize()
    assert rel(dW, 2 * dw_ref) < 2e-2 and rel(db, 2 * db_ref) < 2e-2


def test_vocab_bwd_chunked_rows_equal_one_chunk(lib, monkeypatch):
    """Row chunking of the d-logits operand (MMQG_LH_BYTES) must not change the result."""
    L = lib.lib()
    monkeypatch.setenv("MMQG_LH_BYTES", str(1 << 20))       # 1 MB of bf16 d logits: 128-row chunks, six of them
    R, V, H = 700, 4100, 128
    h, w, b, t, rw = make_case(R, V, H, seed=3)
    _, lse_ref, dh_ref, dw_ref, db_ref, _ = reference(h, w, b, t, rw, 0.5)
    dev = "cuda"
    hd, wd, bd, td = h.to(dev), w.to(dev), b.to(dev), t.to(dev)
    lse = lse_ref.float().to(dev); rs = (0.5 * rw).to(dev)
    n = L.mmqg_vocab_workspace_bytes(R, V, H)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    dH = torch.empty(R, H, device=dev); dW = torch.empty(V, H, device=dev); db = torch.empty(V, device=dev)
    lib.check(L.mmqg_vocab_nll_bwd(hd.data_ptr(), wd.data_ptr(), bd.data_ptr(), td.data_ptr(), lse.data_ptr(), rs.data_ptr(), R, V, H,
                                   ws.data_ptr(), n, dH.data_ptr(), dW.data_ptr(), db.data_ptr(), 0, stream()))
    torch.cuda.synchronize()
    assert max(rel(dH, dh_ref), rel(dW, dw_ref), rel(db, db_ref)) < 2e-2


@pytest.mark.parametrize("R,V,H", [(7, 300, 64), (1024, 10000, 512), (130, 2049, 128)])
def test_decode_step_argmax_lowest_index_on_ties(lib, R, V, H):
    L = lib.lib()
    h, w, b, _, _ = make_case(R, V, H, seed=11 + R, scale_w=8.0)
    # exact ties: duplicate the winning column's weights and bias at a HIGHER index for half of the rows' likely winners
    logits = h.double() @ w.double().t() + b.double()
    win = int(logits[0].argmax())
    dup = (win + 1 + V // 3) % V
    if dup > win:
        w[dup] = w[win]; b[dup] = b[win]
    logits = (h.float() @ w.float().t()).double()      # bf16 products are exact in fp32; accumulation order differs
    ref_logits = h.double() @ w.double().t() + b.double()
    want = ref_logits.argmax(1)
    top2 = ref_logits.topk(2, 1).values
    margin = top2[:, 0] - top2[:, 1]
    dev = "cuda"
    hd, wd, bd = h.to(dev), w.to(dev), b.to(dev)
    n = L.mmqg_vocab_workspace_bytes(R, V, H)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    toks = torch.full((R, 3), -1, dtype=torch.int64, device=dev)
    lib.check(L.mmqg_decode_step_argmax(hd.data_ptr(), wd.data_ptr(), bd.data_ptr(), R, V, H, ws.data_ptr(), n, toks[:, 1:].data_ptr(),
                                        3, stream()))
    torch.cuda.synchronize()
    got = toks[:, 1].cpu()
    assert (toks[:, 0] == -1).all() and (toks[:, 2] == -1).all()          # strided output only
    safe = (margin > 1e-4) | (margin == 0)            # exact ties (identical columns) must resolve to the lower index
    assert torch.equal(got[safe], want[safe]), (got[safe][:10], want[safe][:10])
    if dup > win:
        assert int(got[0]) == win

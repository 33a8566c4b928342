"""mmqg_lstm_seq_fwd / _bwd (one LSTM layer over a whole sequence on the tensor-core path) against torch.nn.LSTM in
fp64 on the CPU with identically bf16-rounded weights and inputs.  Tolerances: outputs 1e-2, gradients 3e-2 relative
per tensor (h_t is rounded to bf16 every step; d gates are bf16 operands of the hoisted products)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("T,B,I,H,state", [(7, 3, 20, 64, True), (33, 130, 300, 128, False), (12, 256, 512, 512, True), (1, 5, 44, 64, True)])
def test_lstm_seq_matches_torch_lstm(T, B, I, H, state):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import ops
    g = torch.Generator().manual_seed(T * 1000 + B)
    ref = torch.nn.LSTM(I, H).double()
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(bf(p.float()).double())
    x = bf(torch.randn(T, B, I, generator=g))
    h0 = 0.5 * torch.randn(B, H, generator=g) if state else torch.zeros(B, H)
    h0 = bf(h0)
    c0 = 0.5 * torch.randn(B, H, generator=g) if state else torch.zeros(B, H)
    gy = torch.randn(T, B, H, generator=g)
    ghn, gcn = torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    xd = x.double().requires_grad_(True)
    h0d, c0d = h0.double().requires_grad_(True), c0.double().requires_grad_(True)
    y, (hn, cn) = ref(xd, (h0d[None], c0d[None]))
    ((y * gy.double()).sum() + (hn[0] * ghn.double()).sum() + (cn[0] * gcn.double()).sum()).backward()
    dev = "cuda"
    w = [p.detach().float().to(dev).contiguous() for p in (ref.weight_ih_l0, ref.weight_hh_l0, ref.bias_ih_l0, ref.bias_hh_l0)]
    yk, hnk, cnk, ws = ops.lstm_seq_fwd(x.to(dev), *w, h0.to(dev) if state else None, c0.to(dev) if state else None)
    gk = ops.lstm_seq_bwd(gy.to(dev), ghn.to(dev), gcn.to(dev), w[1], ws, T, B, I, H)
    torch.cuda.synchronize()
    errs = {"y": rel(yk, y), "hn": rel(hnk, hn[0]), "cn": rel(cnk, cn[0]), "dx": rel(gk["dx"], xd.grad),
            "dw_ih": rel(gk["dw_ih"], ref.weight_ih_l0.grad), "dw_hh": rel(gk["dw_hh"], ref.weight_hh_l0.grad),
            "db": rel(gk["db_ih"], ref.bias_ih_l0.grad), "dh0": rel(gk["dh0"], h0d.grad), "dc0": rel(gk["dc0"], c0d.grad)}
    print(f"lstm_seq T={T} B={B} I={I} H={H}:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert torch.equal(gk["db_ih"], gk["db_hh"])
    assert max(errs["y"], errs["hn"], errs["cn"]) < 1e-2, errs
    assert max(v for k, v in errs.items() if k.startswith("d")) < 3e-2, errs


def test_lstm_seq_rejects_unsupported_shapes():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import _cabi, ops
    x = torch.randn(3, 2, 10, device="cuda")
    w_ih, w_hh = torch.randn(4 * 24, 10, device="cuda"), torch.randn(4 * 24, 24, device="cuda")
    b = torch.randn(4 * 24, device="cuda")
    with pytest.raises(_cabi.MmqgError):          # H = 24 is not a multiple of 64: loud error, no silent fallback
        ops.lstm_seq_fwd(x, w_ih, w_hh, b, b)

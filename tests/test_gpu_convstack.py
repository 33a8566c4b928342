"""f2 (SURVEY section 8 "next"): the conv stack of VideoConvLstmEncoder on raw frames, CUDA kernels of csrc/convstack.cu through
the C ABI.  (1) building blocks against the CPU oracle (oracle/convstack_oracle.py, itself pinned at 1e-9 to the reference's
own module) on seeded tensors, edge shapes included; (2) the drop-in module against fixtures produced by the reference's
VideoConvLstmEncoder (model/encoder.py:31-78): train-mode outputs, every gradient, BatchNorm running buffers, eval-mode
outputs.  Bar: 1e-3 relative (north_star, fp32), measured ~1e-6."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))

from conftest import load_golden  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import ops as o
    return o


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [
    dict(N=3, Cin=3, Cout=4, H=20, W=17, K=3, stride=1),
    dict(N=2, Cin=8, Cout=10, H=34, W=34, K=3, stride=1),       # fast path (convstack3.cu): even width, reference channels
    dict(N=3, Cin=3, Cout=4, H=21, W=30, K=3, stride=1),         # fast path, ragged 4-pixel groups (Wo = 28 + edge), 3 tiles of rows
    dict(N=5, Cin=4, Cout=6, H=110, W=110, K=3, stride=1),       # fast path at the reference's layer-2 size: several blocks per image
    dict(N=2, Cin=6, Cout=8, H=36, W=36, K=3, stride=1),
    dict(N=1, Cin=1, Cout=1, H=3, W=3, K=3, stride=1),            # single output pixel
    dict(N=2, Cin=5, Cout=7, H=23, W=19, K=5, stride=2),
])
def test_conv_blocks_match_oracle(ops, shape):
    from oracle import convstack_oracle as CO
    g = torch.Generator().manual_seed(5)
    N, Cin, Cout, H, W, K, s = (shape[k] for k in ("N", "Cin", "Cout", "H", "W", "K", "stride"))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) * 0.3
    b = torch.randn(Cout, generator=g) * 0.1
    sc, sh = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.2
    xr = (x.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)).requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    y_ref = torch.relu(CO.conv2d(xr, wr, br, s))
    y, stats = ops.conv_relu_fwd(x.cuda(), w.cuda(), b.cuda(), sc.cuda(), sh.cuda(), s)
    torch.cuda.synchronize()
    assert rel(y, y_ref) < 1e-5
    tot = stats.sum(0)
    assert rel(tot[:Cout], y_ref.sum((0, 2, 3))) < 1e-4 and rel(tot[Cout:], (y_ref ** 2).sum((0, 2, 3))) < 1e-4
    # backward of the convolution for a given d(conv output)
    dz = torch.randn(y_ref.shape, generator=g) * (y_ref.detach() > 0).float()
    conv_out = CO.conv2d(xr, wr, br, s)
    gx, gw, gb = torch.autograd.grad(conv_out, [xr, wr, br], dz.double())
    dw, db = ops.conv_bwd_w(x.cuda(), dz.cuda(), K, s, sc.cuda(), sh.cuda())
    dxn = ops.conv_bwd_x(dz.cuda(), w.cuda(), H, W, s)
    torch.cuda.synchronize()
    assert rel(dw, gw) < 1e-4 and rel(db, gb) < 1e-4 and rel(dxn, gx) < 1e-5


@pytest.mark.parametrize("shape", [dict(N=3, C=6, H=36, W=36, K=3), dict(N=2, C=10, H=32, W=32, K=3), dict(N=4, C=2, H=7, W=11, K=2)])
def test_batchnorm_pool_blocks_match_oracle(ops, shape):
    """Train-mode BatchNorm (statistics, running buffers) + MaxPool forward and backward; y >= 0 with many exact zeros
    (it is a ReLU output), so ties inside pooling windows are exercised: the first maximum takes the gradient, as in torch."""
    from oracle import convstack_oracle as CO
    g = torch.Generator().manual_seed(9)
    N, C_, H, W, K = (shape[k] for k in ("N", "C", "H", "W", "K"))
    y = torch.relu(torch.randn(N, C_, H, W, generator=g))
    gamma, beta = torch.rand(C_, generator=g) + 0.5, torch.randn(C_, generator=g) * 0.3
    gamma[0] = -gamma[0]                                             # a negative scale flips which element is the maximum
    rm, rv = torch.randn(C_, generator=g) * 0.1, torch.rand(C_, generator=g) + 0.5
    yr = y.double().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    bn_ref, rm_ref, rv_ref = CO.batchnorm(yr, gr, br, rm.double(), rv.double(), True)
    pooled_ref = torch.nn.functional.max_pool2d(bn_ref, K, K)      # torch's tie rule is the reference's
    yc = y.cuda()
    stats = torch.stack([yc.sum((0, 2, 3)), (yc * yc).sum((0, 2, 3))]).flatten().contiguous()
    rmc, rvc = rm.cuda(), rv.cuda()
    scale, shift, mean, invstd = ops.bn_finalize(stats, N * H * W, gamma.cuda(), beta.cuda(), 1e-5, 0.1, rmc, rvc)
    pooled, idx = ops.bn_maxpool_fwd(yc, scale, shift, K)
    torch.cuda.synchronize()
    assert rel(pooled, pooled_ref) < 1e-5
    assert rel(rmc, rm_ref) < 1e-5 and rel(rvc, rv_ref) < 1e-5
    dpool = torch.randn(pooled_ref.shape, generator=g)
    gy, gg, gb = torch.autograd.grad(pooled_ref, [yr, gr, br], dpool.double())
    dbn = ops.maxpool_bwd(dpool.cuda(), idx, H, W, K)
    dz, sums = ops.bn_relu_bwd(yc, mean, invstd, gamma.cuda(), dbn, train=True)
    dz2, sums2 = ops.bn_relu_pool_bwd(yc, mean, invstd, gamma.cuda(), dpool.cuda(), idx, K, train=True)     # the same in one pass
    torch.cuda.synchronize()
    assert rel(dz, gy * (y.double() > 0)) < 1e-4
    assert rel(sums[:C_], gb) < 1e-4 and rel(sums[C_:], gg) < 1e-4
    assert rel(dz2, gy * (y.double() > 0)) < 1e-4
    assert rel(sums2[:C_], gb) < 1e-4 and rel(sums2[C_:], gg) < 1e-4


@pytest.mark.parametrize("name", ["convstack_a", "convstack_b"])
def test_dropin_video_conv_lstm_encoder_matches_reference_fixture(ops, name):
    from model.encoder import VideoConvLstmEncoder
    fx = load_golden(name)
    enc = VideoConvLstmEncoder(3, 3, 1, fx["hidden"], fx["feat"])
    state = {k: (v.long() if "num_batches" in k else v) for k, v in fx["state0"].items()}
    enc.load_state_dict(state, strict=True)
    enc = enc.cuda().train()
    out = enc(fx["frames"].cuda())
    loss = (out * fx["proj"].cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert tuple(out.shape) == tuple(fx["out_train"].shape)
    assert rel(out, fx["out_train"]) < TOL
    assert abs(float(loss) - fx["loss"]) < TOL * max(1.0, abs(fx["loss"]))
    errs = {n_: rel(p_.grad, fx["grads"][n_]) for n_, p_ in enc.named_parameters()}
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"{name}: out rel err {rel(out, fx['out_train']):.2e}; worst grad rel err {worst}")
    assert worst[1] < TOL, errs
    for k, v in fx["bn_after"].items():
        got = enc.state_dict()[k]
        if "num_batches" in k:
            assert int(got) == int(v)
        else:
            assert rel(got, v) < 1e-5, k
    enc.eval()
    with torch.no_grad():
        out_eval = enc(fx["frames"].cuda())
    torch.cuda.synchronize()
    assert rel(out_eval, fx["out_eval"]) < TOL


def test_conv_stack_rejects_cpu_tensors():
    from model.encoder import VideoConvLstmEncoder
    enc = VideoConvLstmEncoder(3, 3, 1, 32, 40)
    with pytest.raises(RuntimeError):
        enc(torch.randn(1, 3, 2, 40, 40))

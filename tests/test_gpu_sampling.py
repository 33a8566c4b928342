"""The reference's 'sampling' decode strategy (evaluate.py:84-90: np.random.choice(V, p=softmax(logits)))
as an inverse-CDF draw: given the same uniform, the kernel must pick the token numpy's inverse CDF
picks (fp64), the draws must follow the softmax distribution, and the decode loop must be
deterministic in the seed.  'topk' in the reference is topk(1) == greedy."""
import numpy as np
import pytest
import torch

from mmqg.dims import Dims
from mmqg.synth import make_batch, make_params

pytestmark = pytest.mark.gpu


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def test_sample_rows_is_the_inverse_cdf_draw():
    need_gpu()
    from mmqg import ops
    g = torch.Generator().manual_seed(3)
    R, V = 300, 1003
    logits = (torch.randn(R, V, generator=g) * 3).cuda()
    for step in (0, 7):
        tok = ops.sample_rows(logits, seed=11, step=step).cpu().numpy()
        u = ops.sample_uniform(R, seed=11, step=step).double().cpu().numpy()
        p = torch.softmax(logits.double().cpu(), 1).numpy()
        cdf = np.cumsum(p, 1)
        ref = np.array([min(int(np.searchsorted(cdf[r], u[r], side="right")), V - 1) for r in range(R)])
        # identical except where u falls within fp32 rounding of a CDF step
        mism = np.nonzero(tok != ref)[0]
        for r in mism:
            lo, hi = sorted((tok[r], ref[r]))
            assert hi - lo <= 1 and abs(cdf[r, lo] - u[r]) < 1e-5, (r, tok[r], ref[r], u[r])
        assert len(mism) <= 3
        assert (u >= 0).all() and (u < 1).all()


def test_sample_rows_follows_the_softmax_distribution():
    need_gpu()
    from mmqg import ops
    V, N = 7, 40000
    row = torch.tensor([2.0, 0.5, -1.0, 1.0, 0.0, -3.0, 1.5])
    logits = row.repeat(N, 1).cuda()
    tok = ops.sample_rows(logits, seed=5, step=0).cpu()
    freq = torch.bincount(tok, minlength=V).double() / N
    p = torch.softmax(row.double(), 0)
    assert float((freq - p).abs().max()) < 0.01, (freq, p)
    tok2 = ops.sample_rows(logits, seed=5, step=0).cpu()
    assert torch.equal(tok, tok2)                                   # deterministic in (seed, step)
    assert not torch.equal(tok, ops.sample_rows(logits, seed=6, step=0).cpu())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_sampling_decode_loop(mode):
    need_gpu()
    from mmqg import engine
    d = Dims(B=16, T_t=9, T_v=3, T_q=4, V=120, E=52, H=64, L=2, H_a=24, H_v=64, F_v=40, TM=12, AM=5)
    eng = engine.TrainEngine(d, make_params(d, seed=121), mode=mode)
    b = eng.to_device(make_batch(d, seed=122))
    a = eng.greedy(b, 7, strategy="sampling", seed=1).cpu()
    a2 = eng.greedy(b, 7, strategy="sampling", seed=1).cpu()
    c = eng.greedy(b, 7, strategy="sampling", seed=2).cpu()
    g = eng.greedy(b, 7).cpu()
    assert torch.equal(a, a2) and not torch.equal(a, c)
    assert torch.equal(eng.greedy(b, 7, strategy="topk").cpu(), g)   # the reference's topk(1)
    assert a.min() >= 0 and a.max() < d.V
    assert not torch.equal(a, g)                                     # random-init logits are far from one-hot
    with pytest.raises(ValueError):
        eng.greedy(b, 7, strategy="beam")

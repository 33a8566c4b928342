"""Parity AT THE BENCHMARKED SHAPES (BASELINE.json configs[1], [3], [4]).

The other GPU suites use shapes the oracle finishes in a second; the benchmark runs B=256,
H=512, T_t=100 -- two m-tiles in the persistent recurrent kernels, six time chunks on three
layer streams, the multi-group loss head, split-K weight gradients -- paths no small test
reaches together.  These tests run the oracle (fp64, identically bf16-rounded weights) at
exactly those dimensions: ~15 s of host time at cfg-2, a few minutes at cfg-4.

Tolerances: bf16 mode loss 5e-3, gradients 5e-2 relative per tensor (same bar as
tests/test_gpu_bf16_mode.py; the fp32 mode carries the 1e-3 bar, checked here on the greedy
tokens of cfg-5 which must be exact wherever the oracle's own top-2 margin is resolvable).
"""
import os

import pytest
import torch

from mmqg.dims import config
from mmqg.synth import make_batch, make_params, round_params_bf16

pytestmark = pytest.mark.gpu

LOSS_TOL, GRAD_TOL = 5e-3, 5e-2


@pytest.fixture(scope="module")
def eng_mod():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import engine
    return engine


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def cfg2():
    """cfg-2 inputs and the fp64 oracle's loss / gradients for p = 0 (bf16-rounded weights)."""
    from oracle import mmqg_oracle as O
    d = config(2)
    params = make_params(d, seed=0)              # the weights and batch bench.py times
    batch = make_batch(d, seed=1234)
    loss_ref, grads_ref = O.loss_and_grads(round_params_bf16(params), batch, d.L, d.TM, d.AM, torch.float64)
    return d, params, batch, float(loss_ref), grads_ref


def _check(eng, loss, loss_ref, grads_ref, what):
    assert abs(loss - loss_ref) < LOSS_TOL * abs(loss_ref), (what, loss, loss_ref)
    errs = {k: rel(eng.grads[k], g) for k, g in grads_ref.items()}
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"{what}: loss {loss:.5f} vs oracle {loss_ref:.5f}; worst grad rel err {worst[1]:.3e} ({worst[0]}); "
          f"median {sorted(errs.values())[len(errs) // 2]:.3e}")
    assert worst[1] < GRAD_TOL, (what, errs)
    return errs


def test_cfg2_bf16_step_matches_oracle(eng_mod, cfg2):
    """The benchmarked step itself, dropout off: eager, then the same step replayed as a CUDA graph
    (what bench.py times) must give the same loss and gradients."""
    d, params, batch, loss_ref, grads_ref = cfg2
    eng = eng_mod.TrainEngine(d, params, mode="bf16")
    db = eng.to_device(batch)
    loss = float(eng.step(db))
    torch.cuda.synchronize()
    _check(eng, loss, loss_ref, grads_ref, "cfg-2 bf16 eager")
    eager = {k: v.clone() for k, v in eng.grads.items()}
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.step(db)
    for b in eng.grad_buckets:
        b.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert abs(float(eng.loss) - loss) < 1e-5 * abs(loss)
    for k in eager:                  # split-K partial sums are ordered deterministically: replays agree closely
        assert rel(eng.grads[k], eager[k]) < 1e-4, k


def test_cfg2_fp32_tc_step_matches_oracle(eng_mod):
    """The parity mode on the tensor cores (MMQG_MODE_FP32_TC: fp32 operands split into bf16 hi/lo, three tcgen05
    products per contraction) at the benchmarked shape, UNROUNDED weights, north_star's 1e-3 bar on the loss and on
    every gradient tensor; also reports its step time."""
    from oracle import mmqg_oracle as O
    d = config(2)
    params = make_params(d, seed=0)
    batch = make_batch(d, seed=1234)
    loss_ref, grads_ref = O.loss_and_grads(params, batch, d.L, d.TM, d.AM, torch.float64)
    eng = eng_mod.TrainEngine(d, params, mode="fp32_tc")
    db = eng.to_device(batch)
    loss = float(eng.step(db))
    torch.cuda.synchronize()
    errs = {k: rel(eng.grads[k], g) for k, g in grads_ref.items()}
    worst = max(errs.items(), key=lambda kv: kv[1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.step(db)
    e1.record()
    torch.cuda.synchronize()
    print(f"cfg-2 fp32_tc: loss {loss:.6f} vs oracle {float(loss_ref):.6f} (rel {abs(loss - float(loss_ref)) / abs(float(loss_ref)):.2e}); "
          f"worst grad rel err {worst[1]:.3e} ({worst[0]}); {e0.elapsed_time(e1) / 3:.2f} ms per eager step")
    assert abs(loss - float(loss_ref)) < 1e-3 * abs(float(loss_ref))
    assert worst[1] < 1e-3, errs


def test_cfg2_chunked_schedule_equals_unchunked(eng_mod, cfg2, monkeypatch):
    """MMQG_CHUNKS=6 (three layer streams, six time chunks -- the default) against one launch per layer."""
    d, params, batch, loss_ref, grads_ref = cfg2
    out = {}
    for chunks in ("1", "6"):
        monkeypatch.setenv("MMQG_CHUNKS", chunks)
        eng = eng_mod.TrainEngine(d, params, mode="bf16")
        loss = float(eng.step(eng.to_device(batch)))
        torch.cuda.synchronize()
        out[chunks] = (loss, {k: v.clone() for k, v in eng.grads.items()})
        _check(eng, loss, loss_ref, grads_ref, f"cfg-2 bf16 MMQG_CHUNKS={chunks}")
    assert abs(out["1"][0] - out["6"][0]) < 1e-4 * abs(out["1"][0])
    for k in out["1"][1]:
        assert rel(out["6"][1][k], out["1"][1][k]) < 2e-3, k


def test_cfg2_bf16_dropout_step_matches_oracle_with_same_masks(eng_mod, cfg2):
    """p = 0.2 (the reference's train-mode value, the configuration bench.py times): the counter-based
    masks of the step are exported and fed to the oracle."""
    from oracle import mmqg_oracle as O
    d, params, batch, _, _ = cfg2
    eng = eng_mod.TrainEngine(d, params, mode="bf16", dropout_p=0.2)
    eng.seed = 99
    masks = {k: v.cpu() for k, v in eng.dropout_masks().items()}
    loss_ref, grads_ref = O.loss_and_grads(round_params_bf16(params), batch, d.L, d.TM, d.AM, torch.float64, drop=masks)
    loss = float(eng.step(eng.to_device(batch)))
    torch.cuda.synchronize()
    _check(eng, loss, float(loss_ref), grads_ref, "cfg-2 bf16 p=0.2")


def test_cfg2_variable_lengths_two_mtiles(eng_mod):
    """Per-sample lengths at B=256, H=512 (two m-tiles, six chunks).  The oracle's per-sample loop over 256
    samples is too slow for the suite, so the check is a size-independent property of the path: cutting every
    sample to a common shorter length through ctx_len / tgt_len / n_frames must equal the uniform batch of that
    length (right-aligned time steps, shifted memory rows and zero-weight loss rows all exercised at mt = 0, 1)."""
    d = config(2)
    params = make_params(d, seed=0)
    batch = make_batch(d, seed=1234)
    cut_t, cut_q, cut_v = 61, 13, 7
    from mmqg.dims import Dims
    ds = Dims(**{**d.asdict(), "T_t": cut_t, "T_q": cut_q, "T_v": cut_v})
    short = {"context": batch["context"][:, :cut_t].contiguous(), "target": batch["target"][:, :cut_q].contiguous(),
             "frames": batch["frames"][:, :cut_v].contiguous(), "audio": batch["audio"][:, :cut_v].contiguous()}
    e_short = eng_mod.TrainEngine(ds, params, mode="bf16")
    l_short = float(e_short.step(e_short.to_device(short)))
    g_short = {k: v.clone() for k, v in e_short.grads.items()}
    lens = dict(batch)
    lens["ctx_len"] = torch.full((d.B,), cut_t, dtype=torch.int32)
    lens["tgt_len"] = torch.full((d.B,), cut_q, dtype=torch.int32)
    lens["n_frames"] = torch.full((d.B,), cut_v, dtype=torch.int32)
    eng = eng_mod.TrainEngine(d, params, mode="bf16")
    l_len = float(eng.step(eng.to_device(lens)))
    torch.cuda.synchronize()
    assert abs(l_len - l_short) < 2e-3 * abs(l_short), (l_len, l_short)
    for k in g_short:
        assert rel(eng.grads[k], g_short[k]) < 3e-2, k


@pytest.mark.skipif(os.environ.get("MMQG_SKIP_CFG4") == "1", reason="MMQG_SKIP_CFG4=1")
def test_cfg4_bf16_step_matches_oracle(eng_mod):
    """Long-context configuration (B=256, 64 frames, 400 tokens, 50k vocabulary) against the fp32 oracle
    (its own noise vs fp64 is 1e-6; fp32 keeps the run at about half a minute of host time)."""
    from oracle import mmqg_oracle as O
    d = config(4)
    params = make_params(d, seed=0)
    batch = make_batch(d, seed=1234)
    loss_ref, grads_ref = O.loss_and_grads(round_params_bf16(params), batch, d.L, d.TM, d.AM, torch.float32)
    eng = eng_mod.TrainEngine(d, params, mode="bf16")
    loss = float(eng.step(eng.to_device(batch)))
    torch.cuda.synchronize()
    _check(eng, loss, float(loss_ref), grads_ref, "cfg-4 bf16")


def test_cfg4_full_batch_properties(eng_mod):
    """cfg-4 at its full batch (B=256): batch-order invariance of the loss (a permutation of the samples),
    and the loss of the whole batch equals the mean of its two halves run separately."""
    d = config(4)
    params = make_params(d, seed=0)
    batch = make_batch(d, seed=1234)
    eng = eng_mod.TrainEngine(d, params, mode="bf16")
    l_all = float(eng.step(eng.to_device(batch)))
    g_all = {k: v.clone() for k, v in eng.grads.items()}
    perm = torch.randperm(d.B, generator=torch.Generator().manual_seed(3))
    l_perm = float(eng.step(eng.to_device({k: v[perm].contiguous() for k, v in batch.items()})))
    torch.cuda.synchronize()
    assert abs(l_perm - l_all) < 1e-4 * abs(l_all), (l_perm, l_all)
    for k in g_all:
        assert rel(eng.grads[k], g_all[k]) < 1e-2, k
    dh = config(4, d.B // 2)
    eh = eng_mod.TrainEngine(dh, params, mode="bf16")
    halves = []
    for i in range(2):
        sl = slice(i * dh.B, (i + 1) * dh.B)
        halves.append(float(eh.step(eh.to_device({k: v[sl].contiguous() for k, v in batch.items()}))))
    assert abs(0.5 * sum(halves) - l_all) < 1e-4 * abs(l_all), (halves, l_all)


@pytest.fixture(scope="module")
def cfg5():
    from oracle import mmqg_oracle as O
    d = config(5)
    gp = make_params(d, seed=0, bias_scale=0.1, out_weight_scale=10.0)      # input-sensitive weights, SURVEY section 0
    batch = make_batch(d, seed=1234)
    want, margins = O.greedy_decode(gp, batch, d.L, d.TM, d.AM, 30, torch.float64, return_margins=True)
    want16, margins16 = O.greedy_decode(round_params_bf16(gp), batch, d.L, d.TM, d.AM, 30, torch.float64, return_margins=True)
    return d, gp, batch, want, margins, want16, margins16


def test_cfg5_fp32_greedy_token_exact(eng_mod, cfg5):
    """Greedy decode, B=1024 x 30 tokens, fp32 mode: every token equals the oracle's up to the first
    position of a row whose oracle top-1/top-2 margin is within fp32 noise of the logits (1e-4)."""
    d, gp, batch, want, margins, _, _ = cfg5
    eng = eng_mod.TrainEngine(d, gp, mode="fp32")
    toks = eng.greedy(eng.to_device(batch), 30).cpu()
    safe = (margins > 1e-4).long().cumprod(1).bool()
    print(f"cfg-5 fp32: {int((toks == want).sum())}/{toks.numel()} tokens equal; {float(safe.float().mean()):.4f} of positions "
          f"resolvable; oracle margins min {float(margins.min()):.2e} median {float(margins.median()):.2e}; "
          f"{len({tuple(r) for r in want.tolist()})} distinct sequences")
    assert safe.float().mean() > 0.9
    assert torch.equal(toks[safe], want[safe])
    assert len({tuple(r) for r in want.tolist()}) > 8          # sequences are input-dependent


def test_cfg5_bf16_greedy_match_rate(eng_mod, cfg5):
    """The tensor-core decode path against the oracle run with the same bf16-rounded weights: a row may
    leave the oracle's path only at a position whose margin is small (< 0.5 on logits of std ~2.5), and the
    rows must agree on most of their tokens."""
    d, gp, batch, _, _, want16, margins16 = cfg5
    eng = eng_mod.TrainEngine(d, gp, mode="bf16")
    toks = eng.greedy(eng.to_device(batch), 30).cpu()
    first_bad = []
    for b in range(d.B):
        diff = (toks[b] != want16[b]).nonzero()
        if diff.numel():
            first_bad.append(float(margins16[b, int(diff[0])]))
    rate = float((toks == want16).float().mean())
    print(f"cfg-5 bf16: token match rate {rate:.4f}; {d.B - len(first_bad)}/{d.B} rows exact; "
          f"largest margin at a first divergence {max(first_bad) if first_bad else 0.0:.3f}")
    assert not first_bad or max(first_bad) < 0.5
    assert rate > 0.6

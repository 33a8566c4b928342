"""Persistent decoder BPTT kernel (csrc/dec_persist_bwd.cu, selected with MMQG_DEC_BWD_PERSIST=1): one cooperative launch
for the backward of all teacher-forced decoder steps (autograd twin of reference model/decoder.py:74-107 under
train.py:177).  Checked against the fp64 oracle run with identically bf16-rounded weights at the bf16 tolerances of
tests/test_gpu_bf16_mode.py, with dropout (exported masks), and against the default launch-per-phase BPTT loop."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))

pytestmark = pytest.mark.gpu

from mmqg.dims import Dims, config  # noqa: E402
from mmqg.synth import make_batch, make_params, round_params_bf16  # noqa: E402

LOSS_TOL = 5e-3      # bf16 operands, fp32 accumulation (see tests/test_gpu_bf16_mode.py)
GRAD_TOL = 5e-2


@pytest.fixture(scope="module")
def eng_mod():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import engine
    return engine


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


CASES = [
    dict(B=8, T_t=17, T_v=5, T_q=7, V=1003, E=300, H=512, L=3, H_a=128, H_v=512, F_v=2048, TM=283, AM=101),
    dict(B=130, T_t=5, T_v=2, T_q=3, V=520, E=52, H=64, L=2, H_a=24, H_v=128, F_v=40, TM=11, AM=6),      # ragged last group, L = 2
    dict(B=70, T_t=9, T_v=4, T_q=5, V=300, E=36, H=128, L=1, H_a=16, H_v=64, F_v=24, TM=14, AM=7),       # single layer
]


@pytest.mark.parametrize("p_drop", [0.0, 0.2])
@pytest.mark.parametrize("cfg", CASES)
def test_persistent_bptt_matches_oracle_and_default_loop(eng_mod, cfg, p_drop, monkeypatch):
    from oracle import mmqg_oracle as O
    d = Dims(**cfg)
    params = make_params(d, seed=61)
    batch = make_batch(d, seed=62)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MMQG_DEC_BWD_PERSIST", mode)
        eng = eng_mod.TrainEngine(d, params, mode="bf16", dropout_p=p_drop)
        eng.seed = 4321
        masks = {k: v.cpu() for k, v in eng.dropout_masks().items()} if p_drop > 0 else None
        loss = float(eng.step(eng.to_device(batch)))
        torch.cuda.synchronize()
        out[mode] = (loss, {k: v.clone() for k, v in eng.grads.items()}, masks)
    loss_ref, grads_ref = O.loss_and_grads(round_params_bf16(params), batch, d.L, d.TM, d.AM, torch.float64, drop=out["1"][2])
    loss, grads, _ = out["1"]
    assert abs(loss - float(loss_ref)) < LOSS_TOL * abs(float(loss_ref)), (loss, float(loss_ref))
    errs = {k: rel(grads[k], g) for k, g in grads_ref.items()}
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"persistent BPTT p={p_drop}: worst grad rel err {worst}")
    assert worst[1] < GRAD_TOL, errs
    # the same masks in both modes (same seed, same call counter): the two BPTT implementations must agree far inside the bar
    assert abs(out["0"][0] - out["1"][0]) < 1e-5 * abs(out["0"][0])
    for k in grads:
        assert rel(grads[k], out["0"][1][k]) < 2e-2, k


def test_persistent_bptt_cfg2_graph_replay(eng_mod, monkeypatch):
    """The benchmarked shape (B = 256: four 64-row groups x 32 CTAs) against the default loop, eager and replayed as a graph."""
    d = config(2)
    params = make_params(d, seed=0)
    batch = make_batch(d, seed=1)
    monkeypatch.setenv("MMQG_DEC_BWD_PERSIST", "0")
    ref = eng_mod.TrainEngine(d, params, mode="bf16")
    l_ref = float(ref.step(ref.to_device(batch)))
    g_ref = {k: v.clone() for k, v in ref.grads.items()}
    monkeypatch.setenv("MMQG_DEC_BWD_PERSIST", "1")
    eng = eng_mod.TrainEngine(d, params, mode="bf16")
    db = eng.to_device(batch)
    loss = float(eng.step(db))
    torch.cuda.synchronize()
    assert abs(loss - l_ref) < 1e-5 * abs(l_ref)
    for k in g_ref:
        assert rel(eng.grads[k], g_ref[k]) < 2e-2, k
    eager = {k: v.clone() for k, v in eng.grads.items()}
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.step(db)
    for b in eng.grad_buckets:
        b.zero_()
    g.replay()
    torch.cuda.synchronize()
    for k in eager:
        assert rel(eng.grads[k], eager[k]) < 1e-4, k

"""Run the non-headline BASELINE.json configs once on the GPU: cfg 4 (long context: T_t=400, T_v=64, V=50k)
train step in bf16 and fp32, cfg 5 (greedy decode, B=1024, 30 tokens).  Prints samples/s; results are
quoted in DESIGN.md."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg.dims import config
from mmqg.engine import TrainEngine
from mmqg.synth import make_batch, make_params

def timed(fn, iters):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

which = sys.argv[1:] or ["4bf16", "4fp32", "5"]
if "4bf16" in which or "4fp32" in which:
    d = config(4)
    params = make_params(d, seed=0); batch = make_batch(d, seed=1)
    for mode in ("bf16", "fp32"):
        if f"4{mode}" not in which: continue
        eng = TrainEngine(d, params, mode=mode); db = eng.to_device(batch)
        ms = timed(lambda: eng.step(db), 3 if mode == "fp32" else 10)
        print(f"cfg4 {mode}: {ms:.2f} ms/step  {d.B / ms * 1e3:.0f} samples/s  loss {float(eng.loss):.4f}  ws {eng.ws.numel()/2**30:.2f} GiB", flush=True)
        del eng; torch.cuda.empty_cache()
if "5" in which:
    d = config(5)
    params = make_params(d, seed=0, bias_scale=0.1, out_weight_scale=10.0); batch = make_batch(d, seed=1)
    eng = TrainEngine(d, params, mode="fp32"); db = eng.to_device(batch)
    ms = timed(lambda: eng.greedy(db, 30), 3)
    toks = eng.greedy(db, 30)
    print(f"cfg5 greedy fp32: {ms:.2f} ms per batch of {d.B}  {d.B / ms * 1e3:.0f} samples/s  distinct sequences {len({tuple(r) for r in toks.tolist()})}", flush=True)
    e16 = TrainEngine(d, params, mode="bf16"); db16 = e16.to_device(batch)
    ms16 = timed(lambda: e16.greedy(db16, 30), 5)
    t16 = e16.greedy(db16, 30)
    same = (t16 == toks)
    first_diff = torch.where(same.all(1), torch.full((d.B,), 30, device=same.device), (~same).float().argmax(1))
    print(f"cfg5 greedy bf16: {ms16:.2f} ms per batch of {d.B}  {d.B / ms16 * 1e3:.0f} samples/s  rows identical to fp32 "
          f"{int(same.all(1).sum())}/{d.B}, tokens identical {float(same.float().mean()):.3f}, mean first difference at step "
          f"{float(first_diff.float().mean()):.1f}", flush=True)

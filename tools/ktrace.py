"""Timeline of the persistent recurrent kernels inside one step (debug hook mmqg_debug_ktrace):
start/end of every launch relative to the first, per layer and chunk, plus the busy/idle
structure of the text-encoder phases.  python tools/ktrace.py [--graph]"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg import _cabi  # noqa: E402
from mmqg.dims import config  # noqa: E402
from mmqg.engine import TrainEngine  # noqa: E402
from mmqg.synth import make_batch, make_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
d = config(2)
eng = TrainEngine(d, make_params(d, seed=0), mode="bf16", dropout_p=0.2)
b = eng.to_device(make_batch(d, seed=1))
for _ in range(3):
    eng.step(b)
torch.cuda.synchronize()
buf = torch.zeros(1 + 3 * 256, dtype=torch.int64, device="cuda")
L = C.CDLL(_cabi.LIB_PATH)
L.mmqg_debug_ktrace.argtypes = [C.c_void_p]
L.mmqg_debug_ktrace(buf.data_ptr())
if a.graph:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.step(b)
    buf.zero_()
    g.replay()
else:
    eng.step(b)
torch.cuda.synchronize()
L.mmqg_debug_ktrace(None)
t = buf.cpu()
n = int(t[0])
rows = sorted(((int(t[1 + 3 * i]), int(t[2 + 3 * i]), int(t[3 + 3 * i])) for i in range(n)), key=lambda r: r[1])
t0 = rows[0][1]
print(f"{n} persistent launches; times in us from the first start")
for tag, s, e in rows:
    kind = "bwd" if tag >= 1000 else "fwd"
    tg = tag % 1000
    name = "video" if tg == 900 else f"L{tg // 16} c{tg % 16}"
    print(f"{kind} {name:8s} start {(s - t0) / 1e3:9.1f}  end {(e - t0) / 1e3:9.1f}  dur {(e - s) / 1e3:7.1f}")
for kind, lo in (("fwd", 0), ("bwd", 1000)):
    ev = [(s, e) for tag, s, e in rows if lo <= tag < lo + 900]
    if not ev:
        continue
    a0, a1 = min(s for s, _ in ev), max(e for _, e in ev)
    busy = sum(e - s for s, e in ev)
    print(f"text {kind}: window {(a1 - a0) / 1e3:.1f} us, summed kernel time {busy / 1e3:.1f} us, mean concurrency {busy / (a1 - a0):.2f}")

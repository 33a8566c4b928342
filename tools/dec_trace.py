"""Per-step timeline of the persistent decoder kernel (debug hook mmqg_debug_dec_trace): where one decoder step of
CTA 0 spends its time.  python tools/dec_trace.py"""
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

d = config(2)
eng = TrainEngine(d, make_params(d, seed=0), mode="bf16", dropout_p=0.2)
b = eng.to_device(make_batch(d, seed=1))
for _ in range(3):
    eng.step(b)
torch.cuda.synchronize()
buf = torch.zeros(12 * d.T_q, dtype=torch.int64, device="cuda")
L = C.CDLL(_cabi.LIB_PATH)
L.mmqg_debug_dec_trace.argtypes = [C.c_void_p]
L.mmqg_debug_dec_trace(buf.data_ptr())
eng.step(b)
torch.cuda.synchronize()
L.mmqg_debug_dec_trace(None)
t = buf.cpu().view(d.T_q, 12)
names = ["start", "P1 epi", "scores seen", "ctx published", "acc L0", "acc L1", "acc L2", "end"]
print("step  " + "  ".join(f"{n:>13s}" for n in names[1:]) + "   (us since the step's start; last column = step length)")
tot = []
for s in range(d.T_q):
    r = t[s]
    if int(r[0]) == 0:
        continue
    print(f"{s:4d}  " + "  ".join(f"{(int(r[i]) - int(r[0])) / 1e3:13.2f}" for i in range(1, 8)))
    tot.append((int(r[7]) - int(r[0])) / 1e3)
    if int(r[8]):
        print(f"        attention: scores loaded {(int(r[8]) - int(r[0])) / 1e3:.2f}  softmax done {(int(r[9]) - int(r[0])) / 1e3:.2f}  "
              f"2nd sample starts {(int(r[10]) - int(r[0])) / 1e3:.2f}  waited for memory chunks {int(r[11]) / 1e3:.2f}")
if tot:
    print(f"mean step {sum(tot) / len(tot):.2f} us; whole window {(int(t[-1][7]) - int(t[0][0])) / 1e3:.1f} us")

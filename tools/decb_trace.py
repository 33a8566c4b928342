"""Per-step timeline of the persistent decoder BPTT kernel (debug hook mmqg_debug_decb_trace): where one step of CTA 0
spends its time.  python tools/decb_trace.py"""
import ctypes as C
import os
import sys

os.environ["MMQG_DEC_BWD_PERSIST"] = "1"
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
buf = torch.zeros(24 * d.T_q, dtype=torch.int64, device="cuda")
L = C.CDLL(_cabi.LIB_PATH)
L.mmqg_debug_decb_trace.argtypes = [C.c_void_p]
L.mmqg_debug_decb_trace(buf.data_ptr())
eng.step(b)
torch.cuda.synchronize()
L.mmqg_debug_decb_trace(None)
t = buf.cpu().view(d.T_q, 24)
names = ["start", "top operands", "top acc", "top publ", "L1 acc", "L1 publ", "L0 acc", "L0 publ", "ctx acc", "ctx publ", "dctx seen",
         "attn done", "TMA: dG_top seen", "TMA: S_top issued", "TMA: ds seen", "TMA: S_s issued"]
print("step  " + " ".join(f"{n:>9s}" for n in ["t.oper", "t.acc", "t.publ", "L1acc", "L1publ", "L0acc", "L0publ", "ctxacc", "ctxpubl", "dctxseen", "attndone", "gTseen", "STiss", "dsseen", "SSiss"]))
tot = []
for s in range(d.T_q - 1, -1, -1):
    r = t[s]
    if int(r[0]) == 0:
        continue
    print(f"{s:4d}  " + " ".join(f"{(int(r[i]) - int(r[0])) / 1e3:9.2f}" for i in range(1, 16)))
    tot.append((int(r[11]) - int(r[0])) / 1e3)
    print(f"        attention, first sample: operands loaded {(int(r[16]) - int(r[0])) / 1e3:.2f}  chunks done {(int(r[17]) - int(r[0])) / 1e3:.2f} "
          f"softmax backward done {(int(r[20]) - int(r[0])) / 1e3:.2f}")
if tot:
    print(f"mean step {sum(tot) / len(tot):.2f} us; whole window {(int(t[0][11]) - int(t[-1][0])) / 1e3:.1f} us")

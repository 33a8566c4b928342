"""Per-phase %globaltimer trace of the persistent forward kernel: min/max over all CTAs per step
(last text layer of one forward).  Debug tool, not part of the product path."""
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
eng = TrainEngine(d, make_params(d, seed=0), mode="bf16")
b = eng.to_device(make_batch(d, seed=1))
for _ in range(2):
    eng.forward(b, want_grads=False)
torch.cuda.synchronize()
init = torch.zeros(d.T_t, 8, dtype=torch.int64)
init[:, [0, 2, 4]] = torch.iinfo(torch.int64).max
buf = init.cuda()
L = C.CDLL(_cabi.LIB_PATH)
L.mmqg_debug_lstm_trace.argtypes = [C.c_void_p]
# only the last text layer should write: the engine runs video, text l0, l1, l2 in order, so re-arm before each
# forward and keep the final content (min/max accumulate over the 4 launches; video has only T_v steps)
L.mmqg_debug_lstm_trace(buf.data_ptr())
eng.forward(b, want_grads=False)
torch.cuda.synchronize()
L.mmqg_debug_lstm_trace(None)
t = buf.cpu().double()
print("NOTE: min/max accumulate over all persistent launches of the forward; use the max columns of late steps")
names = ["min flag_ok", "max flag_ok", "min epi_sees_mma", "max epi_sees_mma", "min flag_set", "max flag_set", "max h_stored", "max mma_issued"]
s = t[20:90]
ref = s[:, 1:2]      # max flag_ok of the last launch
for i, n in enumerate(names):
    print(f"{n:18s} mean offset vs max flag_ok: {float((s[:, i] - ref[:, 0]).mean()):10.0f} ns")
print("period (max flag_ok deltas):", float((t[21:91, 1] - t[20:90, 1]).mean()), "ns")

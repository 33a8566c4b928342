"""Per-phase %globaltimer trace of the split-K cluster BPTT kernel (lstm_seq_bwd4_kernel), CTA (0,0), per step:
0 poll start, 1 counter seen, 2 loads issued, 3 epilogue operands ready (prefetch + transposes done), 4 accumulator seen,
5 partials sent, 6 partials received, 7 dG stored, 8 CTA barrier passed, 9 fence + counter bump done.
MMQG_CHUNKS=1 (T=100 per launch; the last layer's launch -- text layer 0 -- overwrites the earlier ones).  Debug tool."""
import ctypes as C
import os
import sys

os.environ.setdefault("MMQG_CHUNKS", "1")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg import _cabi  # noqa: E402
from mmqg.dims import config  # noqa: E402
from mmqg.engine import TrainEngine  # noqa: E402
from mmqg.synth import make_batch, make_params  # noqa: E402

import dataclasses  # noqa: E402
d = dataclasses.replace(config(2), L=int(os.environ.get("LAYERS", "1")))      # one text layer: min/max columns then come from a single launch
eng = TrainEngine(d, make_params(d, seed=0), mode="bf16", dropout_p=float(os.environ.get("DROP_P", "0.2")))
b = eng.to_device(make_batch(d, seed=1))
for _ in range(2):
    eng.step(b)
torch.cuda.synchronize()
eng.forward(b, want_grads=True)
torch.cuda.synchronize()
init = torch.zeros(d.T_t, 16, dtype=torch.int64)
init[:, [11, 13]] = torch.iinfo(torch.int64).max
buf = init.cuda()
L = C.CDLL(_cabi.LIB_PATH)
L.mmqg_debug_lstm_trace.argtypes = [C.c_void_p]
L.mmqg_debug_lstm_trace(buf.data_ptr())
eng.backward(b, 0)
torch.cuda.synchronize()
L.mmqg_debug_lstm_trace(None)
t = buf.cpu().double()
names = ["poll start", "counter seen", "loads issued", "epilogue operands ready", "accumulator seen", "partials sent",
         "partials received", "dG stored", "barrier passed", "fence + bump done"]
s = t[20:90]
ref = s[:, 1:2]
for i, n in enumerate(names):
    print(f"{n:26s} mean offset vs counter seen: {float((s[:, i:i+1] - ref).mean()):9.0f} ns")
print("period:", float((t[20:90, 1] - t[21:91, 1]).mean()), "ns")
for i, n in ((10, "bump done, LAST CTA of the m-tile"), (11, "bump done, FIRST CTA"), (12, "counter seen, LAST CTA"), (13, "counter seen, FIRST CTA")):
    print(f"{n:34s} mean offset vs CTA 0 counter seen: {float((s[:, i:i+1] - ref).mean()):9.0f} ns")
for tt in range(40, 43):
    print("step", tt, [int(t[tt, i] - t[tt, 1]) for i in range(10)])

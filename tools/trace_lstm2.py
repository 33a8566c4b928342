"""Per-phase %globaltimer trace of the two-chain persistent forward kernel (lstm_seq_fwd2_kernel), CTA (0,0):
per step and chain -- 0 counter seen, 1 loads issued, 2 MMAs committed, 3 epilogue sees the accumulator,
4 h stored + chain barrier, 5 fence + counter bump done, 6 saved activations stored.  MMQG_CHUNKS=1 (T=100 per launch);
the last text layer's launch overwrites the earlier ones.  Debug tool, not part of the product path."""
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

d = config(2)
eng = TrainEngine(d, make_params(d, seed=0), mode="bf16", dropout_p=float(os.environ.get("DROP_P", "0.2")))
b = eng.to_device(make_batch(d, seed=1))
for _ in range(2):
    eng.forward(b, want_grads=False)
torch.cuda.synchronize()
buf = torch.zeros(d.T_t, 16, dtype=torch.int64, device="cuda")
L = C.CDLL(_cabi.LIB_PATH)
L.mmqg_debug_lstm_trace.argtypes = [C.c_void_p]
L.mmqg_debug_lstm_trace(buf.data_ptr())
eng.forward(b, want_grads=False)
torch.cuda.synchronize()
L.mmqg_debug_lstm_trace(None)
t = buf.cpu().double().view(d.T_t, 2, 8)
names = ["counter seen", "loads issued", "MMAs committed", "epilogue sees acc", "h stored + barrier", "fence + bump done", "saved acts stored"]
s = t[20:90]
ref = s[:, 0, 0:1]
for q in range(2):
    for i, n in enumerate(names):
        print(f"chain {q} {n:20s} mean offset vs chain 0 counter seen: {float((s[:, q, i:i+1] - ref).mean()):9.0f} ns")
print("period chain 0:", float((t[21:91, 0, 0] - t[20:90, 0, 0]).mean()), "ns; chain 1:", float((t[21:91, 1, 0] - t[20:90, 1, 0]).mean()), "ns")
print("poll started (chain 0 / 1) vs chain 0 counter seen:", float((s[:, 0, 7:8] - ref).mean()), float((s[:, 1, 7:8] - ref).mean()))
for tt in range(40, 44):
    print("step", tt, [[int(t[tt, q, i] - t[tt, 0, 0]) for i in range(7)] for q in range(2)])

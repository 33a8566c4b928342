"""Per-step time of the persistent recurrent kernels vs the number of CTAs streaming at once:
B=128 (one m-tile, 32 CTAs) against B=256 (two m-tiles, 64 CTAs), layers run one after the other
(MMQG_CHUNKS=1).  If the backward kernel were bound by aggregate L2 bandwidth its step time would
drop with half the CTAs; bound by per-SM ingest it stays.  Uses the in-library probe (class 7)."""
import ctypes as C
import os
import sys

os.environ["MMQG_CHUNKS"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg import _cabi  # noqa: E402
from mmqg.dims import config  # noqa: E402
from mmqg.engine import TrainEngine  # noqa: E402
from mmqg.synth import make_batch, make_params  # noqa: E402


def probe(fn):
    L = _cabi.lib()
    _cabi.check(L.mmqg_probe_start(7))
    fn()
    torch.cuda.synchronize()
    ms, n, fl, by = C.c_double(), C.c_ulonglong(), C.c_double(), C.c_double()
    _cabi.check(L.mmqg_probe_stop(C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)))
    return ms.value, n.value


for B in (128, 256):
    d = config(2, B)
    eng = TrainEngine(d, make_params(d, seed=0), mode="bf16")
    b = eng.to_device(make_batch(d, seed=1))
    for _ in range(2):
        eng.step(b)
    torch.cuda.synchronize()
    f_ms, f_n = probe(lambda: eng.forward(b, True))
    t_ms, t_n = probe(lambda: eng.step(b))
    steps = 3 * d.T_t + d.T_v
    print(f"B={B}: forward kernels {f_ms * 1e3 / steps:.2f} us/step ({f_n} launches), backward kernels "
          f"{(t_ms - f_ms) * 1e3 / steps:.2f} us/step ({t_n - f_n} launches)", flush=True)

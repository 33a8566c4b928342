"""f2 measurement: forward + backward of the conv stack (csrc/convstack.cu) at the reference's frame size (3 x 112 x 112,
config.py) for N frames in one call, CUDA-event timed, with the in-library probe's algorithmic bytes per kernel class against
the measured HBM peak.  python tools/bench_convstack.py [N]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg import _cabi  # noqa: E402
from mmqg import functional as MF  # noqa: E402
from model.encoder import VideoConvLstmEncoder  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2560            # 256 samples x 10 salient frames
enc = VideoConvLstmEncoder(3, 3, 1, 512, 1000).cuda().train()
x = torch.randn(N, 3, 112, 112, device="cuda")
params = [p for n, p in enc.named_parameters() if not n.startswith("lstm")]


def step():
    for p in params:
        p.grad = None
    y = MF.conv_stack(x, enc)
    y.backward(torch.ones_like(y))


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
L = _cabi.lib()
_cabi.check(L.mmqg_probe_start(0))
step()
torch.cuda.synchronize()
tms, n, fl, by = C.c_double(), C.c_ulonglong(), C.c_double(), C.c_double()
_cabi.check(L.mmqg_probe_stop(C.byref(tms), C.byref(n), C.byref(fl), C.byref(by)))
peak = 6547.8
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak = float(pk.get("hbm_gbs", pk.get("hbm_gb_s", peak)))
except Exception:
    pass
gbs = by.value / (tms.value * 1e-3) / 1e9
print(json.dumps({"what": "conv stack fwd+bwd (4 x conv3x3+ReLU+BatchNorm(train), 2 x MaxPool3), 3x112x112 frames", "frames": N,
                  "ms_per_call": ms, "frames_per_s": N / (ms * 1e-3), "kernel_launches": int(n.value), "summed_kernel_ms": tms.value,
                  "algorithmic_gbytes": by.value / 1e9, "achieved_gb_s": gbs, "hbm_peak_gb_s": peak, "frac": gbs / peak,
                  "gflop": fl.value / 1e9}))

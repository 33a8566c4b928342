"""Where the step time goes: CUDA events at the section boundaries of eager steps (debug hook in
engine_bf16.cu).  python tools/sections.py [--config 2] [--dropout 0.2]"""
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

NAMES = ["start", "weights packed", "encoders fwd", "decoder hoisted fwd", "decoder step loop fwd", "loss head (fwd+bwd)",
         "backward start", "decoder BPTT loop", "text BPTT (main stream)", "end (joins + embedding scatter)"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    d = config(a.config)
    eng = TrainEngine(d, make_params(d, seed=0), mode="bf16", dropout_p=a.dropout)
    batch = eng.to_device(make_batch(d, seed=1))
    lib = _cabi.lib()
    lib.mmqg_debug_sections.argtypes = [C.c_int]
    lib.mmqg_debug_section_times.argtypes = [C.POINTER(C.c_float)]
    for _ in range(3):
        eng.step(batch)
    torch.cuda.synchronize()
    lib.mmqg_debug_sections(1)
    acc = [0.0] * 10
    for _ in range(a.reps):
        eng.step(batch)
        torch.cuda.synchronize()
        out = (C.c_float * 10)()
        lib.mmqg_debug_section_times(out)
        for i in range(10):
            acc[i] += out[i] / a.reps
    lib.mmqg_debug_sections(0)
    prev = 0.0
    for i in range(1, 10):
        print(f"{NAMES[i]:34s} +{acc[i] - prev:7.3f} ms   (t = {acc[i]:7.3f})")
        prev = acc[i]


if __name__ == "__main__":
    main()

"""Per-kernel-class roofline table of one bf16 train step (cfg-2): the in-library CUDA-event probe
(mmqg_probe_start/stop) around every launch of one class at a time, on the launching stream, in an
eager step; achieved = booked algorithmic FLOPs (or bytes) / summed kernel time, against the
measured peaks of MEASURED_PEAKS.json (sustained bf16 TF/s, HBM GB/s).
    python tools/roofline_table.py [--config 2] [--markdown]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg import _cabi  # noqa: E402
from mmqg.dims import config  # noqa: E402
from mmqg.engine import TrainEngine  # noqa: E402
from mmqg.synth import make_batch, make_params  # noqa: E402

NAMES = {1: ("per-step products of the decoder (gemm_tc_kernel, B x 4H x K)", "tensor"),
         2: ("hoisted whole-sequence products (gemm_tc / gemm_tc_persist)", "tensor"),
         3: ("LSTM cell kernels of the decoder (lstm_pointwise_*)", "hbm"),
         4: ("attention step forward / backward (attn_*_fast_kernel)", "hbm"),
         5: ("log-softmax + NLL + dlogits rows (nll_rows_bf16)", "hbm"),
         6: ("embedding gather / scatter-add", "hbm"),
         7: ("persistent recurrent kernels (lstm_seq_fwd/bwd_kernel)", "tensor")}


def peaks():
    tf, gb = 1383.1, 6547.8          # fallback = the values the driver measured on this pool
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        tf = float(p.get("bf16_tflops_sustained", tf))
        gb = float(p.get("hbm_gbs", gb))
    except (OSError, ValueError):
        pass
    return tf, gb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--markdown", action="store_true")
    a = ap.parse_args()
    d = config(a.config)
    eng = TrainEngine(d, make_params(d, seed=0), mode="bf16", dropout_p=0.2)
    b = eng.to_device(make_batch(d, seed=1))
    for _ in range(3):
        eng.step(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.step(b)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 5
    L = _cabi.lib()
    tf_peak, gb_peak = peaks()
    rows = []
    for k, (name, bound) in NAMES.items():
        _cabi.check(L.mmqg_probe_start(k))
        for _ in range(2):
            eng.step(b)
        torch.cuda.synchronize()
        ms, n, fl, by = C.c_double(), C.c_ulonglong(), C.c_double(), C.c_double()
        _cabi.check(L.mmqg_probe_stop(C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)))
        if not n.value or ms.value <= 0:
            continue
        if bound == "tensor":
            ach = fl.value / (ms.value * 1e-3) / 1e12
            rows.append((name, n.value // 2, 1e3 * ms.value / n.value, ms.value / 2, f"{ach:.1f} TFLOP/s", ach / tf_peak))
        else:
            ach = by.value / (ms.value * 1e-3) / 1e9
            rows.append((name, n.value // 2, 1e3 * ms.value / n.value, ms.value / 2, f"{ach:.0f} GB/s", ach / gb_peak))
    print(f"eager step {step_ms:.2f} ms; peaks: {tf_peak:.1f} TFLOP/s bf16 sustained, {gb_peak:.1f} GB/s HBM")
    if a.markdown:
        print("| kernel class | launches/step | avg us | summed ms/step | achieved | of peak |")
        print("|---|---|---|---|---|---|")
        for r in rows:
            print(f"| {r[0]} | {r[1]} | {r[2]:.1f} | {r[3]:.2f} | {r[4]} | {100 * r[5]:.1f} % |")
    else:
        for r in rows:
            print(f"{r[0]:72s} n={r[1]:4d} avg {r[2]:7.1f} us  sum {r[3]:5.2f} ms  {r[4]:>14s}  {100 * r[5]:5.1f} % of peak")


if __name__ == "__main__":
    main()

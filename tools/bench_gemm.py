"""Micro-benchmark of the tcgen05 GEMM at the shapes the path uses (run under gpurun).
Prints achieved TFLOP/s per shape next to torch.matmul (cuBLAS) on the same operands."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg import ops  # noqa: E402

SHAPES = [  # name, M, N, K, a_mn, b_mn
    ("text hoisted X W_ih^T", 25600, 2048, 512, False, False),
    ("text dW = dG^T X", 2048, 512, 25600, True, True),
    ("text dX = dG W", 25600, 512, 2048, False, True),
    ("step h W_hh^T", 256, 2048, 512, False, False),
    ("step dG W_hh", 256, 512, 2048, False, True),
    ("vocab logits chunk", 1664, 10000, 512, False, False),
    ("vocab dW chunk", 10000, 512, 1664, True, True),
    ("square 4096", 4096, 4096, 4096, False, False),
]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, M, N, K, amn, bmn in SHAPES:
        A = torch.randn((K, M) if amn else (M, K), device="cuda", generator=g).bfloat16()
        B = torch.randn((K, N) if bmn else (N, K), device="cuda", generator=g).bfloat16()
        out = torch.empty(M, N, device="cuda")
        t = timeit(lambda: ops.gemm_bf16(A, B, amn, bmn, out=out))
        Am = A.t() if amn else A
        Bm = B if bmn else B.t()
        tt = timeit(lambda: torch.matmul(Am, Bm))
        fl = 2.0 * M * N * K
        print(f"{name:28s} M={M:6d} N={N:6d} K={K:6d}  mmqg {t*1e6:8.1f} us {fl/t/1e12:7.1f} TF/s | "
              f"cuBLAS {tt*1e6:8.1f} us {fl/tt/1e12:7.1f} TF/s")


if __name__ == "__main__":
    main()

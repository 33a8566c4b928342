"""GEMM time vs K at fixed M,N (fp32 and bf16 output): separates epilogue cost from main-loop cost."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
from mmqg import ops

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

M, N = 25600, 2048
for K in (64, 128, 256, 512, 1024, 2048):
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    o32 = torch.empty(M, N, device="cuda"); o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t32 = timeit(lambda: ops.gemm_bf16(A, B, out=o32)); t16 = timeit(lambda: ops.gemm_bf16(A, B, out=o16))
    print(f"K={K:5d}  fp32 out {t32:7.1f} us   bf16 out {t16:7.1f} us")
src = torch.empty(M, N, device="cuda"); dst = torch.empty(M, N, device="cuda")
print("copy 210 MB fp32 (read+write):", timeit(lambda: dst.copy_(src)), "us;  fill:", timeit(lambda: dst.fill_(1.0)), "us")

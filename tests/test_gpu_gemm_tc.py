"""tcgen05/TMA bf16 GEMM vs PyTorch on the same bf16-rounded operands (GPU only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mmqg import ops as o
    return o


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


SHAPES = [(128, 128, 64), (128, 128, 512), (256, 2048, 512), (5120, 2048, 304), (300, 200, 72), (2048, 512, 2560),
          (1000, 10000, 512), (64, 488, 512)]


@pytest.mark.parametrize("amn,bmn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bf16(ops, M, N, K, amn, bmn):
    if amn and M % 8 or bmn and N % 8:
        pytest.skip("MN-major leading dimension must be a multiple of 8")
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((K, M) if amn else (M, K), device="cuda", generator=g).bfloat16()
    B = torch.randn((K, N) if bmn else (N, K), device="cuda", generator=g).bfloat16()
    ref = (A.t() if amn else A).double() @ (B if bmn else B.t()).double()
    out = ops.gemm_bf16(A, B, amn, bmn)
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-5, rel(out, ref)
    outb = ops.gemm_bf16(A, B, amn, bmn, out_dtype=torch.bfloat16)
    assert rel(outb, ref) < 5e-3
    # epilogue options + second operand pair + split-K
    K2 = 128
    A2 = torch.randn((K2, M) if amn else (M, K2), device="cuda", generator=g).bfloat16()
    B2 = torch.randn((K2, N) if bmn else (N, K2), device="cuda", generator=g).bfloat16()
    Cin = torch.randn(M, N, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    ref2 = 0.5 * (ref + (A2.t() if amn else A2).double() @ (B2 if bmn else B2.t()).double()) + 2 * Cin.double() + bias.double()
    out2 = ops.gemm_bf16(A, B, amn, bmn, A2=A2, B2=B2, Cin=Cin, beta=2.0, bias=bias, alpha=0.5)
    assert rel(out2, ref2) < 1e-5
    if K >= 128:
        parts = ops.gemm_bf16(A, B, amn, bmn, split_k=2)
        assert rel(parts.sum(0), ref) < 1e-5


def test_gemm_bf16_strided(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(512, 1152 + 304, device="cuda", generator=g).bfloat16()
    W = torch.randn(2048, 1152, device="cuda", generator=g).bfloat16()
    out = torch.zeros(512, 4096, device="cuda")
    ops.gemm_bf16(X[:, 304:], W, out=out[:, 1024:3072])
    assert rel(out[:, 1024:3072], X[:, 304:].double() @ W.double().t()) < 1e-5
    assert float(out[:, :1024].abs().max()) == 0 and float(out[:, 3072:].abs().max()) == 0

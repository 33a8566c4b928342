// FP32 contraction on the tcgen05 tensor cores by bf16 splitting (the parity mode's tensor-core twin of gemm_f32.cu).
//
//   x = hi + lo,  hi = bf16(x),  lo = bf16(x - hi)          (|x - hi - lo| <= 2^-17 |x|)
//   A B  ~=  hi(A) lo(B) + lo(A) hi(B) + hi(A) hi(B)        (the dropped lo*lo term is <= 2^-16 relative)
//
// with fp32 accumulation in tensor memory, small terms first.  Each operand is written once as a bf16 matrix
// twice as long in K — A as [hi | lo], B as [lo | hi] — so that the three products are TWO operand pairs of
// ONE gemm_bf16 launch: pair 1 = the whole 2K (hi*lo + lo*hi), pair 2 = A's first half against B's second
// half (hi*hi).  A second operand pair of the fp32 call (x W_ih^T + h W_hh^T, gemm_f32 semantics) is
// concatenated along K before splitting.  Same argument struct, epilogue (alpha, beta, Cin, bias) and
// split-K convention as gemm_f32 (reference call sites: SURVEY.md section 2.3; decoder.py:78,84,92,106,
// encoder.py:69,98 and their autograd twins).
//
// Constant B operands (parameters, the packed attention Linear) are split once per C-ABI call and looked up
// by (pointer, view) afterwards: the per-timestep products then cost one split of the (B x K) activation and
// one tensor-core launch.
#include <cuda_bf16.h>
#include <vector>
#include "kernels.h"

namespace mmqg {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void split2(float x, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// K contiguous: src (rows, K1) [| src2 (rows, K2)] -> dst (rows, 2*Kc), Kc % 8 == 0; columns K1+K2 .. Kc are zero.
// lo_first = 0: [hi | lo];  1: [lo | hi].  One thread per column pair.
__global__ void split_kc_kernel(const float* __restrict__ s1, long long ld1, int K1, const float* __restrict__ s2, long long ld2,
                                int K2, bf16* __restrict__ dst, int Kc, long long rows, int lo_first) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = Kc >> 1;
  const long long total = rows * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / half;
    const int c = 2 * (int)(i - r * half);
    float x[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int cc = c + j;
      x[j] = cc < K1 ? s1[r * ld1 + cc] : (cc - K1 < K2 ? s2[r * ld2 + (cc - K1)] : 0.f);
    }
    bf16 h0, l0, h1, l1;
    split2(x[0], h0, l0);
    split2(x[1], h1, l1);
    __nv_bfloat162 hh, ll;
    hh.x = h0; hh.y = h1; ll.x = l0; ll.y = l1;
    bf16* row = dst + r * (2ll * Kc);
    *reinterpret_cast<__nv_bfloat162*>(row + (lo_first ? Kc : 0) + c) = hh;
    *reinterpret_cast<__nv_bfloat162*>(row + (lo_first ? 0 : Kc) + c) = ll;
  }
}

// K along the rows: src (K1, n) [; src2 (K2, n)] -> dst (2*Kc, ldd), Kc % 8 == 0 (rows K1+K2 .. Kc are zero), ldd % 8 == 0 >= n.
__global__ void split_mn_kernel(const float* __restrict__ s1, long long ld1, int K1, const float* __restrict__ s2, long long ld2,
                                int K2, bf16* __restrict__ dst, long long ldd, int n, int Kc, int lo_first) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = (n + 1) >> 1;
  const long long total = (long long)Kc * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i / half);
    const int c = 2 * (int)(i - (long long)k * half);
    const float* src = k < K1 ? s1 + (long long)k * ld1 : (k - K1 < K2 ? s2 + (long long)(k - K1) * ld2 : nullptr);
    const float x0 = src ? src[c] : 0.f, x1 = (src && c + 1 < n) ? src[c + 1] : 0.f;
    bf16 h0, l0, h1, l1;
    split2(x0, h0, l0);
    split2(x1, h1, l1);
    __nv_bfloat162 hh, ll;
    hh.x = h0; hh.y = h1; ll.x = l0; ll.y = l1;
    *reinterpret_cast<__nv_bfloat162*>(dst + ((long long)(lo_first ? Kc : 0) + k) * ldd + c) = hh;
    *reinterpret_cast<__nv_bfloat162*>(dst + ((long long)(lo_first ? 0 : Kc) + k) * ldd + c) = ll;
  }
}

struct X3Ent { const float* p; const float* p2; int ld, ld2, K, K2, n, mn; bf16* out; };
struct X3Ctx {
  char* arena = nullptr; size_t arena_bytes = 0, arena_off = 0;      // constant B operands of the current call
  bf16* sa = nullptr; size_t sa_elems = 0;                            // A operand of the launch being issued
  bf16* sb = nullptr; size_t sb_elems = 0;                            // non-constant B operand
  std::vector<X3Ent> ents;
};
static thread_local X3Ctx tl_x3;
static thread_local const void* tl_presplit_a = nullptr;
void f32x3_presplit_a(const void* a_split) { tl_presplit_a = a_split; }

void f32x3_bind(void* arena, size_t arena_bytes, void* sa, size_t sa_bytes, void* sb, size_t sb_bytes) {
  X3Ctx& c = tl_x3;
  c.arena = reinterpret_cast<char*>(arena); c.arena_bytes = arena_bytes; c.arena_off = 0;
  c.sa = reinterpret_cast<bf16*>(sa); c.sa_elems = sa_bytes / 2;
  c.sb = reinterpret_cast<bf16*>(sb); c.sb_elems = sb_bytes / 2;
  c.ents.clear();
}

// elements of the split copy of an operand with `n` rows/columns on the M|N side and reduction length K (+K2)
static inline size_t split_elems(bool mn, int n, int K, int K2, int* Kc_out, long long* ld_out) {
  const int Kc = (K + K2 + 7) / 8 * 8;
  *Kc_out = Kc;
  if (mn) {
    const long long ld = (n + 7) / 8 * 8;
    *ld_out = ld;
    return (size_t)2 * Kc * ld;
  }
  *ld_out = 2ll * Kc;
  return (size_t)n * 2 * Kc;
}

size_t f32x3_split_bytes(int mn, long long n, long long K) {      // upper bound used to size the workspace regions
  return (size_t)(2 * ((K + 7) / 8 * 8) * (mn ? (n + 7) / 8 * 8 : n)) * 2;
}

static int run_split(bool mn, const float* p, int ld, int K, const float* p2, int ld2, int K2, int n, bf16* out, int Kc,
                     long long ldd, int lo_first, cudaStream_t st) {
  MMQG_PROBE(KC_OTHER, 0, 8.0 * n * (K + K2));
  if (mn) {
    const long long total = (long long)Kc * ((n + 1) / 2);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    MMQG_CUDA(launch_k(split_mn_kernel, dim3(blocks), dim3(256), 0, st, p, (long long)ld, K, p2, (long long)ld2, K2, out, ldd, n, Kc, lo_first));
  } else {
    const long long total = (long long)n * (Kc / 2);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    MMQG_CUDA(launch_k(split_kc_kernel, dim3(blocks), dim3(256), 0, st, p, (long long)ld, K, p2, (long long)ld2, K2, out, Kc, (long long)n, lo_first));
  }
  MMQG_LAUNCH_CHECK();
  return 0;
}

int gemm_f32x3(const mmqg_gemm_args& a, bool b_const, cudaStream_t st) {
  X3Ctx& c = tl_x3;
  MMQG_REQUIRE(c.sa && c.sb && c.arena, "gemm_f32x3: no split buffers bound");
  MMQG_REQUIRE(a.A && a.B && a.C && a.M > 0 && a.N > 0 && a.K > 0, "gemm_f32x3: bad args");
  MMQG_REQUIRE(a.K2 == 0 || (a.A2 && a.B2), "gemm_f32x3: K2>0 needs A2,B2");
  const bool amn = a.transA != 0, bmn = a.transB == 0;
  int KcA, KcB;
  long long ldA, ldB;
  const size_t ea = split_elems(amn, a.M, a.K, a.K2, &KcA, &ldA);
  const size_t eb = split_elems(bmn, a.N, a.K, a.K2, &KcB, &ldB);
  const bf16* A_hi = c.sa;
  if (tl_presplit_a) {      // the producer kernel wrote [hi | lo] itself
    A_hi = reinterpret_cast<const bf16*>(tl_presplit_a);
    tl_presplit_a = nullptr;
    MMQG_REQUIRE(!amn && a.K2 == 0 && a.K % 8 == 0, "gemm_f32x3: pre-split A needs a K-contiguous single operand with K %% 8 == 0");
  } else {
    MMQG_REQUIRE(ea <= c.sa_elems, "gemm_f32x3: A split needs %zu elements, scratch holds %zu", ea, c.sa_elems);
    MMQG_TRY(run_split(amn, a.A, a.lda, a.K, a.A2, a.lda2, a.K2, a.M, c.sa, KcA, ldA, 0, st));
  }
  bf16* sbp = nullptr;
  if (b_const) {
    for (const X3Ent& e : c.ents)
      if (e.p == a.B && e.p2 == (a.K2 ? a.B2 : nullptr) && e.ld == a.ldb && e.ld2 == (a.K2 ? a.ldb2 : 0) && e.K == a.K && e.K2 == a.K2 &&
          e.n == a.N && e.mn == (int)bmn) { sbp = e.out; break; }
    if (!sbp) {
      const size_t off = align_up(c.arena_off, 256);
      MMQG_REQUIRE(off + eb * 2 <= c.arena_bytes, "gemm_f32x3: constant-operand arena exhausted (%zu + %zu > %zu)", off, eb * 2, c.arena_bytes);
      sbp = reinterpret_cast<bf16*>(c.arena + off);
      c.arena_off = off + eb * 2;
      MMQG_TRY(run_split(bmn, a.B, a.ldb, a.K, a.B2, a.ldb2, a.K2, a.N, sbp, KcB, ldB, 1, st));
      c.ents.push_back(X3Ent{a.B, a.K2 ? a.B2 : nullptr, a.ldb, a.K2 ? a.ldb2 : 0, a.K, a.K2, a.N, (int)bmn, sbp});
    }
  } else {
    MMQG_REQUIRE(eb <= c.sb_elems, "gemm_f32x3: B split needs %zu elements, scratch holds %zu", eb, c.sb_elems);
    sbp = c.sb;
    MMQG_TRY(run_split(bmn, a.B, a.ldb, a.K, a.B2, a.ldb2, a.K2, a.N, sbp, KcB, ldB, 1, st));
  }
  mmqg_gemm_bf16_args g{};
  g.a_mn_major = amn; g.b_mn_major = bmn;
  g.lda = (int)ldA; g.ldb = (int)ldB; g.lda2 = (int)ldA; g.ldb2 = (int)ldB;
  g.C = a.C; g.ldc = a.ldc; g.c_bf16 = 0;
  g.Cin = a.Cin; g.ldcin = a.ldcin; g.bias = a.bias; g.M = a.M; g.N = a.N; g.alpha = a.alpha; g.beta = a.beta;
  g.split_k = a.split_k; g.c_split_stride = a.c_split_stride;
  const bf16* B_lo = sbp;
  const bf16* B_hi = bmn ? sbp + (long long)KcB * ldB : sbp + KcB;
  g.A = A_hi; g.B = B_lo; g.K = 2 * KcA;            // hi*lo + lo*hi
  g.A2 = A_hi; g.B2 = B_hi; g.K2 = KcA;             // hi*hi
  return gemm_bf16(g, st);
}

}  // namespace mmqg

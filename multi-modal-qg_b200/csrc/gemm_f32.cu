// FP32 SIMT GEMM: the parity-mode contraction (fp32 FMA accumulation, no tensor cores).
//
//   C(M,N) = alpha * [ op(A) op(B) + op(A2) op(B2) ] + beta * Cin + bias
//
// Replaces the addmm / mm calls the reference issues through torch.nn.Linear and inside
// aten::lstm (decoder.py:78,84,92,106; encoder.py:69,98; and their
// autograd twins).  Two operand pairs share one accumulator so that x W_ih^T + h W_hh^T is
// a single launch.  split_k > 1 writes per-slice partial sums that the consumer adds
// (lstm_pointwise_bwd does), which is how the skinny per-timestep products fill 148 SMs.
#include "common.cuh"

namespace mmqg {

struct GemmP {
  const float* A[2]; const float* B[2]; int lda[2], ldb[2], K[2];
  float* C; int ldc; const float* Cin; int ldcin; const float* bias;
  int M, N; float alpha, beta; int split_k; long long c_split_stride;
};

// Tile of an operand whose storage is (row, k) with k contiguous -> smem S[k][row].
template <int ROWS, int BK, int LDS_, bool VEC>
struct LoadKContig {
  static constexpr int NV = VEC ? (ROWS * BK / 4 / 256) : (ROWS * BK / 256);
  float4 v4[VEC ? NV : 1];
  float v1[VEC ? 1 : NV];
  __device__ __forceinline__ void load(const float* p, int ld, int row0, int k0, int rows, int K, int tid) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        int r = idx / (BK / 4), kq = idx % (BK / 4);
        int gr = row0 + r, gk = k0 + 4 * kq;
        v4[i] = (gr < rows && gk < K) ? *reinterpret_cast<const float4*>(p + (size_t)gr * ld + gk)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        int r = idx / BK, k = idx % BK;
        int gr = row0 + r, gk = k0 + k;
        v1[i] = (gr < rows && gk < K) ? p[(size_t)gr * ld + gk] : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(float (*S)[LDS_], int tid) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        int r = idx / (BK / 4), kq = idx % (BK / 4);
        S[4 * kq + 0][r] = v4[i].x; S[4 * kq + 1][r] = v4[i].y;
        S[4 * kq + 2][r] = v4[i].z; S[4 * kq + 3][r] = v4[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        S[idx % BK][idx / BK] = v1[i];
      }
    }
  }
};

// Tile of an operand whose storage is (k, row) with row contiguous -> smem S[k][row].
template <int ROWS, int BK, int LDS_, bool VEC>
struct LoadRowContig {
  static constexpr int NV = VEC ? (ROWS * BK / 4 / 256) : (ROWS * BK / 256);
  float4 v4[VEC ? NV : 1];
  float v1[VEC ? 1 : NV];
  __device__ __forceinline__ void load(const float* p, int ld, int row0, int k0, int rows, int K, int tid) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        int k = idx / (ROWS / 4), rq = idx % (ROWS / 4);
        int gk = k0 + k, gr = row0 + 4 * rq;
        v4[i] = (gk < K && gr < rows) ? *reinterpret_cast<const float4*>(p + (size_t)gk * ld + gr)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        int k = idx / ROWS, r = idx % ROWS;
        int gk = k0 + k, gr = row0 + r;
        v1[i] = (gk < K && gr < rows) ? p[(size_t)gk * ld + gr] : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(float (*S)[LDS_], int tid) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        int k = idx / (ROWS / 4), rq = idx % (ROWS / 4);
        *reinterpret_cast<float4*>(&S[k][4 * rq]) = v4[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int idx = tid + i * 256;
        S[idx / ROWS][idx % ROWS] = v1[i];
      }
    }
  }
};

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmP p) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  constexpr int RM = TM / 4, RN = TN / 4;
  constexpr int LA = BM + 4, LB = BN + 4;
  __shared__ __align__(16) float As[BK][LA];
  __shared__ __align__(16) float Bs[BK][LB];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  const int nt0 = (p.K[0] + BK - 1) / BK;
  const int nt1 = p.K[1] > 0 ? (p.K[1] + BK - 1) / BK : 0;
  const int nt = nt0 + nt1;
  const int per = (nt + p.split_k - 1) / p.split_k;
  const int t_begin = blockIdx.z * per;
  const int t_end = min(nt, t_begin + per);

  // A is (M,K) k-contiguous unless TA; B is (K,N) n-contiguous unless TB (then (N,K)).
  typename std::conditional<TA, LoadRowContig<BM, BK, LA, VEC>, LoadKContig<BM, BK, LA, VEC>>::type la;
  typename std::conditional<TB, LoadKContig<BN, BK, LB, VEC>, LoadRowContig<BN, BK, LB, VEC>>::type lb;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto fetch = [&](int t) {
    int s = t >= nt0 ? 1 : 0;
    int k0 = (s ? t - nt0 : t) * BK;
    la.load(p.A[s], p.lda[s], m0, k0, p.M, p.K[s], tid);
    lb.load(p.B[s], p.ldb[s], n0, k0, p.N, p.K[s], tid);
  };

  if (t_begin < t_end) fetch(t_begin);
  for (int t = t_begin; t < t_end; ++t) {
    la.store(As, tid);
    lb.store(Bs, tid);
    __syncthreads();
    if (t + 1 < t_end) fetch(t + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < RM; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&As[k][g * (BM / RM) + ty * 4]);
        a[4 * g] = v.x; a[4 * g + 1] = v.y; a[4 * g + 2] = v.z; a[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < RN; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[k][g * (BN / RN) + tx * 4]);
        b[4 * g] = v.x; b[4 * g + 1] = v.y; b[4 * g + 2] = v.z; b[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* C = p.C + (size_t)blockIdx.z * p.c_split_stride;
  const bool lead = blockIdx.z == 0;
#pragma unroll
  for (int gi = 0; gi < RM; ++gi)
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      int m = m0 + gi * (BM / RM) + ty * 4 + ii;
      if (m >= p.M) continue;
#pragma unroll
      for (int gj = 0; gj < RN; ++gj) {
        int n = n0 + gj * (BN / RN) + tx * 4;
        float v[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float x = p.alpha * acc[gi * 4 + ii][gj * 4 + jj];
          if (lead && n + jj < p.N) {
            if (p.Cin) x += p.beta * p.Cin[(size_t)m * p.ldcin + n + jj];
            if (p.bias) x += p.bias[n + jj];
          }
          v[jj] = x;
        }
        float* dst = C + (size_t)m * p.ldc + n;
        if (VEC && n + 3 < p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            if (n + jj < p.N) dst[jj] = v[jj];
        }
      }
    }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_cfg(const GemmP& p, bool ta, bool tb, bool vec, cudaStream_t st) {
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), p.split_k);
#define MMQG_GEMM_CASE(TA_, TB_, V_)                                                     \
  if (ta == TA_ && tb == TB_ && vec == V_) {                                             \
    gemm_f32_kernel<BM, BN, BK, TM, TN, TA_, TB_, V_><<<grid, 256, 0, st>>>(p);          \
    MMQG_LAUNCH_CHECK();                                                                 \
    return 0;                                                                            \
  }
  MMQG_GEMM_CASE(false, false, false) MMQG_GEMM_CASE(false, false, true)
  MMQG_GEMM_CASE(false, true, false)  MMQG_GEMM_CASE(false, true, true)
  MMQG_GEMM_CASE(true, false, false)  MMQG_GEMM_CASE(true, false, true)
  MMQG_GEMM_CASE(true, true, false)   MMQG_GEMM_CASE(true, true, true)
#undef MMQG_GEMM_CASE
  return set_err(MMQG_ERR_BAD_ARG, "gemm: unreachable");
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int gemm_f32(const mmqg_gemm_args& a, cudaStream_t st) {
  MMQG_REQUIRE(a.A && a.B && a.C, "gemm: null operand");
  MMQG_REQUIRE(a.M > 0 && a.N > 0 && a.K >= 0, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  MMQG_REQUIRE(a.K2 == 0 || (a.A2 && a.B2), "gemm: K2>0 needs A2,B2");
  GemmP p;
  p.A[0] = a.A; p.B[0] = a.B; p.lda[0] = a.lda; p.ldb[0] = a.ldb; p.K[0] = a.K;
  p.A[1] = a.K2 > 0 ? a.A2 : a.A; p.B[1] = a.K2 > 0 ? a.B2 : a.B;
  p.lda[1] = a.K2 > 0 ? a.lda2 : a.lda; p.ldb[1] = a.K2 > 0 ? a.ldb2 : a.ldb; p.K[1] = a.K2 > 0 ? a.K2 : 0;
  p.C = a.C; p.ldc = a.ldc; p.Cin = a.Cin; p.ldcin = a.ldcin; p.bias = a.bias;
  p.M = a.M; p.N = a.N; p.alpha = a.alpha; p.beta = a.beta;
  p.split_k = a.split_k > 1 ? a.split_k : 1; p.c_split_stride = a.c_split_stride;
  const bool ta = a.transA != 0, tb = a.transB != 0;
  // Vector path: every contiguous extent a float4 can straddle must be a multiple of 4
  // and every base pointer / leading dimension 16-byte aligned.
  bool vec = true;
  for (int s = 0; s < (p.K[1] > 0 ? 2 : 1); ++s) {
    vec = vec && al16(p.A[s]) && al16(p.B[s]) && p.lda[s] % 4 == 0 && p.ldb[s] % 4 == 0;
    if (!ta) vec = vec && p.K[s] % 4 == 0; else vec = vec && p.M % 4 == 0;
    if (tb) vec = vec && p.K[s] % 4 == 0; else vec = vec && p.N % 4 == 0;
  }
  MMQG_PROBE(tl_gemm_class, 2.0 * p.M * p.N * ((double)p.K[0] + p.K[1]),
             4.0 * ((double)p.M * (p.K[0] + p.K[1]) + (double)p.N * (p.K[0] + p.K[1]) + (double)p.M * p.N));
  const long long big_tiles = (long long)ceil_div(p.M, 128) * ceil_div(p.N, 128);
  if (big_tiles >= 120 && p.split_k == 1) return launch_cfg<128, 128, 8, 8, 8>(p, ta, tb, vec, st);
  return launch_cfg<64, 64, 16, 4, 4>(p, ta, tb, vec, st);
}

}  // namespace mmqg

extern "C" int mmqg_gemm_f32(const mmqg_gemm_args* a, void* stream) {
  if (!a) return mmqg::set_err(MMQG_ERR_BAD_ARG, "gemm: null args");
  return mmqg::gemm_f32(*a, mmqg::as_stream(stream));
}

// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory
// ring -> tcgen05.mma (kind::f16, fp32 accumulators in TMEM) -> tcgen05.ld epilogue.
//
//   C(M,N) = alpha * [ op(A) op(B) + op(A2) op(B2) ] + beta * Cin + bias
//
// Operands are bf16 in global memory in either major-ness, consumed without transposition:
//   A K-major : stored (M,K) row-major      A MN-major : stored (K,M) row-major
//   B K-major : stored (N,K) row-major      B MN-major : stored (K,N) row-major
// which covers the three contractions of the path: X W^T (both K-major), dG W (B MN-major) and
// dG^T X (both MN-major).  Replaces the addmm/mm calls of the reference (SURVEY.md section
// 2.3) in bf16 mode; the fp32 SIMT twin is gemm_f32.cu.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2-9 epilogue
// (TMEM lane quarter = warp % 4, two warps per quarter split the columns).  Tile 128 x BN x 64, 6-stage (BN=128) ring, one tile per CTA;
// split_k > 1 writes fp32 partial tiles that the caller reduces.
#include "kernels.h"
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace mmqg {

using namespace tc;

static constexpr int TBM = 128, TBK = 64;


template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, TcGemmP p) {
  constexpr int A_BYTES = TBM * TBK * 2, B_BYTES = BN * TBK * 2;
  constexpr int STAGES = (BN == 128) ? 6 : (BN == 64 ? 8 : 4);
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; the launch reserves 1 KB of slack for this.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  __shared__ uint64_t full[STAGES], empty[STAGES], accum_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
  const int nk = p.nk1 + p.nk2;
  const int per = (nk + p.split_k - 1) / p.split_k;
  const int kb_begin = blockIdx.z * per;
  const int kb_end = min(nk, kb_begin + per);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&accum_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (p.nk2 > 0) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
  }
  if (warp == 1) tmem_alloc(&tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  // PDL: everything above overlapped the previous kernel's tail; from here on we read / write
  // memory the predecessor may still be producing
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = kb_begin, i = 0; kb < kb_end; ++kb, ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        const bool second = kb >= p.nk1;
        const int k0 = (second ? kb - p.nk1 : kb) * TBK;
        const CUtensorMap* ma = second ? &tmA2 : &tmA;
        const CUtensorMap* mb = second ? &tmB2 : &tmB;
        uint8_t* a = sA + s * A_BYTES;
        uint8_t* b = sB + s * B_BYTES;
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < TBM / 64; ++j) tma_load_2d(a + j * 8192, ma, &full[s], m0 + 64 * j, k0);
        } else {
          tma_load_2d(a, ma, &full[s], k0, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b + j * 8192, mb, &full[s], n0 + 64 * j, k0);
        } else {
          tma_load_2d(b, mb, &full[s], k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(TBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = kb_begin, i = 0; kb < kb_end; ++kb, ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(sA + s * A_BYTES), b_addr = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < TBK / 16; ++k) {
          const uint64_t ad = A_MN ? umma_smem_desc(a_addr + k * 2048, 8192, 1024) : umma_smem_desc(a_addr + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? umma_smem_desc(b_addr + k * 2048, 8192, 1024) : umma_smem_desc(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);      // frees the smem slot once these MMAs have read it
      }
      umma_commit(&accum_full);      // accumulator complete
    }
  } else {
    // ---- epilogue (8 warps: 2 per TMEM lane quarter, each half of the tile's columns) ----
    // The operand ring is idle once accum_full fires, so it doubles as the staging buffer.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const EpiOut out = make_epi_out(p, blockIdx.z, blockIdx.z == 0);
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * EPI_STG_FLOATS;
    mbar_wait(&accum_full, 0);
    tc_fence_after_sync();
    constexpr int NCH = BN / 64;
    epilogue_block<NCH>(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + half * (BN / 2), stg, out, m0 + 32 * q,
                        n0 + half * (BN / 2), lane, nullptr);
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// ---- host ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(MMQG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  MMQG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p is not 16-byte aligned", base);
  MMQG_REQUIRE(ld % 8 == 0, "tensor map: leading dimension %llu is not a multiple of 8 bf16", (unsigned long long)ld);
  MMQG_REQUIRE(box_cols * 2 == 128 && box_rows <= 256, "tensor map: box %ux%u", box_rows, box_cols);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(MMQG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
                   (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2,
                     const TcGemmP& p, cudaStream_t st) {
  constexpr int STAGES = (BN == 128) ? 6 : (BN == 64 ? 8 : 4);
  constexpr int SMEM = STAGES * (TBM * TBK * 2 + BN * TBK * 2) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, TBM), p.split_k);
  MMQG_CUDA(launch_k(gemm_tc_kernel<BN, A_MN, B_MN>, grid, dim3(320), SMEM, st, a, b, a2, b2, p));
  MMQG_LAUNCH_CHECK();
  return 0;
}

int gemm_bf16(const mmqg_gemm_bf16_args& g, cudaStream_t st) {
  MMQG_REQUIRE(g.A && g.B && g.C && g.M > 0 && g.N > 0 && g.K > 0, "gemm_bf16: bad args");
  MMQG_REQUIRE(g.K2 == 0 || (g.A2 && g.B2), "gemm_bf16: K2>0 needs A2,B2");
  const bool amn = g.a_mn_major != 0, bmn = g.b_mn_major != 0;
  // Tile width: skinny problems (the per-timestep products, M = B) take 64-wide tiles so that
  // twice as many SMs pull operands; everything else 128.
  const int want_split = g.split_k > 1 ? g.split_k : 1;
  const bool narrow = (long long)ceil_div(g.M, TBM) * ceil_div(g.N, 128) * want_split < 96 && g.N > 64;
  // Large problems go to the persistent 128x256-tile kernel (double-buffered accumulators).
  static const bool persist_on = []() { const char* e = getenv("MMQG_GEMM_PERSIST"); return !(e && e[0] == '0'); }();
  const bool persist = persist_on && want_split == 1 && (long long)ceil_div(g.M, TBM) * ceil_div(g.N, 256) >= 120;
  const int BN = persist ? 256 : (narrow ? 64 : 128);
  CUtensorMap ta, tb, ta2, tb2;
  auto mk_a = [&](CUtensorMap* t, const void* A, int lda, int K) {
    return amn ? make_tmap_bf16_2d(t, A, K, g.M, lda, 64, 64) : make_tmap_bf16_2d(t, A, g.M, K, lda, TBM, 64);
  };
  auto mk_b = [&](CUtensorMap* t, const void* Bp, int ldb, int K) {
    return bmn ? make_tmap_bf16_2d(t, Bp, K, g.N, ldb, 64, 64) : make_tmap_bf16_2d(t, Bp, g.N, K, ldb, BN, 64);
  };
  MMQG_TRY(mk_a(&ta, g.A, g.lda, g.K));
  MMQG_TRY(mk_b(&tb, g.B, g.ldb, g.K));
  if (g.K2 > 0) {
    MMQG_TRY(mk_a(&ta2, g.A2, g.lda2, g.K2));
    MMQG_TRY(mk_b(&tb2, g.B2, g.ldb2, g.K2));
  } else {
    ta2 = ta; tb2 = tb;
  }
  TcGemmP p{};
  p.M = g.M; p.N = g.N; p.nk1 = ceil_div(g.K, TBK); p.nk2 = g.K2 > 0 ? ceil_div(g.K2, TBK) : 0;
  p.C = g.C; p.ldc = g.ldc; p.c_bf16 = g.c_bf16;
  p.Cin = g.Cin; p.ldcin = g.ldcin; p.beta = g.beta; p.alpha = g.alpha; p.bias = g.bias;
  p.split_k = g.split_k > 1 ? g.split_k : 1; p.c_split_stride = g.c_split_stride;
  MMQG_REQUIRE(p.split_k == 1 || !g.c_bf16, "gemm_bf16: split-K partials are fp32");
  const int nk_all = p.nk1 + p.nk2, per_slice = ceil_div(nk_all, p.split_k);
  if (ceil_div(nk_all, per_slice) < p.split_k) {
    // the k-blocks do not fill every requested slice (fewer blocks than slices, or the rounded-up slice
    // length leaves trailing slices empty): the surplus partial tiles are defined as zero
    const int used = ceil_div(nk_all, per_slice);
    MMQG_REQUIRE(g.ldc == g.N, "gemm_bf16: clamped split-K needs contiguous partials");
    MMQG_CUDA(cudaMemsetAsync(reinterpret_cast<float*>(g.C) + (size_t)used * g.c_split_stride, 0,
                              sizeof(float) * (size_t)(p.split_k - used) * g.c_split_stride, st));
    p.split_k = used;
  }
  MMQG_PROBE(tl_gemm_class, 2.0 * g.M * g.N * ((double)g.K + g.K2),
             2.0 * ((double)g.M + g.N) * ((double)g.K + g.K2) + (g.c_bf16 ? 2.0 : 4.0) * g.M * g.N);
  if (persist) return gemm_tc_persist_launch(ta, tb, ta2, tb2, p, amn, bmn, st);
  if (narrow) {
    if (!amn && !bmn) return launch_tc<64, false, false>(ta, tb, ta2, tb2, p, st);
    if (!amn && bmn) return launch_tc<64, false, true>(ta, tb, ta2, tb2, p, st);
    if (amn && bmn) return launch_tc<64, true, true>(ta, tb, ta2, tb2, p, st);
    return launch_tc<64, true, false>(ta, tb, ta2, tb2, p, st);
  }
  if (!amn && !bmn) return launch_tc<128, false, false>(ta, tb, ta2, tb2, p, st);
  if (!amn && bmn) return launch_tc<128, false, true>(ta, tb, ta2, tb2, p, st);
  if (amn && bmn) return launch_tc<128, true, true>(ta, tb, ta2, tb2, p, st);
  return launch_tc<128, true, false>(ta, tb, ta2, tb2, p, st);
}

}  // namespace mmqg

extern "C" int mmqg_gemm_bf16(const mmqg_gemm_bf16_args* a, void* stream) {
  if (!a) return mmqg::set_err(MMQG_ERR_BAD_ARG, "gemm_bf16: null args");
  return mmqg::gemm_bf16(*a, mmqg::as_stream(stream));
}

// Persistent bf16 tensor-core GEMM for the large hoisted products (whole-sequence input
// projections, weight gradients, input gradients, loss-head chunks).
//
// Same contract as gemm_tc.cu (K- or MN-major operands, two operand pairs, alpha/beta/bias,
// fp32 or bf16 output), different schedule:
//   * one CTA per SM loops over 128 x 256 output tiles (tile = blockIdx.x + i * gridDim.x; the
//     tile index runs fastest along N so that CTAs working at the same time share the A rows),
//   * 4-stage TMA ring of 48 KB stages (A 128x64 + B 256x64 bf16, SWIZZLE_128B),
//   * tcgen05.mma 128x256x16 with TWO accumulators in tensor memory (2 x 256 of the 512 columns):
//     the MMA warp starts tile i+1 while the four epilogue warps drain tile i, so the epilogue
//     (TMEM -> registers -> padded smem transpose -> coalesced global stores, optional addend
//     and bias) is off the tensor pipe's critical path,
//   * the epilogue has its own staging buffer because the operand ring is never idle.
// The one-tile-per-CTA kernel in gemm_tc.cu stays for skinny problems (per-timestep products).
#include "kernels.h"
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace mmqg {

using namespace tc;

static constexpr int PBM = 128, PBN = 256, PBK = 64, PSTAGES = 4;
static constexpr int PA_BYTES = PBM * PBK * 2, PB_BYTES = PBN * PBK * 2;
static constexpr int PSTG_FLOATS = 4 * EPI_STG_FLOATS;
static constexpr int PSMEM = PSTAGES * (PA_BYTES + PB_BYTES) + PSTG_FLOATS * 4 + 1024;

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, TcGemmP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + PSTAGES * PA_BYTES;
  float* stg_all = reinterpret_cast<float*>(smem + PSTAGES * (PA_BYTES + PB_BYTES));
  __shared__ uint64_t full[PSTAGES], empty[PSTAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + PBN - 1) / PBN, tiles_m = (p.M + PBM - 1) / PBM;
  const int n_tiles = tiles_n * tiles_m;
  const int nk = p.nk1 + p.nk2;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (p.nk2 > 0) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int i = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * PBM, n0 = (tile % tiles_n) * PBN;
        for (int kb = 0; kb < nk; ++kb, ++i) {
          const int s = i % PSTAGES, ph = (i / PSTAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], PA_BYTES + PB_BYTES);
          const bool second = kb >= p.nk1;
          const int k0 = (second ? kb - p.nk1 : kb) * PBK;
          const CUtensorMap* ma = second ? &tmA2 : &tmA;
          const CUtensorMap* mb = second ? &tmB2 : &tmB;
          uint8_t* a = sA + s * PA_BYTES;
          uint8_t* b = sB + s * PB_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < PBM / 64; ++j) tma_load_2d(a + j * 8192, ma, &full[s], m0 + 64 * j, k0);
          } else {
            tma_load_2d(a, ma, &full[s], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < PBN / 64; ++j) tma_load_2d(b + j * 8192, mb, &full[s], n0 + 64 * j, k0);
          } else {
            tma_load_2d(b, mb, &full[s], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(PBM, PBN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int i = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);          // epilogue drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * PBN;
        for (int kb = 0; kb < nk; ++kb, ++i) {
          const int s = i % PSTAGES, ph = (i / PSTAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * PA_BYTES), b_addr = smem_u32(sB + s * PB_BYTES);
#pragma unroll
          for (int k = 0; k < PBK / 16; ++k) {
            const uint64_t ad = A_MN ? umma_smem_desc(a_addr + k * 2048, 8192, 1024) : umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? umma_smem_desc(b_addr + k * 2048, 8192, 1024) : umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int q = warp & 3;
    float* stg = stg_all + q * EPI_STG_FLOATS;
    const EpiOut out = make_epi_out(p, 0, true);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int m0 = (tile / tiles_n) * PBM, n0 = (tile % tiles_n) * PBN;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after_sync();
      epilogue_block<PBN / 32>(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * PBN, stg, out, m0 + 32 * q, n0, lane,
                               &acc_empty[buf]);
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool A_MN, bool B_MN>
static int launch_persist(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2,
                          const TcGemmP& p, cudaStream_t st) {
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(gemm_tc_persist_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM));
    int dev = 0;
    MMQG_CUDA(cudaGetDevice(&dev));
    MMQG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr = true;
  }
  const int n_tiles = ceil_div(p.M, PBM) * ceil_div(p.N, PBN);
  const int grid = n_tiles < sms ? n_tiles : sms;
  gemm_tc_persist_kernel<A_MN, B_MN><<<grid, 192, PSMEM, st>>>(a, b, a2, b2, p);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int gemm_tc_persist_launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2,
                           const TcGemmP& p, bool a_mn, bool b_mn, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_persist<false, false>(a, b, a2, b2, p, st);
  if (!a_mn && b_mn) return launch_persist<false, true>(a, b, a2, b2, p, st);
  if (a_mn && b_mn) return launch_persist<true, true>(a, b, a2, b2, p, st);
  return launch_persist<true, false>(a, b, a2, b2, p, st);
}

}  // namespace mmqg

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

// ---- fused vocabulary epilogues (decoder.py:106 + train.py:174 / train.py:107-108) --------------------------
// One epilogue warp owns 32 rows of the 128 x 256 logits tile; tcgen05.ld hands every thread ONE row, so the
// row-wise softmax statistics / arg-max / d logits are plain register arithmetic, 32 columns at a time.
//   VE_STATS  : per (tile, row) running max m and s = sum exp(x - m) over the tile's columns, plus the target
//               column's logit; a tiny merge kernel turns the tiles_n partials of a row into lse and the NLL.
//   VE_DLOGITS: recomputes the tile and stores row_scale * (exp(x - lse) - onehot) as bf16 (coalesced through a
//               per-warp shared-memory transpose): the operand of the dH / dW_out products.
//   VE_ARGMAX : per (tile, row) maximum and its lowest column (greedy decode); merged the same way.
// x = acc + bias; columns >= N get bias = -inf, i.e. probability 0.
template <int MODE>
__device__ __forceinline__ void vocab_epilogue(uint32_t taddr, float* stg, const TcGemmP& p, int m0w, int n0, int tile_n, int lane,
                                               uint64_t* release_bar) {
  using namespace tc;
  const int row = m0w + lane;
  const bool rvalid = row < p.M;
  // bias of the tile's 256 columns -> this warp's staging area (broadcast reads below)
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + 32 * i + lane;
    stg[32 * i + lane] = n < p.N ? (p.bias ? p.bias[n] : 0.f) : -INFINITY;
  }
  __syncwarp();
  long long tgt = -1;
  if (MODE != VE_ARGMAX && rvalid && p.targets) {
    tgt = p.targets[row];
    tgt = tgt < 0 ? 0 : (tgt >= p.N ? p.N - 1 : tgt);
  }
  if (MODE == VE_STATS || MODE == VE_ARGMAX) {
    float m = -INFINITY, ssum = 0.f, tl = 0.f;
    int bi = 0x7fffffff;
    bool has_t = false;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      float v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      tmem_ld_wait();
      if (c == 7 && release_bar) {
        tc_fence_before_sync();
        mbar_arrive(release_bar);
      }
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] += stg[32 * c + j];
        cm = fmaxf(cm, v[j]);
      }
      if (MODE == VE_STATS) {
        const float mn = fmaxf(m, cm);
        float acc = 0.f;
        if (mn > -INFINITY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += __expf(v[j] - mn);
          ssum = ssum * __expf(m - mn) + acc;      // m = -inf on the first chunk: exp(-inf) = 0, ssum is 0 anyway
        }
        m = mn;
        const long long offl = tgt - (n0 + 32 * c);
        const int off = (offl >= 0 && offl < 32) ? (int)offl : -1;      // branch-free select keeps v[] in registers
        has_t = has_t || off >= 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) tl = (j == off) ? v[j] : tl;
      } else {
        if (cm > m) {        // strict: an equal maximum in a later chunk keeps the earlier (lower) column
          m = cm;
          int k = 31;
#pragma unroll
          for (int j = 0; j < 32; ++j) k = min(k, v[j] == cm ? j : 31);
          bi = n0 + 32 * c + k;
        }
      }
    }
    if (rvalid) {
      const size_t o = (size_t)tile_n * p.M + row;
      p.stat_a[o] = m;
      if (MODE == VE_STATS) {
        p.stat_b[o] = ssum;
        if (has_t) p.tgt_logit[row] = tl;
      } else {
        p.stat_i[o] = bi;
      }
    }
  } else {      // VE_DLOGITS
    const float lse = rvalid ? p.lse[row] : 0.f;
    const float rs = rvalid ? p.row_scale[row] : 0.f;
    __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C);
    uint32_t* stw = reinterpret_cast<uint32_t*>(stg + 256);        // 32 rows x 20 words behind the bias tile
    const int rows = min(32, p.M - m0w);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      float v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      tmem_ld_wait();
      if (c == 7 && release_bar) {
        tc_fence_before_sync();
        mbar_arrive(release_bar);
      }
      const long long off = tgt - (n0 + 32 * c);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pr = __expf(v[j] + stg[32 * c + j] - lse);
        v[j] = rs * (pr - ((long long)j == off ? 1.f : 0.f));
      }
      __syncwarp();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * q4 + 2 * e], v[8 * q4 + 2 * e + 1]);
          w[e] = *reinterpret_cast<uint32_t*>(&t2);
        }
        *reinterpret_cast<uint4*>(stw + lane * 20 + 4 * q4) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      __syncwarp();
      const int nb = n0 + 32 * c + 8 * (lane & 3);       // first of this lane's 8 columns
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + (lane >> 2);
        const uint4 x = *reinterpret_cast<const uint4*>(stw + r * 20 + 4 * (lane & 3));
        if (r < rows && nb < p.ldc) *reinterpret_cast<uint4*>(C + (size_t)(m0w + r) * p.ldc + nb) = x;
      }
    }
  }
}

template <bool A_MN, bool B_MN, int EPI = VE_NONE>
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
      if (EPI == VE_NONE)
        epilogue_block<PBN / 32>(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * PBN, stg, out, m0 + 32 * q, n0, lane,
                                 &acc_empty[buf]);
      else
        vocab_epilogue<EPI>(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * PBN, stg, p, m0 + 32 * q, n0, tile % tiles_n, lane,
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

template <int EPI>
static int launch_vocab(const CUtensorMap& a, const CUtensorMap& b, const TcGemmP& p, cudaStream_t st) {
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(gemm_tc_persist_kernel<false, false, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM));
    int dev = 0;
    MMQG_CUDA(cudaGetDevice(&dev));
    MMQG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr = true;
  }
  // MMQG_LH_CTAS caps the CTAs of the loss-head products: they share the GPU with the latency-bound decoder loops,
  // whose kernels cannot co-reside with a ~200 KB-shared-memory CTA on the same SM
  static const int cap = []() { const char* e = getenv("MMQG_LH_CTAS"); return e ? atoi(e) : 0; }();
  const int n_tiles = ceil_div(p.M, PBM) * ceil_div(p.N, PBN);
  int grid = n_tiles < sms ? n_tiles : sms;
  if (cap > 0 && grid > cap) grid = cap;
  gemm_tc_persist_kernel<false, false, EPI><<<grid, 192, PSMEM, st>>>(a, b, a, b, p);
  MMQG_LAUNCH_CHECK();
  return 0;
}

// X (M,K) K-major  x  W (N,K) K-major with one of the fused vocabulary epilogues
int gemm_tc_vocab_launch(const CUtensorMap& a, const CUtensorMap& b, const TcGemmP& p, int mode, cudaStream_t st) {
  if (mode == VE_STATS) return launch_vocab<VE_STATS>(a, b, p, st);
  if (mode == VE_DLOGITS) return launch_vocab<VE_DLOGITS>(a, b, p, st);
  if (mode == VE_ARGMAX) return launch_vocab<VE_ARGMAX>(a, b, p, st);
  return set_err(MMQG_ERR_BAD_ARG, "gemm_tc_vocab_launch: mode %d", mode);
}

int gemm_tc_persist_launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2,
                           const TcGemmP& p, bool a_mn, bool b_mn, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_persist<false, false>(a, b, a2, b2, p, st);
  if (!a_mn && b_mn) return launch_persist<false, true>(a, b, a2, b2, p, st);
  if (a_mn && b_mn) return launch_persist<true, true>(a, b, a2, b2, p, st);
  return launch_persist<true, false>(a, b, a2, b2, p, st);
}

}  // namespace mmqg

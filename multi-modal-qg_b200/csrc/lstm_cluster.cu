// Cluster variant of the persistent recurrent-cell kernels (bf16 mode, H = 32 * cluster size).
//
// Same math as lstm_persist.cu; what changes is how the CTAs that share a batch tile talk:
//   * the H/32 CTAs of one 128-row batch tile form ONE thread-block cluster (16 CTAs at H=512,
//     non-portable size), each owning 32 hidden units (forward: 128 gate rows x H of W_hh,
//     128 KB resident in shared memory; backward: 32 rows x 4H of W_hh^T, 128 KB),
//   * the per-step exchange of h_t / dG_t still goes through global memory (L2), but the
//     "everyone has published" signal is the hardware cluster barrier
//     (barrier.cluster.arrive.release / wait.acquire, ~0.2 us) instead of a global arrival
//     counter polled through L2 (measured ~2.5 us per step with fences on both sides),
//   * the streamed operand (h_{t-1}: 128 KB, dG_{t+1}: 512 KB per batch tile) is read from L2
//     ONCE per cluster: k-block kb is fetched by CTA (kb mod cluster size) with TMA multicast
//     into the same ring slot of every CTA (before: once per CTA, which made the backward
//     kernel L2-bandwidth-bound at 32 MB per step),
//   * ring slots are recycled cluster-wide: every CTA's MMA warp commits (tcgen05.commit
//     ...multicast::cluster) to the `empty` barrier of all CTAs, so the CTA whose turn it is to
//     refill a slot knows that all consumers are done with it.
// Batch tiles are independent clusters: no cross-cluster synchronisation at all.
#include <cuda_bf16.h>
#include "kernels.h"
#include "tc_common.cuh"

namespace mmqg {
namespace lc {

using namespace tc;
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// TMA 2-D load multicast to every CTA in `mask`: data and the mbarrier complete_tx land at the
// same shared-memory offsets in each destination CTA.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the same barrier offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// ---- warp-cooperative tile movers (see lstm_persist.cu) ------------------------------------
static constexpr int STG_LD = 20;
static constexpr int STG_WARP = 32 * STG_LD;

__device__ __forceinline__ void coop_ldg(const float* base, size_t row_stride, int rows_valid, int lane, float4 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    v[i] = r < rows_valid ? *reinterpret_cast<const float4*>(base + (size_t)r * row_stride + 4 * (lane & 3))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void coop_stg(float* base, size_t row_stride, int rows_valid, int lane, const float4 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    if (r < rows_valid) *reinterpret_cast<float4*>(base + (size_t)r * row_stride + 4 * (lane & 3)) = v[i];
  }
}
__device__ __forceinline__ void coop_to_row(float* stg, int lane, const float4 (&v)[4], float* mine) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(stg + (8 * i + (lane >> 2)) * STG_LD + 4 * (lane & 3)) = v[i];
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(stg + lane * STG_LD + 4 * q);
    mine[4 * q] = x.x; mine[4 * q + 1] = x.y; mine[4 * q + 2] = x.z; mine[4 * q + 3] = x.w;
  }
}
__device__ __forceinline__ void row_to_coop(float* stg, int lane, const float* mine, float4 (&v)[4]) {
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<float4*>(stg + lane * STG_LD + 4 * q) = make_float4(mine[4 * q], mine[4 * q + 1], mine[4 * q + 2], mine[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(stg + (8 * i + (lane >> 2)) * STG_LD + 4 * (lane & 3));
}
__device__ __forceinline__ void row_bf16_to_global(uint32_t* stg, int lane, const uint32_t (&w8)[8], bf16* base, size_t row_stride,
                                                   int rows_valid) {
  __syncwarp();
  *reinterpret_cast<uint4*>(stg + lane * 12) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
  *reinterpret_cast<uint4*>(stg + lane * 12 + 4) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = 16 * i + (lane >> 1), hsel = lane & 1;
    const uint4 x = *reinterpret_cast<const uint4*>(stg + r * 12 + 4 * hsel);
    if (r < rows_valid) *reinterpret_cast<uint4*>(base + (size_t)r * row_stride + 8 * hsel) = x;
  }
}

static constexpr int RING = 4;            // 16 KB stages of the streamed operand
static constexpr int NTHREADS = 320;      // warps 0-7 epilogue (2 per TMEM lane quarter), 8 producer, 9 MMA

struct FwdP {
  float* gates; float* cs; bf16* hs; float* mem; long long mem_ld;
  int T, B, H, KB, CS;
};

// grid (CS, n_mt), cluster (CS,1,1).  CTA rank r owns hidden units [32r, 32r+32).
__global__ void __launch_bounds__(NTHREADS, 1)
lstm_seq_fwd_cluster_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, FwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                          // KB x (128 rows x 128 B)
  uint8_t* sA = smem + p.KB * 16384;           // RING x (128 rows x 128 B)
  float* stg_all = reinterpret_cast<float*>(sA + RING * 16384);
  __shared__ uint64_t w_full, full[RING], empty[RING], mma_done, tmem_free;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int mt = blockIdx.y;
  const int H = p.H, B = p.B, G = 4 * p.H, CS = p.CS;
  const uint16_t mask = (uint16_t)((1u << CS) - 1u);

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int s = 0; s < RING; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CS); }
    mbar_init(&mma_done, 1);
    mbar_init(&tmem_free, 256);
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmH);
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  // every CTA's barriers must be initialised before any peer multicasts into them
  cluster_arrive();
  cluster_wait();

  if (warp == 8) {
    // ---- producer: W slice once, then the h_{t-1} k-blocks this CTA is responsible for ----
    if (lane == 0) {
      mbar_expect_tx(&w_full, p.KB * 16384);
      for (int kb = 0; kb < p.KB; ++kb) tma_load_2d(sW + kb * 16384, &tmW, &w_full, kb * 64, rank * 128);
    }
    int i = 0;
    for (int t = 0; t < p.T; ++t) {
      if (t > 0) {
        cluster_wait();                       // phase t-1: every CTA has published its h_{t-1} slice
        fence_proxy_async();
      }
      if (lane == 0) {
        for (int kb = 0; kb < p.KB; ++kb, ++i) {
          const int s = i % RING, ph = (i / RING) & 1;
          mbar_wait(&empty[s], ph ^ 1);       // all CS consumers released this slot
          mbar_expect_tx(&full[s], 16384);
          if (kb % CS == rank) tma_load_2d_mc(sA + s * 16384, &tmH, &full[s], kb * 64, t * B + mt * 128, mask);
        }
      }
      __syncwarp();
      cluster_arrive();                       // phase t (this warp has nothing to publish)
    }
    cluster_wait();                           // phase T-1
  } else if (warp == 9) {
    // ---- MMA issuer ----
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    if (lane == 0) mbar_wait(&w_full, 0);
    __syncwarp();
    int i = 0;
    for (int t = 0; t < p.T; ++t) {
      if (t > 0) cluster_wait();
      if (lane == 0) {
        if (t > 0) mbar_wait(&tmem_free, (t - 1) & 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < p.KB; ++kb, ++i) {
          const int s = i % RING, ph = (i / RING) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * 16384), b_addr = smem_u32(sW + kb * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_mc(&empty[s], mask);    // slot s is free in this CTA once these MMAs retire
        }
        umma_commit(&mma_done);
      }
      __syncwarp();
      cluster_arrive();
    }
    cluster_wait();
  } else {
    // ---- epilogue warps: lane quarter q = warp % 4 (rows), unit half = warp / 4 ----
    const int q = warp & 3, half = warp >> 2;
    const int m0w = mt * 128 + q * 32;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = rank * 32 + half * 16;
    float* stg = stg_all + warp * STG_WARP;
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    for (int t = 0; t < p.T; ++t) {
      float* gbase = p.gates + ((size_t)t * B + m0w) * G + j0;
      float4 gxc[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) coop_ldg(gbase + g * H, G, rows_valid, lane, gxc[g]);
      if (t > 0) cluster_wait();              // pairs with last step's arrive (already complete by now)
      mbar_wait(&mma_done, t & 1);
      tc_fence_after_sync();
      float acc[64];
#pragma unroll
      for (int g = 0; g < 4; ++g)
        tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + g * 32 + half * 16, acc + g * 16);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&tmem_free);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float gxr[16];
        coop_to_row(stg, lane, gxc[g], gxr);
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[g * 16 + u] += gxr[u];
      }
      float hv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float ig = sigm_fast(acc[u]);
        const float fg = sigm_fast(acc[16 + u]);
        const float gg = tanh_fast(acc[32 + u]);
        const float og = sigm_fast(acc[48 + u]);
        c[u] = fmaf(fg, c[u], ig * gg);
        hv[u] = og * tanh_fast(c[u]);
        acc[u] = ig; acc[16 + u] = fg; acc[32 + u] = gg; acc[48 + u] = og;
      }
      {
        uint32_t hp[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          __nv_bfloat162 t2 = __floats2bfloat162_rn(hv[2 * v], hv[2 * v + 1]);
          hp[v] = *reinterpret_cast<uint32_t*>(&t2);
        }
        row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, hp, p.hs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid);
      }
      __syncwarp();
      cluster_arrive();                       // release: this warp's h_t stores are published (phase t)
      {
        float4 tmp[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          row_to_coop(stg, lane, acc + g * 16, tmp);
          coop_stg(gbase + g * H, G, rows_valid, lane, tmp);
        }
        row_to_coop(stg, lane, c, tmp);
        coop_stg(p.cs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid, lane, tmp);
        if (p.mem) {
          row_to_coop(stg, lane, hv, tmp);
          coop_stg(p.mem + (size_t)m0w * p.mem_ld + (size_t)t * H + j0, (size_t)p.mem_ld, rows_valid, lane, tmp);
        }
      }
    }
    cluster_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
  // no CTA may exit while a peer could still multicast into its shared memory
  cluster_arrive();
  cluster_wait();
}

// ---------------------------------------------------------------------------------------------
struct BwdP {
  const float* acts; const float* cs; bf16* dg;
  const float* dh_ext; long long ext_ts, ext_ld;
  const float* dh_last; const float* dc_last;
  int T, B, H, NKB, CS;
};

// CTA rank r owns hidden units [32r, 32r+32): W_hh^T slice (32 x 4H) resident, dG_{t+1} streamed.
__global__ void __launch_bounds__(NTHREADS, 1)
lstm_seq_bwd_cluster_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, BwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                          // NKB x (32 rows x 128 B)
  uint8_t* sA = smem + p.NKB * 4096;           // RING x (128 rows x 128 B)
  float* stg_all = reinterpret_cast<float*>(sA + RING * 16384);
  __shared__ uint64_t w_full, full[RING], empty[RING], mma_done, tmem_free;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int mt = blockIdx.y;
  const int H = p.H, B = p.B, G = 4 * p.H, T = p.T, CS = p.CS;
  const uint16_t mask = (uint16_t)((1u << CS) - 1u);

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int s = 0; s < RING; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CS); }
    mbar_init(&mma_done, 1);
    mbar_init(&tmem_free, 256);
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 32);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  cluster_arrive();
  cluster_wait();

  // Step order t = T-1 .. 0.  Phase index ph = T-1-t: arrive(ph) after dG_t is published; the
  // loads of step t (t <= T-2) need phase ph-1 = dG_{t+1}.
  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(&w_full, p.NKB * 4096);
      for (int kb = 0; kb < p.NKB; ++kb) tma_load_2d(sW + kb * 4096, &tmW, &w_full, kb * 64, rank * 32);
    }
    int i = 0;
    for (int t = T - 1; t >= 0; --t) {
      if (t < T - 1) {
        cluster_wait();
        fence_proxy_async();
        if (lane == 0) {
          for (int kb = 0; kb < p.NKB; ++kb, ++i) {
            const int s = i % RING, ph = (i / RING) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], 16384);
            if (kb % CS == rank) tma_load_2d_mc(sA + s * 16384, &tmG, &full[s], kb * 64, (t + 1) * B + mt * 128, mask);
          }
        }
        __syncwarp();
      }
      cluster_arrive();
    }
    cluster_wait();
  } else if (warp == 9) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
    if (lane == 0) mbar_wait(&w_full, 0);
    __syncwarp();
    int i = 0, it = 0;
    for (int t = T - 1; t >= 0; --t) {
      if (t < T - 1) {
        cluster_wait();
        if (lane == 0) {
          if (it > 0) mbar_wait(&tmem_free, (it - 1) & 1);
          tc_fence_after_sync();
          for (int kb = 0; kb < p.NKB; ++kb, ++i) {
            const int s = i % RING, ph = (i / RING) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(sA + s * 16384), b_addr = smem_u32(sW + kb * 4096);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_mc(&empty[s], mask);
          }
          umma_commit(&mma_done);
        }
        __syncwarp();
        ++it;
      }
      cluster_arrive();
    }
    cluster_wait();
  } else {
    const int q = warp & 3, half = warp >> 2;
    const int m0w = mt * 128 + q * 32;
    const int m = m0w + lane;
    const bool valid = m < B;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = rank * 32 + half * 16;
    float* stg = stg_all + warp * STG_WARP;
    float dc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) dc[u] = (p.dc_last && valid) ? p.dc_last[(size_t)m * H + j0 + u] : 0.f;
    int it = 0;
    for (int t = T - 1; t >= 0; --t) {
      // cooperative prefetch of everything that does not depend on the recurrence
      float4 a4[4][4], cn4[4], cp4[4], ex4[4];
      const float* abase = p.acts + ((size_t)t * B + m0w) * G + j0;
#pragma unroll
      for (int g = 0; g < 4; ++g) coop_ldg(abase + g * H, G, rows_valid, lane, a4[g]);
      coop_ldg(p.cs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid, lane, cn4);
      coop_ldg(p.cs + ((size_t)t * B + m0w) * H + j0, H, rows_valid, lane, cp4);
      if (p.dh_ext) coop_ldg(p.dh_ext + (size_t)t * p.ext_ts + (size_t)m0w * p.ext_ld + j0, (size_t)p.ext_ld, rows_valid, lane, ex4);
      else {
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) ex4[i2] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float dh[16];
      if (t == T - 1) {
#pragma unroll
        for (int u = 0; u < 16; ++u) dh[u] = (p.dh_last && valid) ? p.dh_last[(size_t)m * H + j0 + u] : 0.f;
      } else {
        cluster_wait();                        // pairs with the previous step's arrive
        mbar_wait(&mma_done, it & 1);
        tc_fence_after_sync();
        tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + half * 16, dh);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(&tmem_free);
        ++it;
      }
      float a[64], cn[16], cp[16], ex[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) coop_to_row(stg, lane, a4[g], a + g * 16);
      coop_to_row(stg, lane, cn4, cn);
      coop_to_row(stg, lane, cp4, cp);
      coop_to_row(stg, lane, ex4, ex);
      float dgv[64];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float ig = a[u], fg = a[16 + u], gg = a[32 + u], og = a[48 + u];
        const float d = dh[u] + ex[u];
        const float tc_ = tanh_fast(cn[u]);
        const float dct = dc[u] + d * og * (1.f - tc_ * tc_);
        dgv[u] = dct * gg * ig * (1.f - ig);
        dgv[16 + u] = dct * cp[u] * fg * (1.f - fg);
        dgv[32 + u] = dct * ig * (1.f - gg * gg);
        dgv[48 + u] = d * tc_ * og * (1.f - og);
        dc[u] = dct * fg;
      }
      bf16* dbase = p.dg + ((size_t)t * B + m0w) * G + j0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t w8[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          __nv_bfloat162 t2 = __floats2bfloat162_rn(dgv[g * 16 + 2 * v], dgv[g * 16 + 2 * v + 1]);
          w8[v] = *reinterpret_cast<uint32_t*>(&t2);
        }
        row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, w8, dbase + g * H, G, rows_valid);
      }
      __syncwarp();
      cluster_arrive();                        // dG_t published
    }
    cluster_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 32);
  }
  cluster_arrive();
  cluster_wait();
}

// 32-unit gate-slice packing for the forward kernel: row (r*128 + g*32 + u) = W_hh[g*H + 32r + u, :]
__global__ void pack_whh_fwd32_kernel(const float* __restrict__ w, bf16* __restrict__ out, int H) {
  const int row = blockIdx.x;
  const int r = row / 128, g = (row % 128) / 32, u = row % 32;
  const float* src = w + (size_t)(g * H + r * 32 + u) * H;
  bf16* dst = out + (size_t)row * H;
  for (int k = threadIdx.x; k < H; k += blockDim.x) dst[k] = __float2bfloat16_rn(src[k]);
}

}  // namespace lc

// ---- host -----------------------------------------------------------------------------------
bool lstm_cluster_ok(int B, int H) {
  if (H % 32 != 0) return false;
  const int cs = H / 32;
  if (cs != 2 && cs != 4 && cs != 8 && cs != 16) return false;
  (void)B;
  return true;
}

int pack_whh_cluster(const float* w_hh, void* fwd_packed, int H, cudaStream_t st) {
  MMQG_REQUIRE(w_hh && fwd_packed && H % 32 == 0, "pack_whh_cluster: bad args");
  lc::pack_whh_fwd32_kernel<<<4 * H, 128, 0, st>>>(w_hh, reinterpret_cast<lc::bf16*>(fwd_packed), H);
  MMQG_LAUNCH_CHECK();
  return 0;
}

template <typename Kern, typename P>
static int launch_cluster(Kern kern, dim3 grid, int cs, size_t smem, const CUtensorMap& m0, const CUtensorMap& m1, const P& p,
                          cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(lc::NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MMQG_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, p));
  return 0;
}

int lstm_seq_fwd_cluster(float* gates, float* cs, void* hs, const void* wp_fwd32, float* mem, long long mem_ld, int T, int B,
                         int H, cudaStream_t st) {
  MMQG_REQUIRE(lstm_cluster_ok(B, H), "lstm_seq_fwd_cluster: shape B=%d H=%d not supported", B, H);
  const int CS = H / 32, KB = H / 64, n_mt = ceil_div(B, 128);
  lc::FwdP p{gates, cs, reinterpret_cast<lc::bf16*>(hs), mem, mem_ld, T, B, H, KB, CS};
  CUtensorMap tmW, tmH;
  MMQG_TRY(make_tmap_bf16_2d(&tmW, wp_fwd32, 4 * (uint64_t)H, H, H, 128, 64));
  MMQG_TRY(make_tmap_bf16_2d(&tmH, hs, (uint64_t)(T + 1) * B, H, H, 128, 64));
  const size_t smem = (size_t)KB * 16384 + lc::RING * 16384 + 8 * lc::STG_WARP * sizeof(float) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(lc::lstm_seq_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   8 * 16384 + lc::RING * 16384 + 8 * lc::STG_WARP * (int)sizeof(float) + 1024));
    MMQG_CUDA(cudaFuncSetAttribute(lc::lstm_seq_fwd_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr = true;
  }
  MMQG_PROBE(KC_LSTM_PERSIST, 2.0 * T * B * 4.0 * H * H, 0);
  MMQG_TRY(launch_cluster(lc::lstm_seq_fwd_cluster_kernel, dim3(CS, n_mt), CS, smem, tmW, tmH, p, st));
  MMQG_LAUNCH_CHECK();
  return 0;
}

int lstm_seq_bwd_cluster(const float* acts, const float* cs, void* dg, const void* wp_bwd, const float* dh_ext, long long ext_ts,
                         long long ext_ld, const float* dh_last, const float* dc_last, int T, int B, int H, cudaStream_t st) {
  MMQG_REQUIRE(lstm_cluster_ok(B, H), "lstm_seq_bwd_cluster: shape B=%d H=%d not supported", B, H);
  const int CS = H / 32, NKB = 4 * H / 64, n_mt = ceil_div(B, 128);
  lc::BwdP p{acts, cs, reinterpret_cast<lc::bf16*>(dg), dh_ext, ext_ts, ext_ld, dh_last, dc_last, T, B, H, NKB, CS};
  CUtensorMap tmW, tmG;
  MMQG_TRY(make_tmap_bf16_2d(&tmW, wp_bwd, H, 4 * (uint64_t)H, 4 * (uint64_t)H, 32, 64));
  MMQG_TRY(make_tmap_bf16_2d(&tmG, dg, (uint64_t)T * B, 4 * (uint64_t)H, 4 * (uint64_t)H, 128, 64));
  const size_t smem = (size_t)NKB * 4096 + lc::RING * 16384 + 8 * lc::STG_WARP * sizeof(float) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(lc::lstm_seq_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   32 * 4096 + lc::RING * 16384 + 8 * lc::STG_WARP * (int)sizeof(float) + 1024));
    MMQG_CUDA(cudaFuncSetAttribute(lc::lstm_seq_bwd_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr = true;
  }
  MMQG_PROBE(KC_LSTM_PERSIST, 2.0 * (T - 1) * B * 4.0 * H * H, 0);
  MMQG_TRY(launch_cluster(lc::lstm_seq_bwd_cluster_kernel, dim3(CS, n_mt), CS, smem, tmW, tmG, p, st));
  MMQG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmqg

// One LSTM layer step of the decoder as ONE launch (bf16 mode): the gate pre-activations
//   pre(B,4H) = X W_in^T + h_prev W_hh^T (+ hoisted addend) (+ bias)
// are produced by tcgen05.mma into TMEM exactly as in gemm_tc.cu, but the CTA's 64 accumulator
// columns are the four gates of the SAME 16 hidden units (the B operand is fetched as four
// 16-row TMA boxes, one per gate block of the PyTorch-layout weight, so no re-packing is needed)
// and the epilogue applies the cell update in registers:
//   i,f,o = sigmoid, g = tanh, c' = f c + i g, h' = o tanh(c')      (aten::lstm cell, gate order i,f,g,o;
//   reference decoder.py:100 via torch.nn.LSTM)
// writing the activated gates (kept for the backward pass), c', h' (bf16) and, with inter-layer
// dropout, the dropped copy of h' -- what used to be a GEMM launch plus a pointwise launch.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2-9 epilogue
// (TMEM lane quarter = warp % 4; the two warps of a quarter take 8 units each).
#include <cuda_bf16.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "dropout.cuh"

namespace mmqg {

using namespace tc;
typedef __nv_bfloat16 bf16;

namespace {

constexpr int SBM = 128, SBK = 64, SBN = 64, SUNITS = 16, SSTAGES = 8;
constexpr int SA_BYTES = SBM * SBK * 2, SB_BYTES = SBN * SBK * 2;

struct LstmStepP {
  int M, H, nk1, nk2;
  const float* bias;                 // (4H) or null
  const float* pre; int ldpre;       // optional addend to the pre-activations (may alias acts)
  float* acts; int ldg;
  const float* c_prev; int ldcp;     // null = zero state
  float* c_out; int ldc;
  bf16* h_out; int ldh;
  DropSpec dr;
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }


__global__ void __launch_bounds__(320, 1)
lstm_step_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, LstmStepP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + SSTAGES * SA_BYTES;
  __shared__ uint64_t full[SSTAGES], empty[SSTAGES], accum_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * SBM, u0 = blockIdx.x * SUNITS;
  const int nk = p.nk1 + p.nk2;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&accum_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (p.nk2 > 0) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
  }
  if (warp == 1) tmem_alloc(&tmem_slot, SBN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % SSTAGES, ph = (kb / SSTAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], SA_BYTES + SB_BYTES);
        const bool second = kb >= p.nk1;
        const int k0 = (second ? kb - p.nk1 : kb) * SBK;
        const CUtensorMap* ma = second ? &tmA2 : &tmA;
        const CUtensorMap* mb = second ? &tmB2 : &tmB;
        tma_load_2d(sA + s * SA_BYTES, ma, &full[s], k0, m0);
        // weight rows g*H + u0 .. +16 of each gate block -> tile rows 16g .. 16g+15 (2 KB apiece,
        // whole 8-row swizzle groups, so the tile looks like one 64-row K-major box to the MMA)
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_2d(sB + s * SB_BYTES + g * (SUNITS * 128), mb, &full[s], k0, g * p.H + u0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(SBM, SBN, 0, 0);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % SSTAGES, ph = (kb / SSTAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(sA + s * SA_BYTES), b_addr = smem_u32(sB + s * SB_BYTES);
#pragma unroll
        for (int k = 0; k < SBK / 16; ++k)
          umma_bf16(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(&accum_full);
    }
  } else {
    // ---- cell-update epilogue: thread = one batch row x 8 hidden units x 4 gates ----
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = m0 + 32 * q + lane, u = u0 + 8 * half;
    const bool live = r < p.M;
    // operands that do not depend on the accumulator are fetched while the MMAs run
    float cp[8], add[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cp[j] = 0.f;
    if (live && p.c_prev) {
      const float4* s4 = reinterpret_cast<const float4*>(p.c_prev + (size_t)r * p.ldcp + u);
      const float4 a = s4[0], b = s4[1];
      cp[0] = a.x; cp[1] = a.y; cp[2] = a.z; cp[3] = a.w; cp[4] = b.x; cp[5] = b.y; cp[6] = b.z; cp[7] = b.w;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int j = 0; j < 8; ++j) add[g][j] = 0.f;
      if (p.bias) {
        const float4* s4 = reinterpret_cast<const float4*>(p.bias + g * p.H + u);
        const float4 a = __ldg(s4), b = __ldg(s4 + 1);
        add[g][0] = a.x; add[g][1] = a.y; add[g][2] = a.z; add[g][3] = a.w;
        add[g][4] = b.x; add[g][5] = b.y; add[g][6] = b.z; add[g][7] = b.w;
      }
      if (live && p.pre) {
        const float4* s4 = reinterpret_cast<const float4*>(p.pre + (size_t)r * p.ldpre + g * p.H + u);
        const float4 a = s4[0], b = s4[1];
        add[g][0] += a.x; add[g][1] += a.y; add[g][2] += a.z; add[g][3] += a.w;
        add[g][4] += b.x; add[g][5] += b.y; add[g][6] += b.z; add[g][7] += b.w;
      }
    }
    mbar_wait(&accum_full, 0);
    tc_fence_after_sync();
    float v[4][8];
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + 8 * half;
#pragma unroll
    for (int g = 0; g < 4; ++g) tmem_ld_32x8(taddr + g * SUNITS, v[g]);
    tmem_ld_wait();
    if (live) {
      float cn[8], hn[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float i = sigm(v[0][j] + add[0][j]), f = sigm(v[1][j] + add[1][j]);
        const float gg = tanhf(v[2][j] + add[2][j]), o = sigm(v[3][j] + add[3][j]);
        const float c = f * cp[j] + i * gg;
        v[0][j] = i; v[1][j] = f; v[2][j] = gg; v[3][j] = o;
        cn[j] = c;
        hn[j] = o * tanhf(c);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float4* d4 = reinterpret_cast<float4*>(p.acts + (size_t)r * p.ldg + g * p.H + u);
        d4[0] = make_float4(v[g][0], v[g][1], v[g][2], v[g][3]);
        d4[1] = make_float4(v[g][4], v[g][5], v[g][6], v[g][7]);
      }
      float4* c4 = reinterpret_cast<float4*>(p.c_out + (size_t)r * p.ldc + u);
      c4[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
      c4[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
      __nv_bfloat162 hb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) hb[j] = __floats2bfloat162_rn(hn[2 * j], hn[2 * j + 1]);
      *reinterpret_cast<uint4*>(p.h_out + (size_t)r * p.ldh + u) = *reinterpret_cast<const uint4*>(hb);
      if (p.dr.out) {
        const float inv_keep = 1.0f / (1.0f - p.dr.p);
        const unsigned long long sd = p.dr.seed + (p.dr.ctr ? *p.dr.ctr : 0ull);
        const unsigned long long base = p.dr.base + (unsigned long long)r * p.H + u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          hb[j] = __floats2bfloat162_rn(hn[2 * j] * drop_scale(sd, p.dr.sid, base + 2 * j, p.dr.p, inv_keep),
                                        hn[2 * j + 1] * drop_scale(sd, p.dr.sid, base + 2 * j + 1, p.dr.p, inv_keep));
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.dr.out) + (size_t)r * p.dr.ld + u) = *reinterpret_cast<const uint4*>(hb);
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, SBN);
  }
}

}  // namespace

bool lstm_step_tc_ok(int H, int ldg, int ldc, int ldh, int ldcp, int ldpre) {
  return H % 16 == 0 && ldg % 4 == 0 && ldc % 4 == 0 && ldh % 8 == 0 && ldcp % 4 == 0 && ldpre % 4 == 0;
}

int lstm_step_tc(const void* X, int ldx, int K1, const void* W_in, int ldw1, const void* Hprev, int ldhp, const void* W_hh,
                 int ldw2, const float* bias, const float* pre, int ldpre, float* acts, int ldg, const float* c_prev, int ldcp,
                 float* c_out, int ldc, void* h_out, int ldh, int B, int H, DropSpec dr, cudaStream_t st) {
  MMQG_REQUIRE(X && W_in && Hprev && W_hh && acts && c_out && h_out && B > 0 && H > 0 && K1 > 0, "lstm_step_tc: bad args");
  MMQG_REQUIRE(lstm_step_tc_ok(H, ldg, ldc, ldh, c_prev ? ldcp : 4, pre ? ldpre : 4), "lstm_step_tc: unsupported shape H=%d", H);
  MMQG_REQUIRE(!dr.out || dr.ld % 8 == 0, "lstm_step_tc: dropped copy needs a row pitch multiple of 8");
  CUtensorMap ta, tb, ta2, tb2;
  MMQG_TRY(make_tmap_bf16_2d(&ta, X, B, K1, ldx, SBM, 64));
  MMQG_TRY(make_tmap_bf16_2d(&tb, W_in, 4 * (uint64_t)H, K1, ldw1, SUNITS, 64));
  MMQG_TRY(make_tmap_bf16_2d(&ta2, Hprev, B, H, ldhp, SBM, 64));
  MMQG_TRY(make_tmap_bf16_2d(&tb2, W_hh, 4 * (uint64_t)H, H, ldw2, SUNITS, 64));
  LstmStepP p;
  p.M = B; p.H = H; p.nk1 = ceil_div(K1, SBK); p.nk2 = ceil_div(H, SBK);
  p.bias = bias; p.pre = pre; p.ldpre = ldpre; p.acts = acts; p.ldg = ldg; p.c_prev = c_prev; p.ldcp = ldcp;
  p.c_out = c_out; p.ldc = ldc; p.h_out = reinterpret_cast<bf16*>(h_out); p.ldh = ldh; p.dr = dr;
  constexpr int SMEM = SSTAGES * (SA_BYTES + SB_BYTES) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(lstm_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  MMQG_PROBE(tl_gemm_class, 2.0 * B * 4.0 * H * ((double)K1 + H),
             2.0 * ((double)B + 4.0 * H) * ((double)K1 + H) + 4.0 * B * 4.0 * H);
  MMQG_CUDA(launch_k(lstm_step_tc_kernel, dim3(H / SUNITS, ceil_div(B, SBM)), dim3(320), SMEM, st, ta, tb, ta2, tb2, p));
  MMQG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmqg

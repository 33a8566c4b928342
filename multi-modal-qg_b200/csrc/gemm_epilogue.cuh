// Epilogue shared by the tcgen05 GEMM kernels: one warp drains a 32-row x (32*NCH)-column block
// of a TMEM accumulator to global memory.
//
// tcgen05.ld hands each thread one accumulator ROW (32 consecutive fp32 columns), global memory
// wants lanes along the columns.  Per 32x32 chunk: TMEM -> registers -> padded smem tile
// (STS.128, row pitch 36 floats) -> LDS.128 in the transposed role (8 lanes cover the 32 columns
// of one row, 4 rows per instruction) -> alpha/bias/addend -> STG.128 (fp32) or STG.64 (bf16).
// 8+8+8 memory instructions per chunk instead of 32+32+32 scalar ones: with one epilogue warp
// per scheduler the scalar version ran at ~1 output element per clock per SM.
#pragma once
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace mmqg {

static constexpr int EPI_LD = 36;                 // staged row pitch in floats (32 + 4, keeps 16-B alignment)
static constexpr int EPI_STG_FLOATS = 32 * EPI_LD; // per warp

struct EpiOut {
  float* Cf; __nv_bfloat16* Cb; int ldc; int c_bf16;
  const float* Cin; int ldcin; float beta, alpha;
  const float* bias;
  int M, N;
  bool vec;        // N, ldc (and ldcin) multiples of 4 and base pointers aligned: vector path
};

__device__ __forceinline__ EpiOut make_epi_out(const TcGemmP& p, int zslice, bool lead) {
  EpiOut o;
  o.Cf = reinterpret_cast<float*>(p.C) + (size_t)zslice * p.c_split_stride;
  o.Cb = reinterpret_cast<__nv_bfloat16*>(p.C);
  o.ldc = p.ldc; o.c_bf16 = p.c_bf16;
  o.Cin = lead ? p.Cin : nullptr; o.ldcin = p.ldcin; o.beta = p.beta; o.alpha = p.alpha;
  o.bias = lead ? p.bias : nullptr;
  o.M = p.M; o.N = p.N;
  const uintptr_t cb = reinterpret_cast<uintptr_t>(p.C) + (size_t)zslice * p.c_split_stride * 4;
  bool v = (p.N % 4 == 0) && (p.ldc % 4 == 0) && (cb % (p.c_bf16 ? 8 : 16) == 0);
  if (o.Cin) v = v && (p.ldcin % 4 == 0) && (reinterpret_cast<uintptr_t>(p.Cin) % 16 == 0);
  if (o.bias) v = v && (reinterpret_cast<uintptr_t>(p.bias) % 16 == 0);
  o.vec = v;
  return o;
}

// taddr: TMEM address of (first lane of this warp's quarter, first column of the block).
// (mrow0, ncol0): global coordinates of the block's first element.  If release_bar != nullptr every
// lane arrives on it right after the last TMEM read (hands the accumulator back to the MMA warp).
template <int NCH>
__device__ __forceinline__ void epilogue_block(uint32_t taddr, float* stg, const EpiOut& o, int mrow0, int ncol0, int lane,
                                               uint64_t* release_bar) {
  using namespace tc;
  const int rows = min(32, o.M - mrow0);
  const int cg = lane & 7, rs = lane >> 3;         // column group (4 columns) and row sub-index of the store role
  float4 cin[2][8];
  auto load_cin = [&](int c, float4 (&dst)[8]) {
    const int n = ncol0 + c * 32 + 4 * cg;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + rs;
      dst[i] = (r < rows && n < o.N) ? *reinterpret_cast<const float4*>(o.Cin + (size_t)(mrow0 + r) * o.ldcin + n)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (o.vec) {
    if (o.Cin) load_cin(0, cin[0]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (c == NCH - 1 && release_bar) {
          tc_fence_before_sync();
          mbar_arrive(release_bar);
        }
        __syncwarp();
#pragma unroll
        for (int qd = 0; qd < 8; ++qd)
          *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * qd) = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
        __syncwarp();
      }
      if (o.Cin && c + 1 < NCH) load_cin(c + 1, cin[(c + 1) & 1]);
      const int n = ncol0 + c * 32 + 4 * cg;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o.bias && n < o.N) b4 = *reinterpret_cast<const float4*>(o.bias + n);
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 s = *reinterpret_cast<const float4*>(stg + (4 * i + rs) * EPI_LD + 4 * cg);
        x[i] = make_float4(fmaf(o.alpha, s.x, b4.x), fmaf(o.alpha, s.y, b4.y), fmaf(o.alpha, s.z, b4.z), fmaf(o.alpha, s.w, b4.w));
        if (o.Cin) {
          const float4 a = cin[c & 1][i];
          x[i].x = fmaf(o.beta, a.x, x[i].x); x[i].y = fmaf(o.beta, a.y, x[i].y);
          x[i].z = fmaf(o.beta, a.z, x[i].z); x[i].w = fmaf(o.beta, a.w, x[i].w);
        }
      }
      if (n < o.N) {
        if (o.c_bf16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rs;
            if (r < rows) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(x[i].x, x[i].y), hi = __floats2bfloat162_rn(x[i].z, x[i].w);
              uint2 u;
              u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(o.Cb + (size_t)(mrow0 + r) * o.ldc + n) = u;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rs;
            if (r < rows) *reinterpret_cast<float4*>(o.Cf + (size_t)(mrow0 + r) * o.ldc + n) = x[i];
          }
        }
      }
    }
  } else {
    // generic path (ragged N / unaligned views): scalar, lanes along the columns
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (c == NCH - 1 && release_bar) {
          tc_fence_before_sync();
          mbar_arrive(release_bar);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * EPI_LD + j] = v[j];
        __syncwarp();
      }
      const int n = ncol0 + c * 32 + lane;
      if (n < o.N) {
        const float b = o.bias ? o.bias[n] : 0.f;
        for (int r = 0; r < rows; ++r) {
          const size_t m = (size_t)(mrow0 + r);
          float x = fmaf(o.alpha, stg[r * EPI_LD + lane], b);
          if (o.Cin) x = fmaf(o.beta, o.Cin[m * o.ldcin + n], x);
          if (o.c_bf16) o.Cb[m * o.ldc + n] = __float2bfloat16_rn(x);
          else o.Cf[m * o.ldc + n] = x;
        }
      }
    }
  }
}

}  // namespace mmqg

// Device helpers shared by the persistent decoder-step kernels (dec_persist.cu forward, dec_persist_bwd.cu BPTT):
// fast activations, warp reductions, named barriers, bulk global->shared copies, arrival-counter waits and the
// warp-cooperative tile movers (4 lanes share a 64-byte row segment, as in lstm_persist.cu).
#pragma once
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace mmqg {
namespace dp {

using namespace tc;
typedef __nv_bfloat16 bf16;


__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void bar_epi(int nthreads) { asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void wait_count(const uint32_t* f, uint32_t target) {
  wait_counter_acquire(f, target);          // relaxed poll + one acquire load (tc_common.cuh)
}

// warp-cooperative tile movers (same scheme as lstm_persist.cu: 4 lanes share a 64-byte row segment)
static constexpr int STG_LD = 20;
static constexpr int STG_WARP = 32 * STG_LD;
__device__ __forceinline__ void coop_ldg(const float* base, size_t row_stride, int rows_valid, int lane, float4 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    v[i] = r < rows_valid ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)r * row_stride + 4 * (lane & 3)))
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

static constexpr int MAXL = 3;
static constexpr int ASLOT = 24 * 1024;         // bytes of one attention-memory chunk

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// chunks of one sample's attention memories, in the order loader and workers both walk them:
// text rows, audio rows, video rows
struct Chunker {
  int cr_t, cr_a, cr_v, n_t, n_a, n_v;
  __device__ Chunker(int H, int H_a, int H_v, int T_t, int T_v) {
    cr_t = ASLOT / (H * 2); cr_a = ASLOT / (H_a * 4); cr_v = ASLOT / (H_v * 2);
    n_t = (T_t + cr_t - 1) / cr_t; n_a = (T_v + cr_a - 1) / cr_a; n_v = (T_v + cr_v - 1) / cr_v;
  }
  __device__ int count() const { return n_t + n_a + n_v; }
};

}  // namespace dp
}  // namespace mmqg

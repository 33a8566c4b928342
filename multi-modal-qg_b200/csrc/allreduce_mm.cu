// Gradient all-reduce over NVSwitch multicast memory (the path's only exchange step).
//
// The gradient buckets live in one symmetric allocation that every rank has mapped twice: its own copy, and a MULTICAST
// address that stands for all ranks' copies at once.  One kernel per bucket and rank, a few CTAs wide so that it fits on
// the SMs the two cooperative 64-CTA recurrent kernels leave free (NCCL's all-reduce kernels take SMs those launches wait for):
//
//   barrier (same-index CTAs of all ranks, flags in the peers' signal pads over NVLink)   -- every rank's bucket is final
//   rank r owns the r-th slice:  v = multimem.ld_reduce.add(mc + i)   -- ONE load, the switch sums the W copies in flight
//                                multimem.st(mc + i, v)               -- ONE store, the switch writes all W copies
//   barrier                                                           -- every slice has landed everywhere
//
// Bytes on each GPU's links: (W-1)/W of the bucket out (its share of the peers' reductions) + 1/W out (its slice, broadcast
// by the switch), the same in -- half of what a ring moves, and no intermediate buffers.  Sums are fp32 in switch order
// (deterministic per topology, not bit-identical to NCCL's ring order).
#include "kernels.h"

namespace mmqg {

__device__ __forceinline__ void sig_put(uint32_t* addr) {       // release: 0 -> 1, waits until the consumer has taken the last one
  __threadfence_system();
  while (atomicCAS_system(addr, 0u, 1u) != 0u) {}
}
__device__ __forceinline__ void sig_wait(uint32_t* addr) {      // acquire: 1 -> 0
  while (atomicCAS_system(addr, 1u, 0u) != 1u) {}
  __threadfence_system();
}

// sig[p]: rank p's signal pad (peer-mapped); slot (blockIdx.x * W + sender) of the RECEIVER's pad
__device__ __forceinline__ void rank_barrier(uint32_t* const* sig, int rank, int W) {
  __syncthreads();
  if ((int)threadIdx.x < W) {
    sig_put(sig[threadIdx.x] + blockIdx.x * W + rank);
    sig_wait(sig[rank] + blockIdx.x * W + threadIdx.x);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mm_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(512)
allreduce_mm_kernel(float* __restrict__ mc, long long n4, uint32_t* const* __restrict__ sig, int rank, int W) {
  rank_barrier(sig, rank, W);
  const long long per = (n4 + W - 1) / W;
  const long long lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < hi; i += 4 * stride) {        // four independent 16-byte reductions in flight per thread
    const float4 a = mm_ld_reduce(mc + 4 * i), b = mm_ld_reduce(mc + 4 * (i + stride)), c = mm_ld_reduce(mc + 4 * (i + 2 * stride)),
                 d = mm_ld_reduce(mc + 4 * (i + 3 * stride));
    mm_st(mc + 4 * i, a); mm_st(mc + 4 * (i + stride), b); mm_st(mc + 4 * (i + 2 * stride), c); mm_st(mc + 4 * (i + 3 * stride), d);
  }
  for (; i < hi; i += stride) mm_st(mc + 4 * i, mm_ld_reduce(mc + 4 * i));
  __threadfence_system();
  rank_barrier(sig, rank, W);
}

}  // namespace mmqg

// mc: multicast address of the bucket (16-byte aligned), n floats (multiple of 4); signal_pads: DEVICE array of `world` pointers
// to the ranks' signal pads (each at least ctas * world * 4 bytes, zero when no call is in flight); every rank calls this with
// the same n and ctas for the same bucket, in the same order.
extern "C" int mmqg_allreduce_multimem(float* mc, long long n, void* const* signal_pads, int rank, int world, int ctas, void* stream) {
  using namespace mmqg;
  MMQG_REQUIRE(mc && signal_pads && n > 0 && n % 4 == 0 && world >= 2 && world <= 32 && rank >= 0 && rank < world && ctas >= 1 && ctas <= 64,
               "allreduce_multimem: bad args");
  MMQG_REQUIRE(reinterpret_cast<uintptr_t>(mc) % 16 == 0, "allreduce_multimem: bucket is not 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  MMQG_PROBE(KC_OTHER, 0, 8.0 * n);
  allreduce_mm_kernel<<<ctas, 512, 0, st>>>(mc, n / 4, reinterpret_cast<uint32_t* const*>(signal_pads), rank, world);
  MMQG_LAUNCH_CHECK();
  return 0;
}

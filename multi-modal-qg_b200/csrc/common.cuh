// Shared helpers for libmmqg.so (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <type_traits>
#include <atomic>
#include "../../include/mmqg.h"

namespace mmqg {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

int set_err(int code, const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Kernel classes for launch accounting / the timing probe (probe.cu, mmqg_probe_start()).
enum KernelClass {
  KC_OTHER = 0,
  KC_GEMM_STEP = 1,   // per-timestep recurrent products (B x 4H x H and their backward twins)
  KC_GEMM_SEQ = 2,    // hoisted whole-sequence products (input projections, weight gradients, loss head)
  KC_POINTWISE = 3,   // LSTM cell update / its gradient
  KC_ATTN = 4,        // fused attention step forward / backward
  KC_LOSS = 5,        // log-softmax + NLL (+ dlogits) over a row chunk
  KC_EMBED = 6,       // embedding gather / scatter-add
  KC_LSTM_PERSIST = 7,  // persistent recurrent-cell kernels (one launch per layer and direction)
  KC_COUNT = 8
};
extern thread_local int tl_gemm_class;    // class the next gemm_f32() launch is booked under
void probe_open(int cls, cudaStream_t st, double flops, double bytes);
void probe_close(cudaStream_t st);
struct StepGemmScope {                    // RAII: GEMMs issued inside are per-timestep products
  int prev;
  StepGemmScope() : prev(tl_gemm_class) { tl_gemm_class = KC_GEMM_STEP; }
  ~StepGemmScope() { tl_gemm_class = prev; }
};
// Programmatic dependent launch (PDL): inside the decoder's per-step loops every kernel is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, calls pdl_launch_dependents()
// at its start and pdl_wait() before touching global memory, so its launch latency, block
// scheduling and prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the tail
// of its predecessor.  Both calls are no-ops for kernels launched the ordinary way.
extern thread_local int tl_pdl;
struct PdlScope {
  int prev;
  explicit PdlScope(bool on) : prev(tl_pdl) { tl_pdl = on ? 1 : 0; }
  ~PdlScope() { tl_pdl = prev; }
};
bool pdl_enabled();
// L2 persistence window for the launches issued inside an L2WindowScope (the decoder's attention
// steps re-read the same bf16 memories every step: decoder.py:81,87,95): accesses to [base, base+bytes)
// are marked persisting, everything else streaming.  The set-aside is sized once (l2_window_reserve).
extern thread_local const void* tl_l2win_base;
extern thread_local size_t tl_l2win_bytes;
struct L2WindowScope {
  const void* pb; size_t pn;
  L2WindowScope(const void* base, size_t bytes) : pb(tl_l2win_base), pn(tl_l2win_bytes) { tl_l2win_base = base; tl_l2win_bytes = bytes; }
  ~L2WindowScope() { tl_l2win_base = pb; tl_l2win_bytes = pn; }
};
// Reserves a persisting-L2 set-aside of at least `bytes` (clamped to the device maximum); returns window_bytes
// if a window of that size can be attached to a launch, else 0 (persistence unavailable, window too large).
size_t l2_window_reserve(size_t bytes, size_t window_bytes);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (tl_pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (tl_l2win_base && tl_l2win_bytes) {
    at[n].id = cudaLaunchAttributeAccessPolicyWindow;
    at[n].val.accessPolicyWindow.base_ptr = const_cast<void*>(tl_l2win_base);
    at[n].val.accessPolicyWindow.num_bytes = tl_l2win_bytes;
    at[n].val.accessPolicyWindow.hitRatio = 1.0f;
    at[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// Place directly before a launch on stream `st`; MMQG_LAUNCH_CHECK() closes it.
#define MMQG_PROBE(cls, flops, bytes) ::mmqg::probe_open((cls), st, (double)(flops), (double)(bytes))

// Call after every kernel launch: counts it and turns a launch error into a status.
#define MMQG_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    ::mmqg::probe_close(st);                                                             \
    ::mmqg::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess)                                                               \
      return ::mmqg::set_err(MMQG_ERR_CUDA, "%s:%d launch failed: %s", __FILE__, __LINE__, \
                             cudaGetErrorString(e_));                                    \
  } while (0)

#define MMQG_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return ::mmqg::set_err(MMQG_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,   \
                             cudaGetErrorString(e_));                                    \
  } while (0)

#define MMQG_REQUIRE(cond, ...)                                                          \
  do {                                                                                   \
    if (!(cond)) return ::mmqg::set_err(MMQG_ERR_BAD_ARG, __VA_ARGS__);                  \
  } while (0)

#define MMQG_TRY(call)                                                                   \
  do {                                                                                   \
    int s_ = (call);                                                                     \
    if (s_ != 0) return s_;                                                              \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace mmqg

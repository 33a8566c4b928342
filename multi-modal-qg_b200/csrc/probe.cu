// Per-kernel-class timing probe: CUDA events recorded around every launch of ONE selected
// kernel class, on the stream the kernel is launched on, so bench.py can report the
// achieved FLOP/s or GB/s of the dominant kernel measured live inside a run of the same
// step (B200_PROFILING.md "Roofline arithmetic").  Off by default: zero overhead beyond a
// branch.  Not for use under CUDA-graph capture.
#include <stdlib.h>
#include <vector>
#include "common.cuh"

namespace mmqg {

thread_local int tl_gemm_class = KC_GEMM_SEQ;
thread_local int tl_pdl = 0;

thread_local const void* tl_l2win_base = nullptr;
thread_local size_t tl_l2win_bytes = 0;

size_t l2_window_reserve(size_t bytes, size_t window_bytes) {
  static long long max_window = -1, reserved = 0, max_persist = 0;
  if (max_window < 0) {
    const char* e = getenv("MMQG_L2WIN");
    int dev = 0, mw = 0, mp = 0;
    cudaGetDevice(&dev);
    if ((e && e[0] == '0') || cudaDeviceGetAttribute(&mw, cudaDevAttrMaxAccessPolicyWindowSize, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&mp, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess) {
      cudaGetLastError();
      mw = mp = 0;
    }
    max_window = mw;
    max_persist = mp;
  }
  if (max_window <= 0 || max_persist <= 0) return 0;
  long long want = (long long)bytes < max_persist ? (long long)bytes : max_persist;
  if (want > reserved) {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want) != cudaSuccess) {
      cudaGetLastError();
      max_window = 0;
      return 0;
    }
    reserved = want;
  }
  return (long long)window_bytes <= max_window ? window_bytes : 0;      // a clipped window would leave rows unpinned
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MMQG_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

namespace {
struct ProbeState {
  int cls = -1;                       // selected class, -1 = off
  std::vector<cudaEvent_t> ev;        // start/end pairs
  size_t used = 0;                    // events used
  bool open = false;
  unsigned long long launches = 0;
  double flops = 0, bytes = 0;
};
ProbeState g_probe;
const size_t kMaxEvents = 1 << 17;
}  // namespace

void probe_open(int cls, cudaStream_t st, double flops, double bytes) {
  ProbeState& p = g_probe;
  if (p.cls != cls || p.used + 2 > kMaxEvents) return;
  while (p.ev.size() < p.used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    p.ev.push_back(e);
  }
  cudaEventRecord(p.ev[p.used], st);
  p.open = true;
  p.flops += flops;
  p.bytes += bytes;
  p.launches += 1;
}

void probe_close(cudaStream_t st) {
  ProbeState& p = g_probe;
  if (!p.open) return;
  cudaEventRecord(p.ev[p.used + 1], st);
  p.used += 2;
  p.open = false;
}

}  // namespace mmqg

using namespace mmqg;

extern "C" {

int mmqg_probe_start(int kernel_class) {
  if (kernel_class < 0 || kernel_class >= KC_COUNT) return set_err(MMQG_ERR_BAD_ARG, "probe: class %d", kernel_class);
  g_probe.cls = kernel_class;
  g_probe.used = 0;
  g_probe.open = false;
  g_probe.launches = 0;
  g_probe.flops = g_probe.bytes = 0;
  return 0;
}

int mmqg_probe_stop(double* total_ms, unsigned long long* launches, double* flops, double* bytes) {
  ProbeState& p = g_probe;
  p.cls = -1;
  double ms = 0;
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    cudaError_t e = cudaEventSynchronize(p.ev[i + 1]);
    if (e != cudaSuccess) return set_err(MMQG_ERR_CUDA, "probe: %s", cudaGetErrorString(e));
    float t = 0;
    e = cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]);
    if (e != cudaSuccess) return set_err(MMQG_ERR_CUDA, "probe: %s", cudaGetErrorString(e));
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = p.used / 2;
  if (flops) *flops = p.flops;
  if (bytes) *bytes = p.bytes;
  p.used = 0;
  return 0;
}

}  // extern "C"

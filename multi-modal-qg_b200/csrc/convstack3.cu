// f2 fast path: the 3x3 / stride-1 convolutions of VideoConvLstmEncoder (reference model/encoder.py:40-49, channels
// 3 -> 4 -> 6 -> 8 -> 10) specialised on (Cin, Cout).  Same contracts as the generic kernels of convstack.cu, which stay
// as the fallback for every other kernel size / stride / channel count.
//
//   forward / input gradient : one thread = 4 consecutive pixels of one row, all output channels.  The layer's weights sit in
//       __constant__ memory and every loop is unrolled, so the weights reach the FFMAs through the uniform datapath (SASS: LDCU
//       into uniform registers, ~1 per 6 FFMAs): no shared-memory traffic and no per-thread registers for weights, and the
//       6-pixel input window of a (channel, row) is loaded once for 12 * Cout FMAs.
//   weight gradient : persistent blocks walk (image, 8-row) tiles; the tile of the (normalised) input and of d z is staged in
//       shared memory ONCE (each tensor is read exactly once from HBM, against Cin times before), warp (ci, ky) keeps its
//       Cout x 3 partial sums in registers over all tiles of the block and the lanes run over the tile's pixels; one
//       shuffle reduction and Cout*Cin*9 atomics per block at the very end.
#include "kernels.h"

namespace mmqg {
namespace cs3 {

__constant__ float c_w[1024];      // weights (Cout, Cin, 3, 3) of the layer being processed
__constant__ float c_b[16];        // its bias

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// v[j] = row[x0 + j] for j < 6 where 0 <= x0 + j < W, else 0.  x0 and W are even, row + x0 is 8-byte aligned.
__device__ __forceinline__ void load6(const float* __restrict__ row, int x0, int W, float (&v)[6]) {
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const int xx = x0 + 2 * p;
    float2 t = make_float2(0.f, 0.f);
    if (xx >= 0 && xx < W) t = __ldg(reinterpret_cast<const float2*>(row + xx));
    v[2 * p] = t.x; v[2 * p + 1] = t.y;
  }
}

// grid (ceil(Ho * ceil(Wo/4) / 256), N).  stats: `parts` slots of 2*COUT floats per image; this block fills slot blockIdx.x and
// block 0 zeroes the slots the generic kernel's larger grid would have written.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv3_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                      float* __restrict__ y, float* __restrict__ stats, int Hin, int Win, int Ho, int Wo, int parts) {
  const int n = blockIdx.y;
  const int Wq = (Wo + 3) >> 2;
  const int t = blockIdx.x * 256 + threadIdx.x;
  const bool live = t < Ho * Wq;
  const int oy = live ? t / Wq : 0, ox = live ? (t - oy * Wq) * 4 : 0;
  float acc[COUT][4];
#pragma unroll
  for (int co = 0; co < COUT; ++co)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[co][p] = 0.f;
  if (live) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float sc = in_scale ? __ldg(in_scale + ci) : 1.f, sh = in_shift ? __ldg(in_shift + ci) : 0.f;
      const float* xp = x + (((size_t)n * CIN + ci) * Hin + oy) * Win;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float xv[6];
        load6(xp + (size_t)ky * Win, ox, Win, xv);
#pragma unroll
        for (int j = 0; j < 6; ++j) xv[j] = fmaf(xv[j], sc, sh);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int co = 0; co < COUT; ++co)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[co][p] = fmaf(xv[p + kx], c_w[((co * CIN + ci) * 3 + ky) * 3 + kx], acc[co][p]);
      }
    }
  }
  __shared__ float red[8][2 * COUT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    float v[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      v[p] = (live && ox + p < Wo) ? fmaxf(acc[co][p] + c_b[co], 0.f) : 0.f;
      s1 += v[p];
      s2 = fmaf(v[p], v[p], s2);
    }
    if (live) {
      float* yp = y + (((size_t)n * COUT + co) * Ho + oy) * Wo + ox;
      if (ox + 1 < Wo) *reinterpret_cast<float2*>(yp) = make_float2(v[0], v[1]);
      if (ox + 3 < Wo) *reinterpret_cast<float2*>(yp + 2) = make_float2(v[2], v[3]);
    }
    if (stats) {
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { red[warp][co] = s1; red[warp][COUT + co] = s2; }
    }
  }
  if (stats) {
    __syncthreads();
    float* sp = stats + (size_t)n * parts * 2 * COUT;
    if (threadIdx.x < 2 * COUT) {
      float tsum = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) tsum += red[wv][threadIdx.x];
      sp[(size_t)blockIdx.x * 2 * COUT + threadIdx.x] = tsum;
    }
    if (blockIdx.x == 0)
      for (int i = gridDim.x * 2 * COUT + threadIdx.x; i < parts * 2 * COUT; i += 256) sp[i] = 0.f;
  }
}

// grid (ceil(Hin * ceil(Win/4) / 256), N): dxn[ci][iy][ix] = sum_{co,ky,kx} dz[co][iy-ky][ix-kx] * w[co][ci][ky][kx]
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv3_bwd_x_kernel(const float* __restrict__ dz, float* __restrict__ dxn, int Hin, int Win, int Ho, int Wo) {
  const int n = blockIdx.y;
  const int Wq = (Win + 3) >> 2;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= Hin * Wq) return;
  const int iy = t / Wq, ix = (t - iy * Wq) * 4;
  float acc[CIN][4];
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[ci][p] = 0.f;
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    const float* gp = dz + ((size_t)n * COUT + co) * Ho * Wo;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oy = iy - ky;
      if (oy < 0 || oy >= Ho) continue;
      float g[6];                                  // d z at columns ix-2 .. ix+3
      load6(gp + (size_t)oy * Wo, ix - 2, Wo, g);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[ci][p] = fmaf(g[p - kx + 2], c_w[((co * CIN + ci) * 3 + ky) * 3 + kx], acc[ci][p]);
    }
  }
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) {
    float* op = dxn + (((size_t)n * CIN + ci) * Hin + iy) * Win + ix;
    if (ix + 1 < Win) *reinterpret_cast<float2*>(op) = make_float2(acc[ci][0], acc[ci][1]);
    if (ix + 3 < Win) *reinterpret_cast<float2*>(op + 2) = make_float2(acc[ci][2], acc[ci][3]);
  }
}

static constexpr int TH = 8;      // output rows per tile of the weight-gradient kernel
static constexpr int kMaxPitch = 112;   // widest input row the weight-gradient kernel stages (the reference's frames are 112 wide)

__device__ __forceinline__ void cp_async8(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// persistent grid; block = CIN*3 warps (ci, ky) + one warp that sums d z for the bias gradient.  Two shared-memory tile
// buffers of (CIN*(TH+2) + COUT*TH) rows of PITCH floats: the raw input rows and the d z rows of tile i+1 arrive by cp.async
// (8-byte, Win even) while tile i is reduced; the BatchNorm of the layer below is applied on the shared-memory read (scale /
// shift are per-warp constants, warp <-> ci).  PITCH (>= Win) is a compile-time row pitch, so every shared-memory access of
// the inner loop is one base register + an immediate offset (the first version spent 3/4 of its issue slots on index arithmetic).
template <int CIN, int COUT, int PITCH>
__global__ void __launch_bounds__(CIN * 3 * 32 + 32)
conv3_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                   const float* __restrict__ dz, float* __restrict__ dw, float* __restrict__ db, int Hin, int Win, int Ho, int Wo,
                   int tiles_per_image, int ntiles) {
  extern __shared__ __align__(16) float sm[];
  constexpr int XROWS = CIN * (TH + 2), GROWS = COUT * TH, BUF = (XROWS + GROWS) * PITCH;
  constexpr int NW = CIN * 3 + 1;                      // warp NW-1 = bias warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool bias_warp = warp == NW - 1;
  const int ci = bias_warp ? 0 : warp / 3, ky = warp - 3 * ci;
  const float sc = in_scale ? __ldg(in_scale + ci) : 1.f, sh = in_shift ? __ldg(in_shift + ci) : 0.f;
  float acc[COUT][3];                                  // bias warp: acc[co][0] = sum of d z
#pragma unroll
  for (int co = 0; co < COUT; ++co)
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[co][k] = 0.f;
  const int wih = Win >> 1, wh = Wo >> 1;
  const size_t xplane = (size_t)Hin * Win, gplane = (size_t)Ho * Wo;
  auto issue = [&](int tile, float* buf) {
    const int n = tile / tiles_per_image, r0 = (tile - n * tiles_per_image) * TH;
    const int th = min(TH, Ho - r0);
    const float* xb = x + (size_t)n * CIN * xplane + (size_t)r0 * Win;
    const float* gb = dz + (size_t)n * COUT * gplane + (size_t)r0 * Wo;
    // smem row pr of the x block = channel pr / (TH+2), row pr % (TH+2) (compile-time divisors); rows beyond th+2 / th of a
    // ragged last tile are skipped (never read)
    for (int pr = warp; pr < XROWS; pr += NW) {
      const int c = pr / (TH + 2), r = pr - c * (TH + 2);
      if (r < th + 2) {
        const float* src = xb + c * xplane + r * Win;
        float* dst = buf + pr * PITCH;
        for (int col = lane; col < wih; col += 32) cp_async8(dst + 2 * col, src + 2 * col);
      }
    }
    for (int pr = warp; pr < GROWS; pr += NW) {
      const int c = pr / TH, r = pr - c * TH;
      if (r < th) {
        const float* src = gb + c * gplane + r * Wo;
        float* dst = buf + (XROWS + pr) * PITCH;
        for (int col = lane; col < wh; col += 32) cp_async8(dst + 2 * col, src + 2 * col);
      }
    }
  };
  // a lane takes two adjacent pixels per step; consecutive steps advance the flattened (row, pixel pair) index by 32
  const int dr = 32 / wh, dq = 32 - dr * wh;
  int cur = 0;
  if ((int)blockIdx.x < ntiles) issue(blockIdx.x, sm);
  cp_async_commit();
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, cur ^= 1) {
    const int next = tile + gridDim.x;
    if (next < ntiles) issue(next, sm + (cur ^ 1) * BUF);
    cp_async_commit();
    cp_async_wait<1>();                                // this thread's copies of `tile` have landed ...
    __syncthreads();                                   // ... and everybody else's
    const float* sx = sm + cur * BUF + (ci * (TH + 2) + ky) * PITCH;
    const float* sg = sm + cur * BUF + XROWS * PITCH;
    const int r0 = (tile % tiles_per_image) * TH;
    const int npair = min(TH, Ho - r0) * wh;
    int row = lane / wh, oq = lane - row * wh;
    if (!bias_warp) {
      for (int idx = lane; idx < npair; idx += 32) {
        const int off = row * PITCH + 2 * oq;
        const float2 xa = *reinterpret_cast<const float2*>(sx + off), xb = *reinterpret_cast<const float2*>(sx + off + 2);
        const float x0 = fmaf(xa.x, sc, sh), x1 = fmaf(xa.y, sc, sh), x2 = fmaf(xb.x, sc, sh), x3 = fmaf(xb.y, sc, sh);
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          const float2 g = *reinterpret_cast<const float2*>(sg + off + co * TH * PITCH);
          acc[co][0] = fmaf(g.y, x1, fmaf(g.x, x0, acc[co][0]));
          acc[co][1] = fmaf(g.y, x2, fmaf(g.x, x1, acc[co][1]));
          acc[co][2] = fmaf(g.y, x3, fmaf(g.x, x2, acc[co][2]));
        }
        row += dr; oq += dq;
        if (oq >= wh) { oq -= wh; ++row; }
      }
    } else if (db) {
      for (int idx = lane; idx < npair; idx += 32) {
        const int off = row * PITCH + 2 * oq;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          const float2 g = *reinterpret_cast<const float2*>(sg + off + co * TH * PITCH);
          acc[co][0] += g.x + g.y;
        }
        row += dr; oq += dq;
        if (oq >= wh) { oq -= wh; ++row; }
      }
    }
    __syncthreads();                                   // buffer `cur` is refilled by the next iteration's issue
  }
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    if (bias_warp) {
      const float s = warp_sum(acc[co][0]);
      if (lane == 0 && db) atomicAdd(db + co, s);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float s = warp_sum(acc[co][k]);
        if (lane == 0) atomicAdd(dw + ((co * CIN + ci) * 3 + ky) * 3 + k, s);
      }
    }
  }
}

static bool shape_ok(int Cin, int Cout, int K, int stride, int Win) {
  if (K != 3 || stride != 1 || (Win & 1) || Win < 4) return false;
  return (Cin == 3 && Cout == 4) || (Cin == 4 && Cout == 6) || (Cin == 6 && Cout == 8) || (Cin == 8 && Cout == 10);
}

static int set_weights(const float* w, const float* b, int Cin, int Cout, cudaStream_t st) {
  MMQG_CUDA(cudaMemcpyToSymbolAsync(c_w, w, sizeof(float) * Cout * Cin * 9, 0, cudaMemcpyDeviceToDevice, st));
  if (b) MMQG_CUDA(cudaMemcpyToSymbolAsync(c_b, b, sizeof(float) * Cout, 0, cudaMemcpyDeviceToDevice, st));
  return 0;
}

template <int CIN, int COUT>
static int fwd_t(const float* x, const float* sc, const float* sh, float* y, float* stats, int N, int Hin, int Win, int parts, cudaStream_t st) {
  const int Ho = Hin - 2, Wo = Win - 2;
  conv3_relu_fwd_kernel<CIN, COUT><<<dim3(ceil_div(Ho * ((Wo + 3) / 4), 256), N), 256, 0, st>>>(x, sc, sh, y, stats, Hin, Win, Ho, Wo, parts);
  MMQG_LAUNCH_CHECK();
  return 0;
}
template <int CIN, int COUT>
static int bwdx_t(const float* dz, float* dxn, int N, int Hin, int Win, cudaStream_t st) {
  conv3_bwd_x_kernel<CIN, COUT><<<dim3(ceil_div(Hin * ((Win + 3) / 4), 256), N), 256, 0, st>>>(dz, dxn, Hin, Win, Hin - 2, Win - 2);
  MMQG_LAUNCH_CHECK();
  return 0;
}
template <int CIN, int COUT, int PITCH>
static int bwdw_p(const float* x, const float* sc, const float* sh, const float* dz, float* dw, float* db, int N, int Hin, int Win,
                  cudaStream_t st) {
  const int Ho = Hin - 2, Wo = Win - 2;
  constexpr size_t smem = 2 * sizeof(float) * (size_t)(CIN * (TH + 2) + COUT * TH) * PITCH;
  static int per_sm = 0;
  if (!per_sm) {
    MMQG_CUDA(cudaFuncSetAttribute(conv3_bwd_w_kernel<CIN, COUT, PITCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMQG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv3_bwd_w_kernel<CIN, COUT, PITCH>, CIN * 96 + 32, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const int tpi = ceil_div(Ho, TH), ntiles = N * tpi;
  const int grid = ntiles < device_sms() * per_sm ? ntiles : device_sms() * per_sm;
  conv3_bwd_w_kernel<CIN, COUT, PITCH><<<grid, CIN * 96 + 32, smem, st>>>(x, sc, sh, dz, dw, db, Hin, Win, Ho, Wo, tpi, ntiles);
  MMQG_LAUNCH_CHECK();
  return 0;
}
template <int CIN, int COUT>
static int bwdw_t(const float* x, const float* sc, const float* sh, const float* dz, float* dw, float* db, int N, int Hin, int Win,
                  cudaStream_t st) {
  if (Win <= 40) return bwdw_p<CIN, COUT, 40>(x, sc, sh, dz, dw, db, N, Hin, Win, st);
  return bwdw_p<CIN, COUT, kMaxPitch>(x, sc, sh, dz, dw, db, N, Hin, Win, st);
}

}  // namespace cs3

#define CS3_DISPATCH(fn, ...)                                              \
  do {                                                                     \
    if (Cin == 3) return cs3::fn<3, 4>(__VA_ARGS__);                       \
    if (Cin == 4) return cs3::fn<4, 6>(__VA_ARGS__);                       \
    if (Cin == 6) return cs3::fn<6, 8>(__VA_ARGS__);                       \
    return cs3::fn<8, 10>(__VA_ARGS__);                                    \
  } while (0)

bool conv3_fast_ok(int Cin, int Cout, int K, int stride, int Win, int max_win) {
  static const bool on = []() { const char* e = getenv("MMQG_CONV_FAST"); return !(e && e[0] == '0'); }();
  return on && cs3::shape_ok(Cin, Cout, K, stride, Win) && (max_win <= 0 || Win <= max_win);
}

int conv3_relu_fwd(const float* x, const float* in_scale, const float* in_shift, const float* w, const float* b, float* y, float* stats,
                   int N, int Cin, int Hin, int Win, int Cout, int parts, cudaStream_t st) {
  MMQG_TRY(cs3::set_weights(w, b, Cin, Cout, st));
  CS3_DISPATCH(fwd_t, x, in_scale, in_shift, y, stats, N, Hin, Win, parts, st);
}

int conv3_bwd_x(const float* dz, const float* w, float* dxn, int N, int Cin, int Hin, int Win, int Cout, cudaStream_t st) {
  MMQG_TRY(cs3::set_weights(w, nullptr, Cin, Cout, st));
  CS3_DISPATCH(bwdx_t, dz, dxn, N, Hin, Win, st);
}

int conv3_bwd_w(const float* x, const float* in_scale, const float* in_shift, const float* dz, float* dw, float* db, int N, int Cin,
                int Hin, int Win, int Cout, cudaStream_t st) {
  CS3_DISPATCH(bwdw_t, x, in_scale, in_shift, dz, dw, db, N, Hin, Win, st);
}

}  // namespace mmqg

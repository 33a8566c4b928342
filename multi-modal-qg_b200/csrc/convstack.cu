// f2 ("next"): the conv stack of VideoConvLstmEncoder on raw frames, forward and backward
// (reference model/encoder.py:40-52 layers, :58-67 forward):  4 x [Conv2d(k, stride) -> ReLU -> BatchNorm2d] with a
// MaxPool2d(k, k) behind the second and the fourth, channels 3 -> 4 -> 6 -> 8 -> 10, 112x112 frames -> 10x10x10 features.
// Channel counts are tiny (<= 16), so nothing here is GEMM-shaped: every kernel is a coalesced pass over the
// activation tensors (HBM/L2-bound), with the small weight set in shared memory.  Train-mode BatchNorm needs the batch
// statistics of its input before it can normalise, so each BatchNorm is split in two and folded into its neighbours:
//   conv_relu_fwd   : y = relu(conv(x * in_scale + in_shift) + b)       (the PREVIOUS BatchNorm applied on the load)
//                     + per-block partial sums / sums of squares of y   (statistics of the NEXT BatchNorm)
//   bn_finalize     : statistics -> scale / shift, saved mean / invstd, running-stat update (BatchNorm2d, momentum 0.1)
//   bn_maxpool_fwd  : out = maxpool(y * scale + shift), arg-max kept for the backward pass (first maximum on ties)
// Backward, per layer: maxpool_bwd (dense gradient w.r.t. the BatchNorm output) -> bn_bwd_stats (sum d, sum d*xhat)
//   -> bn_relu_bwd (d conv output) -> conv_bwd_w (dW, db) and conv_bwd_x (gradient w.r.t. the normalised input).
// NCHW fp32 throughout (the reference's layout); the `view` that scrambles channels and time is the
// caller's business: these kernels see (N, C, H, W).
#include "kernels.h"

namespace mmqg {
namespace cs {

static constexpr int MAXC = 16;          // channels per tensor
static constexpr int MAXW = 4096;        // weights of one conv layer kept in shared memory (Cout * Cin * K * K)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid (ceil(Ho*Wo / 256), N).  Thread = one output pixel, all output channels.
__global__ void __launch_bounds__(256)
conv_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                     const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y, float* __restrict__ stats,
                     int Cin, int Hin, int Win, int Cout, int K, int stride, int Ho, int Wo) {
  __shared__ float sw[MAXW];
  __shared__ float ssc[MAXC], ssh[MAXC], sb[MAXC];
  const int nw = Cout * Cin * K * K;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < Cin) {
    ssc[threadIdx.x] = in_scale ? in_scale[threadIdx.x] : 1.f;
    ssh[threadIdx.x] = in_shift ? in_shift[threadIdx.x] : 0.f;
  }
  if (threadIdx.x < Cout) sb[threadIdx.x] = b ? b[threadIdx.x] : 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < Ho * Wo;
  const int oy = live ? p / Wo : 0, ox = live ? p % Wo : 0;
  float acc[MAXC];
#pragma unroll
  for (int co = 0; co < MAXC; ++co) acc[co] = 0.f;
  if (live) {
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xp = x + (((size_t)n * Cin + ci) * Hin + (size_t)oy * stride) * Win + (size_t)ox * stride;
      const float sc = ssc[ci], sh = ssh[ci];
      for (int ky = 0; ky < K; ++ky)
        for (int kx = 0; kx < K; ++kx) {
          const float xv = fmaf(xp[(size_t)ky * Win + kx], sc, sh);
          const float* wp = sw + (ci * K + ky) * K + kx;
#pragma unroll
          for (int co = 0; co < MAXC; ++co)
            if (co < Cout) acc[co] = fmaf(xv, wp[co * Cin * K * K], acc[co]);
        }
    }
  }
  // per-block partial statistics (no same-address atomics: N * tiles blocks would serialise on 2*Cout words of L2)
  __shared__ float red[8][2 * MAXC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int co = 0; co < MAXC; ++co) {
    if (co >= Cout) break;
    const float v = live ? fmaxf(acc[co] + sb[co], 0.f) : 0.f;
    if (live) y[(((size_t)n * Cout + co) * Ho + oy) * Wo + ox] = v;
    if (stats) {
      const float s1 = warp_sum(v), s2 = warp_sum(v * v);
      if (lane == 0) { red[warp][co] = s1; red[warp][Cout + co] = s2; }
    }
  }
  if (stats) {
    __syncthreads();
    if (threadIdx.x < 2 * Cout) {
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
      stats[((size_t)n * gridDim.x + blockIdx.x) * 2 * Cout + threadIdx.x] = t;
    }
  }
}

// Partial statistics (nparts, 2C) -> acc[2C] (zeroed by the caller): rows are read coalesced, thread t owns column t % 2C.
__global__ void __launch_bounds__(256)
bn_reduce_parts_kernel(const float* __restrict__ parts, int nparts, int C2, float* __restrict__ acc) {
  __shared__ float red[256];
  const int rows_per_pass = 256 / C2;
  const int col = threadIdx.x % C2, rl = threadIdx.x / C2;
  float s = 0.f;
  if (rl < rows_per_pass)
    for (int r = blockIdx.x * rows_per_pass + rl; r < nparts; r += gridDim.x * rows_per_pass) s += parts[(size_t)r * C2 + col];
  red[threadIdx.x] = rl < rows_per_pass ? s : 0.f;
  __syncthreads();
  if (threadIdx.x < C2) {
    float t = 0.f;
    for (int k = 0; k < rows_per_pass; ++k) t += red[k * C2 + threadIdx.x];
    atomicAdd(acc + threadIdx.x, t);
  }
}

// one thread per channel: acc = (sum | sum of squares) -> the BatchNorm bookkeeping.  acc may alias mean_out / invstd_out storage.
__global__ void bn_finalize_kernel(const float* __restrict__ acc, float count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out, int C) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const float s1 = acc[c], s2 = acc[C + c];
  const float mean = s1 / count;
  const float var = fmaxf(s2 / count - mean * mean, 0.f);     // biased, as BatchNorm normalises with
  const float invstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = bt - mean * g * invstd;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (count > 1.f ? count / (count - 1.f) : 1.f);
}

// thread = one pooled output element; window K x K, stride K, floor mode
__global__ void bn_maxpool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                      float* __restrict__ out, unsigned char* __restrict__ idx, int C, int H, int W, int K, int Hp,
                                      int Wp, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int px = (int)(i % Wp), py = (int)((i / Wp) % Hp);
  const long long nc = i / ((long long)Wp * Hp);
  const int c = (int)(nc % C);
  const float sc = scale[c], sh = shift[c];
  const float* yp = y + (nc * H + (long long)py * K) * W + (long long)px * K;
  float best = -INFINITY;
  int bi = 0;
  for (int ky = 0; ky < K; ++ky)
    for (int kx = 0; kx < K; ++kx) {
      const float v = fmaf(yp[(long long)ky * W + kx], sc, sh);
      if (v > best || v != v) { best = v; bi = ky * K + kx; }      // first maximum on ties, NaN propagates (as torch)
    }
  out[i] = best;
  idx[i] = (unsigned char)bi;
}

// dense gradient w.r.t. the BatchNorm output from the pooled gradient: thread = one element of the (N,C,H,W) tensor
__global__ void maxpool_bwd_kernel(const float* __restrict__ dpool, const unsigned char* __restrict__ idx, float* __restrict__ dbn,
                                   int H, int W, int K, int Hp, int Wp, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xw = (int)(i % W), yh = (int)((i / W) % H);
  const long long nc = i / ((long long)W * H);
  const int py = yh / K, px = xw / K;
  float g = 0.f;
  if (py < Hp && px < Wp) {
    const long long o = (nc * Hp + py) * Wp + px;
    if (idx[o] == (unsigned char)((yh - py * K) * K + (xw - px * K))) g = dpool[o];
  }
  dbn[i] = g;
}

// Gradient w.r.t. the BatchNorm output when a MaxPool2d(K, K) follows: the pooled gradient of the window that holds the
// element if the element was the window's (first) maximum, else 0 -- maxpool_bwd_kernel on the fly, so the dense (N,C,H,W)
// gradient is never written or re-read.  KC = 3: compile-time window (the reference's); KC = 0: pg.K at run time.
struct PoolGrad {
  const float* dpool; const unsigned char* idx; int W, K, Hp, Wp;
};
template <int KC>
__device__ __forceinline__ float pool_grad(const PoolGrad& pg, long long nc, int yh, int xw) {
  const int K = KC ? KC : pg.K;
  const int py = yh / K, px = xw / K;
  if (py >= pg.Hp || px >= pg.Wp) return 0.f;
  const long long o = (nc * pg.Hp + py) * pg.Wp + px;
  return pg.idx[o] == (unsigned char)((yh - py * K) * K + (xw - px * K)) ? pg.dpool[o] : 0.f;
}

// Four consecutive elements i0 .. i0+3 of plane nc: y values and the gradient w.r.t. the BatchNorm output.
// MODE -1: dense gradient dbn; 0 / 3: pooled gradient (run-time / compile-time window).  vec: float4 accesses are legal.
template <int MODE>
__device__ __forceinline__ void load_y_d(const float* __restrict__ yp, const float* __restrict__ dp, const PoolGrad& pg, long long nc,
                                         int i0, int HW, bool vec, float (&yv)[4], float (&d)[4]) {
  if (vec) {
    const float4 t = *reinterpret_cast<const float4*>(yp + i0);
    yv[0] = t.x; yv[1] = t.y; yv[2] = t.z; yv[3] = t.w;
    if (MODE < 0) {
      const float4 u = *reinterpret_cast<const float4*>(dp + i0);
      d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      yv[q] = i0 + q < HW ? yp[i0 + q] : 0.f;
      if (MODE < 0) d[q] = i0 + q < HW ? dp[i0 + q] : 0.f;
    }
  }
  if (MODE >= 0) {
    int yh = i0 / pg.W, xw = i0 - yh * pg.W;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      d[q] = i0 + q < HW ? pool_grad<(MODE > 0 ? MODE : 0)>(pg, nc, yh, xw) : 0.f;
      if (++xw == pg.W) { xw = 0; ++yh; }
    }
  }
}

// sums[c] += sum d, sums[C + c] += sum d * xhat over (n, h, w); grid (N*C, blocks per plane)
template <int MODE>
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ dbn, PoolGrad pg, float* __restrict__ sums, int C, int HW, bool vec) {
  __shared__ float r1[8], r2[8];
  const long long nc = blockIdx.x;
  const int c = (int)(nc % C);
  const float m = mean[c], is = invstd[c];
  const float* yp = y + nc * HW;
  const float* dp = MODE < 0 ? dbn + nc * HW : nullptr;
  float s1 = 0.f, s2 = 0.f;
  for (int i0 = (blockIdx.y * 256 + threadIdx.x) * 4; i0 < HW; i0 += gridDim.y * 1024) {
    float yv[4], d[4];
    load_y_d<MODE>(yp, dp, pg, nc, i0, HW, vec, yv, d);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s1 += d[q];
      s2 = fmaf(d[q], (yv[q] - m) * is, s2);
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { r1[warp] = s1; r2[warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    s1 = lane < 8 ? r1[lane] : 0.f; s2 = lane < 8 ? r2[lane] : 0.f;
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { atomicAdd(sums + c, s1); atomicAdd(sums + C + c, s2); }
  }
}

// dz = gamma * invstd * (d - sum_d / M - xhat * sum_dxhat / M) * (y > 0)     (sums == NULL: eval-mode BatchNorm, dz = d * scale * (y > 0))
// grid (N*C, ceil(HW / 1024)): four consecutive elements per thread
template <int MODE>
__global__ void __launch_bounds__(256)
bn_relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ invstd,
                   const float* __restrict__ gamma, const float* __restrict__ sums, const float* __restrict__ dbn, PoolGrad pg,
                   float* __restrict__ dz, int C, int HW, float inv_count, bool vec) {
  const int i0 = (blockIdx.y * 256 + threadIdx.x) * 4;
  if (i0 >= HW) return;
  const long long nc = blockIdx.x;
  const int c = (int)(nc % C);
  const float is = invstd[c], gi = (gamma ? gamma[c] : 1.f) * is;
  const float m = sums ? mean[c] : 0.f, a = sums ? sums[c] * inv_count : 0.f, b = sums ? sums[C + c] * inv_count * is : 0.f;
  float yv[4], d[4], v[4];
  load_y_d<MODE>(y + nc * HW, MODE < 0 ? dbn + nc * HW : nullptr, pg, nc, i0, HW, vec, yv, d);
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = yv[q] > 0.f ? gi * (d[q] - a - (yv[q] - m) * b) : 0.f;
  float* op = dz + nc * HW + i0;
  if (vec) {
    *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i0 + q < HW) op[q] = v[q];
  }
}

// grid (Cout * Cin, chunks).  Block = one (co, ci) pair over a chunk of the N*Ho*Wo positions; thread accumulates the K*K taps.
__global__ void __launch_bounds__(256)
conv_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                  const float* __restrict__ dz, float* __restrict__ dw, float* __restrict__ db, int N, int Cin, int Hin, int Win,
                  int Cout, int K, int stride, int Ho, int Wo) {
  __shared__ float red[8][26];
  const int co = blockIdx.x / Cin, ci = blockIdx.x % Cin;
  const float sc = in_scale ? in_scale[ci] : 1.f, sh = in_shift ? in_shift[ci] : 0.f;
  const long long P = (long long)N * Ho * Wo;
  float acc[25], accb = 0.f;
#pragma unroll
  for (int k = 0; k < 25; ++k) acc[k] = 0.f;
  for (long long p = (long long)blockIdx.y * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.y * blockDim.x) {
    const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
    const float g = dz[(((size_t)n * Cout + co) * Ho + oy) * Wo + ox];
    accb += g;
    const float* xp = x + (((size_t)n * Cin + ci) * Hin + (size_t)oy * stride) * Win + (size_t)ox * stride;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx)
        if (ky < K && kx < K) acc[ky * 5 + kx] = fmaf(g, fmaf(xp[(size_t)ky * Win + kx], sc, sh), acc[ky * 5 + kx]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 25; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = s;
  }
  {
    const float s = warp_sum(accb);
    if (lane == 0) red[warp][25] = s;
  }
  __syncthreads();
  if (threadIdx.x < 26) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    if (threadIdx.x == 25) {
      if (ci == 0 && db) atomicAdd(db + co, s);
    } else {
      const int ky = threadIdx.x / 5, kx = threadIdx.x % 5;
      if (ky < K && kx < K) atomicAdd(dw + ((size_t)(co * Cin + ci) * K + ky) * K + kx, s);
    }
  }
}

// K = 3 fast path.  grid (chunks, Cin): a block takes ONE input channel and a slice of the N*Ho*Wo positions; a thread keeps the
// 9 taps x Cout partial sums in registers (x patch read once per position, d z once per output channel), so d z is read Cin
// times in total instead of Cin * Cout * 9 / ... times through the generic kernel's (co, ci) blocks.
template <int STRIDE1>
__global__ void __launch_bounds__(256)
conv_bwd_w3_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                   const float* __restrict__ dz, float* __restrict__ dw, float* __restrict__ db, int N, int Cin, int Hin, int Win,
                   int Cout, int stride, int Ho, int Wo) {
  __shared__ float red[8][MAXC * 9 + MAXC];
  const int ci = blockIdx.y;
  const float sc = in_scale ? in_scale[ci] : 1.f, sh = in_shift ? in_shift[ci] : 0.f;
  const long long P = (long long)N * Ho * Wo;
  const size_t plane = (size_t)Ho * Wo;
  float acc[MAXC][9], accb[MAXC];
#pragma unroll
  for (int co = 0; co < MAXC; ++co) {
    accb[co] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[co][k] = 0.f;
  }
  const int st = STRIDE1 ? 1 : stride;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
    const float* xp = x + (((size_t)n * Cin + ci) * Hin + (size_t)oy * st) * Win + (size_t)ox * st;
    float xv[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) xv[ky * 3 + kx] = fmaf(xp[(size_t)ky * Win + kx], sc, sh);
    const float* gp = dz + (size_t)n * Cout * plane + (size_t)oy * Wo + ox;
#pragma unroll
    for (int co = 0; co < MAXC; ++co) {
      if (co >= Cout) break;
      const float g = gp[co * plane];
      accb[co] += g;
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[co][k] = fmaf(g, xv[k], acc[co][k]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int co = 0; co < MAXC; ++co) {
    if (co >= Cout) break;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float s = warp_sum(acc[co][k]);
      if (lane == 0) red[warp][co * 9 + k] = s;
    }
    const float sb = warp_sum(accb[co]);
    if (lane == 0) red[warp][MAXC * 9 + co] = sb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 9 + Cout; i += blockDim.x) {
    const bool is_b = i >= Cout * 9;
    const int slot = is_b ? MAXC * 9 + (i - Cout * 9) : i;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][slot];
    if (is_b) {
      if (ci == 0 && db) atomicAdd(db + (i - Cout * 9), s);
    } else {
      const int co = i / 9, k = i % 9;
      atomicAdd(dw + ((size_t)(co * Cin + ci)) * 9 + k, s);
    }
  }
}

// grid (ceil(Hin*Win / 256), N).  Thread = one input pixel, all input channels: dxn = sum_{co,ky,kx} dz[oy, ox] * w[co][ci][ky][kx]
template <int STRIDE1>
__global__ void __launch_bounds__(256)
conv_bwd_x_kernel(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ dxn, int Cin, int Hin, int Win,
                  int Cout, int K, int stride, int Ho, int Wo) {
  __shared__ float sw[MAXW];
  const int nw = Cout * Cin * K * K;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Hin * Win) return;
  const int iy = p / Win, ix = p % Win;
  float acc[MAXC];
#pragma unroll
  for (int ci = 0; ci < MAXC; ++ci) acc[ci] = 0.f;
  for (int ky = 0; ky < K; ++ky) {
    int oy = iy - ky;
    if (oy < 0) continue;
    if (!STRIDE1) {
      if (oy % stride) continue;
      oy /= stride;
    }
    if (oy >= Ho) continue;
    for (int kx = 0; kx < K; ++kx) {
      int ox = ix - kx;
      if (ox < 0) continue;
      if (!STRIDE1) {
        if (ox % stride) continue;
        ox /= stride;
      }
      if (ox >= Wo) continue;
      const float* gp = dz + ((size_t)n * Cout * Ho + oy) * Wo + ox;
      const float* wp = sw + ky * K + kx;
      for (int co = 0; co < Cout; ++co) {
        const float g = gp[(size_t)co * Ho * Wo];
#pragma unroll
        for (int ci = 0; ci < MAXC; ++ci)
          if (ci < Cin) acc[ci] = fmaf(g, wp[(co * Cin + ci) * K * K], acc[ci]);
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < MAXC; ++ci)
    if (ci < Cin) dxn[(((size_t)n * Cin + ci) * Hin + iy) * Win + ix] = acc[ci];
}

static int conv_out(int in, int K, int stride) { return (in - K) / stride + 1; }

}  // namespace cs
}  // namespace mmqg

using namespace mmqg;

// stats: (mmqg_conv_stats_parts(...) x 2*Cout) floats of per-block partial sums (sum | sum of squares), or NULL
extern "C" int mmqg_conv_stats_parts(int N, int Hin, int Win, int K, int stride) {
  if (N <= 0 || Hin < K || Win < K || K < 1 || stride < 1) return 0;
  return N * ceil_div(cs::conv_out(Hin, K, stride) * cs::conv_out(Win, K, stride), 256);
}

extern "C" int mmqg_conv_relu_fwd(const float* x, const float* in_scale, const float* in_shift, const float* w, const float* b, float* y,
                                  float* stats, int N, int Cin, int Hin, int Win, int Cout, int K, int stride, void* stream) {
  MMQG_REQUIRE(x && w && y && N > 0 && Cin > 0 && Cin <= cs::MAXC && Cout > 0 && Cout <= cs::MAXC && K >= 1 && K <= 5 && stride >= 1 &&
                   Hin >= K && Win >= K && Cout * Cin * K * K <= cs::MAXW,
               "conv_relu_fwd: unsupported shape N=%d Cin=%d Cout=%d K=%d stride=%d H=%d W=%d", N, Cin, Cout, K, stride, Hin, Win);
  cudaStream_t st = as_stream(stream);
  const int Ho = cs::conv_out(Hin, K, stride), Wo = cs::conv_out(Win, K, stride);
  MMQG_PROBE(KC_OTHER, 2.0 * N * Ho * Wo * Cout * Cin * K * K, 4.0 * N * ((double)Cin * Hin * Win + (double)Cout * Ho * Wo));
  if (b && conv3_fast_ok(Cin, Cout, K, stride, Win, 0) && ((uintptr_t)x % 8 == 0) && ((uintptr_t)y % 8 == 0))
    return conv3_relu_fwd(x, in_scale, in_shift, w, b, y, stats, N, Cin, Hin, Win, Cout, mmqg_conv_stats_parts(N, Hin, Win, K, stride) / N, st);
  cs::conv_relu_fwd_kernel<<<dim3(ceil_div(Ho * Wo, 256), N), 256, 0, st>>>(x, in_scale, in_shift, w, b, y, stats, Cin, Hin, Win, Cout, K,
                                                                            stride, Ho, Wo);
  MMQG_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmqg_bn_finalize(const float* stats_parts, int nparts, long long count, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                                float* invstd, int C, void* stream) {
  MMQG_REQUIRE(stats_parts && nparts > 0 && scale && shift && mean && invstd && C > 0 && C <= cs::MAXC && count > 0, "bn_finalize: bad args");
  cudaStream_t st = as_stream(stream);
  // the sums are accumulated in `scale` | `shift` (2C floats when contiguous, else in two steps) -- the outputs double as scratch
  float* acc = scale;
  const bool contig = shift == scale + C;
  if (contig && nparts > 1) {
    MMQG_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * 2 * C, st));
    int blocks = ceil_div(nparts, (256 / (2 * C)) * 8);
    blocks = blocks > 296 ? 296 : blocks;
    cs::bn_reduce_parts_kernel<<<blocks, 256, 0, st>>>(stats_parts, nparts, 2 * C, acc);
    MMQG_LAUNCH_CHECK();
    cs::bn_finalize_kernel<<<1, 32, 0, st>>>(acc, (float)count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift,
                                             mean, invstd, C);
  } else if (nparts == 1) {
    cs::bn_finalize_kernel<<<1, 32, 0, st>>>(stats_parts, (float)count, gamma, beta, eps, momentum, running_mean, running_var, scale,
                                             shift, mean, invstd, C);
  } else {
    return set_err(MMQG_ERR_BAD_ARG, "bn_finalize: scale and shift must be one contiguous (2, C) buffer when nparts > 1");
  }
  MMQG_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmqg_bn_maxpool_fwd(const float* y, const float* scale, const float* shift, float* out, unsigned char* idx, int N, int C,
                                   int H, int W, int K, void* stream) {
  MMQG_REQUIRE(y && scale && shift && out && idx && N > 0 && C > 0 && C <= cs::MAXC && K >= 1 && K <= 15 && H >= K && W >= K,
               "bn_maxpool_fwd: bad args");
  cudaStream_t st = as_stream(stream);
  const int Hp = (H - K) / K + 1, Wp = (W - K) / K + 1;
  const long long total = (long long)N * C * Hp * Wp;
  MMQG_PROBE(KC_OTHER, 0, 4.0 * N * C * ((double)H * W + (double)Hp * Wp));
  cs::bn_maxpool_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(y, scale, shift, out, idx, C, H, W, K, Hp, Wp, total);
  MMQG_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmqg_maxpool_bwd(const float* dpool, const unsigned char* idx, float* dbn, int N, int C, int H, int W, int K, void* stream) {
  MMQG_REQUIRE(dpool && idx && dbn && N > 0 && C > 0 && K >= 1 && H >= K && W >= K, "maxpool_bwd: bad args");
  cudaStream_t st = as_stream(stream);
  const int Hp = (H - K) / K + 1, Wp = (W - K) / K + 1;
  const long long total = (long long)N * C * H * W;
  MMQG_PROBE(KC_OTHER, 0, 4.0 * total);
  cs::maxpool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dpool, idx, dbn, H, W, K, Hp, Wp, total);
  MMQG_LAUNCH_CHECK();
  return 0;
}

// Backward of train-mode BatchNorm + ReLU: y = relu output (the BatchNorm's input), dbn = gradient w.r.t. the BatchNorm output,
// dz = gradient w.r.t. the conv output (may alias dbn).  sums (2*C floats: sum d = d beta, sum d*xhat = d gamma) is filled here.
// sums == NULL: eval-mode BatchNorm (fixed affine map, invstd = 1/sqrt(running_var + eps)).
template <int MODE>
static int bn_relu_bwd_impl(const float* y, const float* mean, const float* invstd, const float* gamma, const float* dbn, cs::PoolGrad pg,
                            float* dz, float* sums, int N, int C, int H, int W, cudaStream_t st) {
  const int HW = H * W;
  const long long total = (long long)N * C * HW;
  const bool vec = HW % 4 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)dz % 16 == 0 && (MODE >= 0 || (uintptr_t)dbn % 16 == 0);
  if (sums) {
    MMQG_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, st));
    int by = ceil_div(HW, 1024 * 4);
    if (by > 16) by = 16;
    MMQG_PROBE(KC_OTHER, 0, (MODE >= 0 ? 4.0 : 8.0) * total);
    cs::bn_bwd_stats_kernel<MODE><<<dim3(N * C, by), 256, 0, st>>>(y, mean, invstd, dbn, pg, sums, C, HW, vec);
    MMQG_LAUNCH_CHECK();
  }
  MMQG_PROBE(KC_OTHER, 0, (MODE >= 0 ? 8.0 : 12.0) * total);
  cs::bn_relu_bwd_kernel<MODE><<<dim3(N * C, ceil_div(HW, 1024)), 256, 0, st>>>(y, mean, invstd, gamma, sums, dbn, pg, dz, C, HW,
                                                                                1.0f / ((float)N * HW), vec);
  MMQG_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmqg_bn_relu_bwd(const float* y, const float* mean, const float* invstd, const float* gamma, const float* dbn, float* dz,
                                float* sums, int N, int C, int H, int W, void* stream) {
  MMQG_REQUIRE(y && invstd && dbn && dz && N > 0 && C > 0 && C <= cs::MAXC && H > 0 && W > 0 && (!sums || mean), "bn_relu_bwd: bad args");
  return bn_relu_bwd_impl<-1>(y, mean, invstd, gamma, dbn, cs::PoolGrad{}, dz, sums, N, C, H, W, as_stream(stream));
}

// The same with a MaxPool2d(K, K) between the BatchNorm and the incoming gradient: dpool / idx are the pooled gradient and the
// arg-max saved by mmqg_bn_maxpool_fwd; equals mmqg_maxpool_bwd followed by mmqg_bn_relu_bwd without the dense intermediate.
extern "C" int mmqg_bn_relu_pool_bwd(const float* y, const float* mean, const float* invstd, const float* gamma, const float* dpool,
                                     const unsigned char* idx, int K, float* dz, float* sums, int N, int C, int H, int W, void* stream) {
  MMQG_REQUIRE(y && invstd && dpool && idx && dz && N > 0 && C > 0 && C <= cs::MAXC && K >= 1 && K <= 15 && H >= K && W >= K &&
                   (!sums || mean), "bn_relu_pool_bwd: bad args");
  cs::PoolGrad pg{dpool, idx, W, K, (H - K) / K + 1, (W - K) / K + 1};
  if (K == 3) return bn_relu_bwd_impl<3>(y, mean, invstd, gamma, nullptr, pg, dz, sums, N, C, H, W, as_stream(stream));
  return bn_relu_bwd_impl<0>(y, mean, invstd, gamma, nullptr, pg, dz, sums, N, C, H, W, as_stream(stream));
}

extern "C" int mmqg_conv_bwd_w(const float* x, const float* in_scale, const float* in_shift, const float* dz, float* dw, float* db, int N,
                               int Cin, int Hin, int Win, int Cout, int K, int stride, void* stream) {
  MMQG_REQUIRE(x && dz && dw && N > 0 && Cin > 0 && Cin <= cs::MAXC && Cout > 0 && Cout <= cs::MAXC && K >= 1 && K <= 5 && stride >= 1 &&
                   Hin >= K && Win >= K, "conv_bwd_w: unsupported shape");
  cudaStream_t st = as_stream(stream);
  const int Ho = cs::conv_out(Hin, K, stride), Wo = cs::conv_out(Win, K, stride);
  MMQG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * K * K, st));
  if (db) MMQG_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cout, st));
  const long long P = (long long)N * Ho * Wo;
  int chunks = (int)((P + 256 * 16 - 1) / (256 * 16));
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  if (conv3_fast_ok(Cin, Cout, K, stride, Win, 112) && ((uintptr_t)x % 8 == 0) && ((uintptr_t)dz % 8 == 0)) {
    MMQG_PROBE(KC_OTHER, 2.0 * P * Cout * Cin * K * K, 4.0 * (N * (double)Cin * Hin * Win + (double)P * Cout));
    return conv3_bwd_w(x, in_scale, in_shift, dz, dw, db, N, Cin, Hin, Win, Cout, st);
  }
  MMQG_PROBE(KC_OTHER, 2.0 * P * Cout * Cin * K * K, 4.0 * (N * (double)Cin * Hin * Win + (double)Cin * P * Cout));
  if (K == 3) {
    int cx = (int)((P + 256 * 32 - 1) / (256 * 32));
    if (cx < 1) cx = 1;
    if (cx > 148 * 2) cx = 148 * 2;
    if (stride == 1) cs::conv_bwd_w3_kernel<1><<<dim3(cx, Cin), 256, 0, st>>>(x, in_scale, in_shift, dz, dw, db, N, Cin, Hin, Win, Cout, stride, Ho, Wo);
    else cs::conv_bwd_w3_kernel<0><<<dim3(cx, Cin), 256, 0, st>>>(x, in_scale, in_shift, dz, dw, db, N, Cin, Hin, Win, Cout, stride, Ho, Wo);
  } else {
    cs::conv_bwd_w_kernel<<<dim3(Cout * Cin, chunks), 256, 0, st>>>(x, in_scale, in_shift, dz, dw, db, N, Cin, Hin, Win, Cout, K, stride, Ho, Wo);
  }
  MMQG_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmqg_conv_bwd_x(const float* dz, const float* w, float* dxn, int N, int Cin, int Hin, int Win, int Cout, int K, int stride,
                               void* stream) {
  MMQG_REQUIRE(dz && w && dxn && N > 0 && Cin > 0 && Cin <= cs::MAXC && Cout > 0 && Cout <= cs::MAXC && K >= 1 && K <= 5 && stride >= 1 &&
                   Hin >= K && Win >= K && Cout * Cin * K * K <= cs::MAXW, "conv_bwd_x: unsupported shape");
  cudaStream_t st = as_stream(stream);
  const int Ho = cs::conv_out(Hin, K, stride), Wo = cs::conv_out(Win, K, stride);
  MMQG_PROBE(KC_OTHER, 2.0 * N * Ho * Wo * Cout * Cin * K * K, 4.0 * N * ((double)Cin * Hin * Win + (double)Cout * Ho * Wo));
  if (conv3_fast_ok(Cin, Cout, K, stride, Win, 0) && ((uintptr_t)dz % 8 == 0) && ((uintptr_t)dxn % 8 == 0))
    return conv3_bwd_x(dz, w, dxn, N, Cin, Hin, Win, Cout, st);
  if (stride == 1) cs::conv_bwd_x_kernel<1><<<dim3(ceil_div(Hin * Win, 256), N), 256, 0, st>>>(dz, w, dxn, Cin, Hin, Win, Cout, K, stride, Ho, Wo);
  else cs::conv_bwd_x_kernel<0><<<dim3(ceil_div(Hin * Win, 256), N), 256, 0, st>>>(dz, w, dxn, Cin, Hin, Win, Cout, K, stride, Ho, Wo);
  MMQG_LAUNCH_CHECK();
  return 0;
}

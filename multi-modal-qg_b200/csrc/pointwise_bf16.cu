// Elementwise / row-wise kernels of the bf16 mode: same arithmetic as pointwise.cu (fp32 math,
// fp32 cell state and reductions) but the tensors that feed tensor-core GEMMs are written as
// bf16 (h sequences, gate gradients, gathered embeddings, dlogits, packed weights).
#include <cuda_bf16.h>
#include "kernels.h"
#include "dropout.cuh"

namespace mmqg {

typedef __nv_bfloat16 bf16;

// dst(r, c) = bf16(src(r, c)) for c < cols, 0 for cols <= c < cols_dst  (weight packing, input conversion)
__global__ void cvt_f32_bf16_2d_kernel(const float* __restrict__ src, long long ld_src, bf16* __restrict__ dst,
                                       long long ld_dst, long long rows, int cols, int cols_dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = rows * cols_dst;
  if (i >= total) return;
  long long r = i / cols_dst;
  int c = (int)(i % cols_dst);
  float v = c < cols ? src[r * ld_src + c] : 0.f;
  dst[r * ld_dst + c] = __float2bfloat16_rn(v);
}

__global__ void embedding_gather_bf16_kernel(const float* __restrict__ emb, const int64_t* __restrict__ idx,
                                             bf16* __restrict__ out, int ldo, int N, int E, int E_pad, int V) {
  int n = blockIdx.x;
  long long w = idx[n];
  w = w < 0 ? 0 : (w >= V ? V - 1 : w);
  const float* src = emb + (size_t)w * E;
  bf16* dst = out + (size_t)n * ldo;
  for (int e = threadIdx.x; e < E_pad; e += blockDim.x) dst[e] = __float2bfloat16_rn(e < E ? src[e] : 0.f);
}

// ---- inter-layer dropout (torch.nn.LSTM(dropout=p): outputs of every layer but the last) ----
// Counter-based mask: element `idx` of logical stream `sid` is kept iff u(seed, sid, idx) >= p, so
// the forward (y = x * m / (1-p)) and the backward (dx *= m / (1-p)) regenerate the same mask from
// the seed and chunked application equals whole-tensor application.  Not bit-compatible with
// ATen's Philox stream (SURVEY section 7 "Hard parts"): validated against the oracle run with the
// exported mask (mmqg_dropout_mask) and statistically.

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void lstm_pointwise_fwd_bf16_kernel(float* __restrict__ gates, int ldg, const float* __restrict__ c_prev,
                                               int ldcp, float* __restrict__ c_out, int ldc, bf16* __restrict__ h_out,
                                               int ldh, float* __restrict__ h2, int ldh2, int B, int H, DropSpec dr,
                                               PreSpec ps) {
  pdl_launch_dependents();
  pdl_wait();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  int b = idx / H, j = idx % H;
  float* g = gates + (size_t)b * ldg;
  float pi, pf, pg, po;
  if (ps.part) {       // pre-activations arrive as split-K partial sums (+ bias, + what `gates` already holds)
    pi = pf = pg = po = 0.f;
    if (ps.add_gates) { pi = g[j]; pf = g[H + j]; pg = g[2 * H + j]; po = g[3 * H + j]; }
    if (ps.bias) { pi += ps.bias[j]; pf += ps.bias[H + j]; pg += ps.bias[2 * H + j]; po += ps.bias[3 * H + j]; }
    for (int k = 0; k < ps.n_part; ++k) {
      const float* q = ps.part + (size_t)k * ps.stride + (size_t)b * ps.ld;
      pi += q[j]; pf += q[H + j]; pg += q[2 * H + j]; po += q[3 * H + j];
    }
  } else {
    pi = g[j]; pf = g[H + j]; pg = g[2 * H + j]; po = g[3 * H + j];
  }
  float i = sigm(pi), f = sigm(pf), gg = tanhf(pg), o = sigm(po);
  float cp = c_prev ? c_prev[(size_t)b * ldcp + j] : 0.f;
  float c = f * cp + i * gg;
  float h = o * tanhf(c);
  g[j] = i; g[H + j] = f; g[2 * H + j] = gg; g[3 * H + j] = o;
  c_out[(size_t)b * ldc + j] = c;
  h_out[(size_t)b * ldh + j] = __float2bfloat16_rn(h);
  if (h2) h2[(size_t)b * ldh2 + j] = h;
  if (dr.out)   // dropped copy: the input of the next layer
    reinterpret_cast<bf16*>(dr.out)[(size_t)b * dr.ld + j] =
        __float2bfloat16_rn(h * drop_scale(dr.seed + (dr.ctr ? *dr.ctr : 0ull), dr.sid, dr.base + (unsigned long long)idx, dr.p, 1.0f / (1.0f - dr.p)));
}

// The plain case (pre-activations in `gates`, no dropout) four units per thread with 16-byte accesses; no_save: the activated
// gates are not written back (forward-only callers: greedy / sampling decode have no backward pass to feed).
__global__ void lstm_pointwise_fwd_bf16_v4_kernel(float* __restrict__ gates, int ldg, const float* __restrict__ c_prev, int ldcp,
                                                  float* __restrict__ c_out, int ldc, bf16* __restrict__ h_out, int ldh,
                                                  float* __restrict__ h2, int ldh2, int B, int H, int no_save) {
  pdl_launch_dependents();
  pdl_wait();
  const int H4 = H >> 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H4) return;
  const int b = idx / H4, j = (idx - b * H4) * 4;
  float* g = gates + (size_t)b * ldg + j;
  const float4 pi = *reinterpret_cast<const float4*>(g), pf = *reinterpret_cast<const float4*>(g + H),
               pg = *reinterpret_cast<const float4*>(g + 2 * H), po = *reinterpret_cast<const float4*>(g + 3 * H);
  const float4 cp = c_prev ? *reinterpret_cast<const float4*>(c_prev + (size_t)b * ldcp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 i = make_float4(sigm(pi.x), sigm(pi.y), sigm(pi.z), sigm(pi.w)), f = make_float4(sigm(pf.x), sigm(pf.y), sigm(pf.z), sigm(pf.w)),
               gg = make_float4(tanhf(pg.x), tanhf(pg.y), tanhf(pg.z), tanhf(pg.w)), o = make_float4(sigm(po.x), sigm(po.y), sigm(po.z), sigm(po.w));
  const float4 c = make_float4(f.x * cp.x + i.x * gg.x, f.y * cp.y + i.y * gg.y, f.z * cp.z + i.z * gg.z, f.w * cp.w + i.w * gg.w);
  const float4 h = make_float4(o.x * tanhf(c.x), o.y * tanhf(c.y), o.z * tanhf(c.z), o.w * tanhf(c.w));
  if (!no_save) {
    *reinterpret_cast<float4*>(g) = i; *reinterpret_cast<float4*>(g + H) = f;
    *reinterpret_cast<float4*>(g + 2 * H) = gg; *reinterpret_cast<float4*>(g + 3 * H) = o;
  }
  *reinterpret_cast<float4*>(c_out + (size_t)b * ldc + j) = c;
  const __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(h_out + (size_t)b * ldh + j) = u;
  if (h2) *reinterpret_cast<float4*>(h2 + (size_t)b * ldh2 + j) = h;
}

__global__ void lstm_pointwise_bwd_bf16_kernel(const float* __restrict__ acts, int ldg, const float* __restrict__ c_prev,
                                               int ldcp, const float* __restrict__ c_new, int ldc,
                                               const float* __restrict__ dh0, int ldh0, int n0, long long s0,
                                               const float* __restrict__ dh1, int ldh1, int n1, long long s1,
                                               const float* __restrict__ dh2, int ldh2, float* __restrict__ dc, int lddc,
                                               int dc_is_zero, bf16* __restrict__ dg, int lddg, int B, int H, DropSpec dr) {
  pdl_launch_dependents();
  pdl_wait();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  int b = idx / H, j = idx % H;
  float dh = 0.f;
  if (dh0)
    for (int s = 0; s < n0; ++s) dh += dh0[(size_t)s * s0 + (size_t)b * ldh0 + j];
  if (dh1) {   // gradient from the layer above, through this layer's dropout mask when dr.p > 0
    float d1 = 0.f;
    for (int s = 0; s < n1; ++s) d1 += dh1[(size_t)s * s1 + (size_t)b * ldh1 + j];
    if (dr.p > 0.f) d1 *= drop_scale(dr.seed + (dr.ctr ? *dr.ctr : 0ull), dr.sid, dr.base + (unsigned long long)idx, dr.p, 1.0f / (1.0f - dr.p));
    dh += d1;
  }
  if (dh2) dh += dh2[(size_t)b * ldh2 + j];
  const float* a = acts + (size_t)b * ldg;
  float i = a[j], f = a[H + j], gg = a[2 * H + j], o = a[3 * H + j];
  float cp = c_prev ? c_prev[(size_t)b * ldcp + j] : 0.f;
  float tc = tanhf(c_new[(size_t)b * ldc + j]);
  float dct = (dc_is_zero ? 0.f : dc[(size_t)b * lddc + j]) + dh * o * (1.f - tc * tc);
  bf16* d = dg + (size_t)b * lddg;
  d[j] = __float2bfloat16_rn(dct * gg * i * (1.f - i));
  d[H + j] = __float2bfloat16_rn(dct * cp * f * (1.f - f));
  d[2 * H + j] = __float2bfloat16_rn(dct * i * (1.f - gg * gg));
  d[3 * H + j] = __float2bfloat16_rn(dh * tc * o * (1.f - o));
  dc[(size_t)b * lddc + j] = dct * f;
}

// out(n) (+)= sum_m X(m,n), X bf16.  grid (ceil(N/64), row chunks); each warp row covers 64 columns
// (128 contiguous bytes), 8 row lanes per block; chunk sums are added atomically, so the
// caller zeroes `out` first unless it accumulates.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ X, int ldx, float* __restrict__ out,
                                                          float* __restrict__ out2, int M, int N, int rows_per_block) {
  __shared__ float sh[8][65];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 64 + 2 * tx;
  const int m_begin = blockIdx.y * rows_per_block, m_end = min(M, m_begin + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (n + 1 < N && (ldx & 1) == 0) {
    for (int m = m_begin + ty; m < m_end; m += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(X + (size_t)m * ldx + n);
      s0 += __low2float(v); s1 += __high2float(v);
    }
  } else {
    for (int m = m_begin + ty; m < m_end; m += 8) {
      if (n < N) s0 += __bfloat162float(X[(size_t)m * ldx + n]);
      if (n + 1 < N) s1 += __bfloat162float(X[(size_t)m * ldx + n + 1]);
    }
  }
  sh[ty][2 * tx] = s0; sh[ty][2 * tx + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < N) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
      atomicAdd(out + c, t);
      if (out2) atomicAdd(out2 + c, t);
    }
  }
}

__device__ __forceinline__ float wmax_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Row-wise log-softmax + NLL on fp32 logits; dlogits = scale*(softmax - onehot) written as bf16.
__global__ void __launch_bounds__(256) nll_rows_bf16_kernel(const float* __restrict__ logits, int ldl,
                                                            const int64_t* __restrict__ targets,
                                                            float* __restrict__ nll, int R, int V, float scale,
                                                            bf16* __restrict__ dl, int lddl, const float* __restrict__ row_w) {
  __shared__ float sh[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + (size_t)r * ldl;
  float m = -INFINITY;
  for (int v = tid; v < V; v += 256) m = fmaxf(m, x[v]);
  m = wmax_(m);
  if ((tid & 31) == 0) sh[tid >> 5] = m;
  __syncthreads();
  m = sh[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, sh[w]);
  __syncthreads();
  float s = 0.f;
  for (int v = tid; v < V; v += 256) s += expf(x[v] - m);
  s = wsum_(s);
  if ((tid & 31) == 0) sh[tid >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < 8; ++w) s += sh[w];
  const float lse = m + logf(s);
  long long t = targets[r];
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);
  const float wr = row_w ? row_w[r] : 1.f;      // 0 for target steps beyond the sample's length
  if (tid == 0) nll[r] = wr * (lse - x[t]);
  if (dl) {
    bf16* d = dl + (size_t)r * lddl;
    const float sc = scale * wr;
    for (int v = tid; v < V; v += 256) d[v] = __float2bfloat16_rn(sc * (expf(x[v] - lse) - (v == t ? 1.f : 0.f)));
  }
}

// ---- variable-length batches (mmqg_batch.ctx_len / tgt_len / n_frames) -----------------------------
// Sequences are RIGHT-aligned in time inside the fixed (T, B) layout: sample b's first real step is
// t = shift[b] = T - len[b], so every sample's final state sits at t = T-1 (what the decoder takes
// over) and the masked steps before it simply hold the zero initial state.
__global__ void prep_lengths_kernel(const int* __restrict__ ctx_len, const int* __restrict__ tgt_len, const int* __restrict__ n_frames,
                                    int* __restrict__ shift_t, int* __restrict__ shift_v, float* __restrict__ row_w, int B, int T_t,
                                    int T_v, int T_q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    const int cl = ctx_len ? min(max(ctx_len[i], 1), T_t) : T_t;
    const int nf = n_frames ? min(max(n_frames[i], 1), T_v) : T_v;
    shift_t[i] = T_t - cl;
    shift_v[i] = T_v - nf;
  }
  if (i < B * T_q) {
    const int t = i / B, b = i % B;
    const int tl = tgt_len ? min(max(tgt_len[b], 1), T_q) : T_q;
    row_w[i] = t < tl ? 1.f : 0.f;
  }
}
int prep_lengths(const int* ctx_len, const int* tgt_len, const int* n_frames, int* shift_t, int* shift_v, float* row_w, int B,
                 int T_t, int T_v, int T_q, cudaStream_t st) {
  const int n = B * (T_q > 1 ? T_q : 1);
  prep_lengths_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx_len, tgt_len, n_frames, shift_t, shift_v, row_w, B, T_t, T_v, T_q);
  MMQG_LAUNCH_CHECK();
  return 0;
}
// frames (B, T_v, F) fp32 batch-major -> (T_v*B, F) bf16 time-major, right-aligned by shift_v (zeros before)
__global__ void frames_tm_kernel(const float* __restrict__ frames, bf16* __restrict__ out, const int* __restrict__ shift_v, int B,
                                 int T_v, int F) {
  const int row = blockIdx.x;                 // t*B + b
  const int t = row / B, b = row % B;
  const int sh = shift_v ? shift_v[b] : 0;
  bf16* dst = out + (size_t)row * F;
  if (t < sh) {
    for (int f = threadIdx.x; f < F; f += blockDim.x) dst[f] = __float2bfloat16_rn(0.f);
  } else {
    const float* src = frames + ((size_t)b * T_v + (t - sh)) * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) dst[f] = __float2bfloat16_rn(src[f]);
  }
}
int frames_to_time_major_bf16(const float* frames, void* out, const int* shift_v, int B, int T_v, int F, cudaStream_t st) {
  frames_tm_kernel<<<B * T_v, 256, 0, st>>>(frames, reinterpret_cast<bf16*>(out), shift_v, B, T_v, F);
  MMQG_LAUNCH_CHECK();
  return 0;
}
// fp32 twin (parity modes): frames (B, T_v, F) -> (T_v*B, F) time-major, right-aligned by shift_v (zeros before)
__global__ void frames_tm_f32_kernel(const float* __restrict__ frames, float* __restrict__ out, const int* __restrict__ shift_v, int B,
                                     int T_v, int F) {
  const int row = blockIdx.x;                 // t*B + b
  const int t = row / B, b = row % B;
  const int sh = shift_v ? shift_v[b] : 0;
  float* dst = out + (size_t)row * F;
  const float* src = frames + ((size_t)b * T_v + (t - sh)) * F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) dst[f] = t < sh ? 0.f : src[f];
}
int frames_to_time_major_f32(const float* frames, float* out, const int* shift_v, int B, int T_v, int F, cudaStream_t st) {
  frames_tm_f32_kernel<<<B * T_v, 256, 0, st>>>(frames, out, shift_v, B, T_v, F);
  MMQG_LAUNCH_CHECK();
  return 0;
}
// audio (B, T_v, H_a) -> zero-padded memory (B, AM, H_a); rows >= n_frames[b] are zero (train.py:156)
__global__ void audio_pad_kernel(const float* __restrict__ audio, float* __restrict__ m_aud, const int* __restrict__ n_frames, int B,
                                 int T_v, int AM, int H_a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * AM * H_a) return;
  const int h = (int)(i % H_a);
  const int k = (int)((i / H_a) % AM);
  const int b = (int)(i / ((long long)H_a * AM));
  const int nf = n_frames ? min(max(n_frames[b], 1), T_v) : T_v;
  m_aud[i] = k < nf ? audio[((size_t)b * T_v + k) * H_a + h] : 0.f;
}
int audio_pad(const float* audio, float* m_aud, const int* n_frames, int B, int T_v, int AM, int H_a, cudaStream_t st) {
  const long long n = (long long)B * AM * H_a;
  audio_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(audio, m_aud, n_frames, B, T_v, AM, H_a);
  MMQG_LAUNCH_CHECK();
  return 0;
}

__global__ void bump_counter_kernel(unsigned long long* ctr) { *ctr += 1ull; }
int bump_counter(unsigned long long* ctr, cudaStream_t st) {
  bump_counter_kernel<<<1, 1, 0, st>>>(ctr);
  MMQG_LAUNCH_CHECK();
  return 0;
}

__global__ void dropout_bf16_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n, unsigned long long seed,
                                    const unsigned long long* __restrict__ ctr, int sid,
                                    unsigned long long base, float p, float inv_keep) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(__bfloat162float(x[i]) * drop_scale(seed + (ctr ? *ctr : 0ull), sid, base + i, p, inv_keep));
}

// fp32 twin of dropout_bf16 (fp32 parity mode): y = x * mask
__global__ void dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, unsigned long long seed,
                                   const unsigned long long* __restrict__ ctr, int sid, unsigned long long base, float p, float inv_keep) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * drop_scale(seed + (ctr ? *ctr : 0ull), sid, base + i, p, inv_keep);
}

// x[k*stride + i] *= mask(i) for k < n_part (gradient w.r.t. a dropped tensor, possibly split-K partials)
__global__ void dropout_scale_f32_kernel(float* __restrict__ x, int n_part, long long stride, long long n, unsigned long long seed,
                                         const unsigned long long* __restrict__ ctr,
                                         int sid, unsigned long long base, float p, float inv_keep) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float m = drop_scale(seed + (ctr ? *ctr : 0ull), sid, base + i, p, inv_keep);
  for (int k = 0; k < n_part; ++k) x[(size_t)k * stride + i] *= m;
}

__global__ void dropout_mask_kernel(float* __restrict__ out, long long n, unsigned long long seed, int sid, float p, float inv_keep) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = drop_scale(seed, sid, i, p, inv_keep);
}

int dropout_bf16(const void* x, void* y, long long n, unsigned long long seed, const unsigned long long* ctr, int sid,
                 unsigned long long base, float p, cudaStream_t st) {
  MMQG_REQUIRE(x && y && n > 0 && p >= 0.f && p < 1.f, "dropout_bf16: bad args");
  dropout_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), n, seed, ctr,
                                                                  sid, base, p, 1.0f / (1.0f - p));
  MMQG_LAUNCH_CHECK();
  return 0;
}
int dropout_f32(const float* x, float* y, long long n, unsigned long long seed, const unsigned long long* ctr, int sid,
                unsigned long long base, float p, cudaStream_t st) {
  MMQG_REQUIRE(x && y && n > 0 && p >= 0.f && p < 1.f, "dropout_f32: bad args");
  dropout_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, n, seed, ctr, sid, base, p, 1.0f / (1.0f - p));
  MMQG_LAUNCH_CHECK();
  return 0;
}
int dropout_scale_f32(float* x, int n_part, long long stride, long long n, unsigned long long seed, const unsigned long long* ctr, int sid, unsigned long long base,
                      float p, cudaStream_t st) {
  MMQG_REQUIRE(x && n > 0 && n_part > 0 && p >= 0.f && p < 1.f, "dropout_scale_f32: bad args");
  dropout_scale_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n_part, stride, n, seed, ctr, sid, base, p, 1.0f / (1.0f - p));
  MMQG_LAUNCH_CHECK();
  return 0;
}
int dropout_mask(float* out, long long n, unsigned long long seed, int sid, float p, cudaStream_t st) {
  MMQG_REQUIRE(out && n > 0 && p >= 0.f && p < 1.f, "dropout_mask: bad args");
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, seed, sid, p, 1.0f / (1.0f - p));
  MMQG_LAUNCH_CHECK();
  return 0;
}

// out(r, c) = sum_k part[k*stride + r*N + c]   (split-K partial tiles -> the gradient tensor)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int n_part, long long stride, float* __restrict__ out,
                                       int ldo, long long rows, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 (or tail) per thread
  const int n4 = (N + 3) / 4;
  if (i >= rows * n4) return;
  const long long r = i / n4;
  const int c = (int)(i % n4) * 4;
  if (c + 4 <= N && (N & 3) == 0 && (ldo & 3) == 0) {
    float4 a = *reinterpret_cast<const float4*>(part + r * N + c);
    for (int k = 1; k < n_part; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(part + (size_t)k * stride + r * N + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(out + r * ldo + c) = a;
  } else {
    for (int j = c; j < N && j < c + 4; ++j) {
      float a = 0.f;
      for (int k = 0; k < n_part; ++k) a += part[(size_t)k * stride + r * N + j];
      out[r * ldo + j] = a;
    }
  }
}
int reduce_partials(const float* part, int n_part, long long stride, float* out, int ldo, long long rows, int N, cudaStream_t st) {
  MMQG_REQUIRE(part && out && n_part > 0 && rows > 0 && N > 0, "reduce_partials: bad args");
  MMQG_REQUIRE((N & 3) != 0 || ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 || (ldo & 3) != 0,
               "reduce_partials: unaligned pointers");
  const long long n = rows * ((N + 3) / 4);
  reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, n_part, stride, out, ldo, rows, N);
  MMQG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
int cvt_f32_bf16_2d(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols,
                    int cols_dst, cudaStream_t st) {
  MMQG_REQUIRE(src && dst && rows > 0 && cols > 0 && cols_dst >= cols, "cvt_f32_bf16_2d: bad args");
  long long total = rows * cols_dst;
  cvt_f32_bf16_2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, ld_src, reinterpret_cast<bf16*>(dst), ld_dst,
                                                                       rows, cols, cols_dst);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int embedding_gather_bf16(const float* emb, const int64_t* idx, void* out, int ldo, int N, int E, int E_pad, int V,
                          cudaStream_t st) {
  MMQG_REQUIRE(emb && idx && out && N > 0 && ldo >= E_pad && E_pad >= E, "embedding_gather_bf16: bad args");
  MMQG_PROBE(KC_EMBED, 0, 6.0 * N * E);
  embedding_gather_bf16_kernel<<<N, 128, 0, st>>>(emb, idx, reinterpret_cast<bf16*>(out), ldo, N, E, E_pad, V);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int lstm_pointwise_fwd_bf16(float* gates, int ldg, const float* c_prev, int ldcp, float* c_out, int ldc, void* h_out,
                            int ldh, float* h2, int ldh2, int B, int H, cudaStream_t st, DropSpec dr, PreSpec ps) {
  MMQG_REQUIRE(gates && c_out && h_out && B > 0 && H > 0, "lstm_pointwise_fwd_bf16: bad args");
  int n = B * H;
  MMQG_PROBE(KC_POINTWISE, 0, 4.0 * n * (c_prev ? 10 : 9) + 2.0 * n + (h2 ? 4.0 * n : 0));
  auto al = [](const void* p, int a) { return reinterpret_cast<uintptr_t>(p) % a == 0; };
  if (!ps.part && !dr.out && H % 4 == 0 && ldg % 4 == 0 && ldc % 4 == 0 && ldh % 4 == 0 && (!c_prev || ldcp % 4 == 0) && (!h2 || ldh2 % 4 == 0) &&
      al(gates, 16) && al(c_out, 16) && al(h_out, 8) && (!c_prev || al(c_prev, 16)) && (!h2 || al(h2, 16))) {
    MMQG_CUDA(launch_k(lstm_pointwise_fwd_bf16_v4_kernel, dim3(ceil_div(n / 4, 256)), dim3(256), 0, st, gates, ldg, c_prev, ldcp, c_out, ldc,
                       reinterpret_cast<bf16*>(h_out), ldh, h2, ldh2, B, H, ps.no_save));
    MMQG_LAUNCH_CHECK();
    return 0;
  }
  MMQG_CUDA(launch_k(lstm_pointwise_fwd_bf16_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, st, gates, ldg, c_prev, ldcp, c_out, ldc,
                     reinterpret_cast<bf16*>(h_out), ldh, h2, ldh2, B, H, dr, ps));
  MMQG_LAUNCH_CHECK();
  return 0;
}

int lstm_pointwise_bwd_bf16(const float* acts, int ldg, const float* c_prev, int ldcp, const float* c_new, int ldc,
                            const float* dh0, int ldh0, int n0, long long s0, const float* dh1, int ldh1, int n1,
                            long long s1, const float* dh2, int ldh2, float* dc, int lddc, int dc_is_zero, void* dg,
                            int lddg, int B, int H, cudaStream_t st, DropSpec dr) {
  MMQG_REQUIRE(acts && c_new && dc && dg && B > 0 && H > 0, "lstm_pointwise_bwd_bf16: bad args");
  int n = B * H;
  MMQG_PROBE(KC_POINTWISE, 0, 4.0 * n * (6 + (c_prev ? 1 : 0) + (dh0 ? n0 : 0) + (dh1 ? n1 : 0) + (dh2 ? 1 : 0)) + 8.0 * n);
  MMQG_CUDA(launch_k(lstm_pointwise_bwd_bf16_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, st, acts, ldg, c_prev, ldcp, c_new, ldc,
                     dh0, ldh0, n0, s0, dh1, ldh1, n1, s1, dh2, ldh2, dc, lddc, dc_is_zero, reinterpret_cast<bf16*>(dg), lddg,
                     B, H, dr));
  MMQG_LAUNCH_CHECK();
  return 0;
}

int colsum_bf16(const void* X, int ldx, float* out, float* out2, int M, int N, float beta, cudaStream_t st) {
  MMQG_REQUIRE(X && out && M > 0 && N > 0, "colsum_bf16: bad args");
  MMQG_REQUIRE(beta == 0.f || beta == 1.f, "colsum_bf16: beta must be 0 or 1");
  if (beta == 0.f) {
    MMQG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
    if (out2) MMQG_CUDA(cudaMemsetAsync(out2, 0, sizeof(float) * N, st));
  }
  const int col_blocks = ceil_div(N, 64);
  int chunks = ceil_div(4 * 148, col_blocks);
  if (chunks > ceil_div(M, 64)) chunks = ceil_div(M, 64);
  const int rows_per_block = ceil_div(M, chunks);
  colsum_bf16_kernel<<<dim3(col_blocks, ceil_div(M, rows_per_block)), 256, 0, st>>>(reinterpret_cast<const bf16*>(X), ldx, out,
                                                                                     out2, M, N, rows_per_block);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int nll_rows_bf16(const float* logits, int ldl, const int64_t* targets, float* nll, int R, int V, float scale,
                  void* dlogits, int lddl, cudaStream_t st, const float* row_w) {
  MMQG_REQUIRE(logits && targets && nll && R > 0 && V > 0, "nll_rows_bf16: bad args");
  MMQG_PROBE(KC_LOSS, 0, 4.0 * R * V + (dlogits ? 2.0 * R * V : 0));
  nll_rows_bf16_kernel<<<R, 256, 0, st>>>(logits, ldl, targets, nll, R, V, scale, reinterpret_cast<bf16*>(dlogits), lddl, row_w);
  MMQG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmqg

// Exported for tests: the multiplicative mask (0 or 1/(1-p)) of dropout stream `sid` (text layer l:
// 10+l, decoder layer l: 20+l), element index (t*B + b)*H + j.
extern "C" int mmqg_dropout_mask(float* out, long long n, unsigned long long seed, int sid, float p, void* stream) {
  return mmqg::dropout_mask(out, n, seed, sid, p, mmqg::as_stream(stream));
}

// ---- bf16 gradient exchange: flat fp32 bucket <-> bf16 staging buffer (mmqg/dp.py) ---------------------------------
namespace mmqg {
__global__ void pack_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  for (long long i = 4 * n4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void unpack_bf16_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint2 u = reinterpret_cast<const uint2*>(src)[i];
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    reinterpret_cast<float4*>(dst)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  for (long long i = 4 * n4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}
}  // namespace mmqg

extern "C" int mmqg_pack_bf16(const float* src, void* dst_bf16, long long n, void* stream) {
  using namespace mmqg;
  MMQG_REQUIRE(src && dst_bf16 && n > 0, "pack_bf16: bad args");
  MMQG_REQUIRE(((reinterpret_cast<uintptr_t>(src) & 15) | (reinterpret_cast<uintptr_t>(dst_bf16) & 7)) == 0, "pack_bf16: src must be 16-byte, dst 8-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int blocks = (int)std::min<long long>(148 * 8, (n / 4 + 255) / 256 + 1);
  MMQG_PROBE(KC_EMBED, 0, 6.0 * n);
  pack_bf16_kernel<<<blocks, 256, 0, st>>>(src, reinterpret_cast<bf16*>(dst_bf16), n);
  MMQG_LAUNCH_CHECK();
  return 0;
}
extern "C" int mmqg_unpack_bf16(const void* src_bf16, float* dst, long long n, void* stream) {
  using namespace mmqg;
  MMQG_REQUIRE(src_bf16 && dst && n > 0, "unpack_bf16: bad args");
  MMQG_REQUIRE(((reinterpret_cast<uintptr_t>(dst) & 15) | (reinterpret_cast<uintptr_t>(src_bf16) & 7)) == 0, "unpack_bf16: dst must be 16-byte, src 8-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int blocks = (int)std::min<long long>(148 * 8, (n / 4 + 255) / 256 + 1);
  MMQG_PROBE(KC_EMBED, 0, 6.0 * n);
  unpack_bf16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(src_bf16), dst, n);
  MMQG_LAUNCH_CHECK();
  return 0;
}

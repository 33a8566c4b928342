// Elementwise / row-wise kernels of the hot path (HBM- or L2-bound; fp32).
//   lstm_pointwise_{fwd,bwd}  : the cell update of aten::lstm (reference encoder.py:69,98, decoder.py:104)
//   embedding gather/scatter  : encoder.py:96, decoder.py:75 and embedding_dense_backward
//   nll_rows / argmax_rows    : train.py:174 (CrossEntropyLoss), train.py:107-108 (greedy)
//   colsum                    : bias gradients
#include <cuda_bf16.h>
#include "kernels.h"

namespace mmqg {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ----------------------------------------------------------------------------------------
__global__ void lstm_pointwise_fwd_kernel(float* __restrict__ gates, int ldg, const float* __restrict__ c_prev,
                                          int ldcp, float* __restrict__ c_out, int ldc, float* __restrict__ h_out,
                                          int ldh, float* __restrict__ h2, int ldh2, int B, int H, PreSpec ps, LenSpec len) {
  pdl_launch_dependents();
  pdl_wait();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  int b = idx / H, j = idx % H;
  float* g = gates + (size_t)b * ldg;
  if (len.shift) {       // right-aligned variable-length batch: the steps before a sample's first one hold the zero state
    const int sh = len.shift[b];
    if (len.t_base < sh) {     // zero gates make the backward kernel return zero gradients for this (step, sample) by itself
      g[j] = 0.f; g[H + j] = 0.f; g[2 * H + j] = 0.f; g[3 * H + j] = 0.f;
      c_out[(size_t)b * ldc + j] = 0.f;
      h_out[(size_t)b * ldh + j] = 0.f;
      if (ps.h_split) {
        __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(ps.h_split) + (size_t)b * 2 * H;
        sp[j] = __float2bfloat16_rn(0.f); sp[H + j] = __float2bfloat16_rn(0.f);
      }
      return;
    }
    if (len.mem_shift && h2) h2 += (size_t)(len.t_base - sh) * H;      // memory row = the sample's own position (h2 = row 0)
  }
  float pre[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) pre[q] = (ps.n_part == 0 || ps.add_gates) ? g[q * H + j] : 0.f;
  if (ps.n_part > 0) {      // pre-activations arrive (partly) as split-K partial sums
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ps.bias) pre[q] += ps.bias[q * H + j];
      for (int s = 0; s < ps.n_part; ++s) pre[q] += ps.part[(size_t)s * ps.stride + (size_t)b * ps.ld + q * H + j];
    }
  }
  float i = sigmoidf_acc(pre[0]);
  float f = sigmoidf_acc(pre[1]);
  float gg = tanhf(pre[2]);
  float o = sigmoidf_acc(pre[3]);
  float cp = c_prev ? c_prev[(size_t)b * ldcp + j] : 0.f;
  float c = f * cp + i * gg;
  float h = o * tanhf(c);
  g[j] = i; g[H + j] = f; g[2 * H + j] = gg; g[3 * H + j] = o;
  c_out[(size_t)b * ldc + j] = c;
  h_out[(size_t)b * ldh + j] = h;
  if (h2) h2[(size_t)b * ldh2 + j] = h;
  if (ps.h_split) {
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(ps.h_split) + (size_t)b * 2 * H;
    const __nv_bfloat16 hi = __float2bfloat16_rn(h);
    sp[j] = hi;
    sp[H + j] = __float2bfloat16_rn(h - __bfloat162float(hi));
  }
}

__global__ void lstm_pointwise_bwd_kernel(float* __restrict__ acts, int ldg, const float* __restrict__ c_prev,
                                          int ldcp, const float* __restrict__ c_new, int ldc,
                                          const float* __restrict__ dh0, int ldh0, int n0, long long s0,
                                          const float* __restrict__ dh1, int ldh1, int n1, long long s1, const float* __restrict__ dh2,
                                          int ldh2, float* __restrict__ dc, int lddc, int dc_is_zero, int B, int H,
                                          __nv_bfloat16* __restrict__ dg_split, LenSpec len) {
  pdl_launch_dependents();
  pdl_wait();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  int b = idx / H, j = idx % H;
  if (len.shift) {
    const int sh = len.shift[b];
    if (len.t_base < sh) {     // masked step: no gradient flows through it or out of it
      float* a0 = acts + (size_t)b * ldg;
      a0[j] = 0.f; a0[H + j] = 0.f; a0[2 * H + j] = 0.f; a0[3 * H + j] = 0.f;
      dc[(size_t)b * lddc + j] = 0.f;
      if (dg_split) {
        __nv_bfloat16* sp = dg_split + (size_t)b * 8 * H;
        for (int q = 0; q < 8; ++q) sp[q * H + j] = __float2bfloat16_rn(0.f);
      }
      return;
    }
    if (len.mem_shift && dh2) dh2 += (size_t)(len.t_base - sh) * H;     // external gradient row = the sample's own position
  }
  float dh = 0.f;
  if (dh0)
    for (int s = 0; s < n0; ++s) dh += dh0[(size_t)s * s0 + (size_t)b * ldh0 + j];
  if (dh1)
    for (int s = 0; s < n1; ++s) dh += dh1[(size_t)s * s1 + (size_t)b * ldh1 + j];
  if (dh2) dh += dh2[(size_t)b * ldh2 + j];
  float* a = acts + (size_t)b * ldg;
  float i = a[j], f = a[H + j], gg = a[2 * H + j], o = a[3 * H + j];
  float cp = c_prev ? c_prev[(size_t)b * ldcp + j] : 0.f;
  float tc = tanhf(c_new[(size_t)b * ldc + j]);
  float dct = (dc_is_zero ? 0.f : dc[(size_t)b * lddc + j]) + dh * o * (1.f - tc * tc);
  const float dg[4] = {dct * gg * i * (1.f - i), dct * cp * f * (1.f - f), dct * i * (1.f - gg * gg), dh * tc * o * (1.f - o)};
#pragma unroll
  for (int q = 0; q < 4; ++q) a[q * H + j] = dg[q];
  dc[(size_t)b * lddc + j] = dct * f;
  if (dg_split) {      // [hi(4H) | lo(4H)] rows of pitch 8H: the pre-split A operand of dG . W (gemm_f32x3.cu)
    __nv_bfloat16* sp = dg_split + (size_t)b * 8 * H;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(dg[q]);
      sp[q * H + j] = hi;
      sp[4 * H + q * H + j] = __float2bfloat16_rn(dg[q] - __bfloat162float(hi));
    }
  }
}

// ----------------------------------------------------------------------------------------
__global__ void embedding_gather_kernel(const float* __restrict__ emb, const int64_t* __restrict__ idx,
                                        float* __restrict__ out, int ldo, int N, int E, int V) {
  int n = blockIdx.x;
  long long w = idx[n];
  w = w < 0 ? 0 : (w >= V ? V - 1 : w);
  const float* src = emb + (size_t)w * E;
  float* dst = out + (size_t)n * ldo;
  for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = src[e];
}

__global__ void embedding_scatter_add_kernel(float* __restrict__ demb, const int64_t* __restrict__ idx,
                                             const float* __restrict__ dx, int N, int E, int V) {
  int n = blockIdx.x;
  long long w = idx[n];
  w = w < 0 ? 0 : (w >= V ? V - 1 : w);
  float* dst = demb + (size_t)w * E;
  const float* src = dx + (size_t)n * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(dst + e, src[e]);
}

// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide reductions for 256-thread blocks; result broadcast to every thread.
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) r += sh[w];
  return r;
}

__global__ void __launch_bounds__(256) nll_rows_kernel(float* __restrict__ logits, int ldl,
                                                       const int64_t* __restrict__ targets, long long tgt_stride,
                                                       float* __restrict__ nll, int R, int V, float scale,
                                                       const float* __restrict__ row_w) {
  __shared__ float sh[8];
  int r = blockIdx.x;
  const float rw = row_w ? row_w[r] : 1.f;
  scale *= rw;
  float* x = logits + (size_t)r * ldl;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += 256) m = fmaxf(m, x[v]);
  m = block_max(m, sh);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) s += expf(x[v] - m);
  s = block_sum(s, sh);
  float lse = m + logf(s);
  long long t = targets[(size_t)r * tgt_stride];
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);
  if (threadIdx.x == 0) nll[r] = rw * (lse - x[t]);
  if (scale != 0.f || row_w) {
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += 256) {
      float p = expf(x[v] - lse);
      x[v] = scale * (p - (v == t ? 1.f : 0.f));
    }
  }
}

__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ logits, int ldl,
                                                          int64_t* __restrict__ tokens, long long tok_stride,
                                                          int64_t* __restrict__ tokens2, int R, int V) {
  __shared__ float sv[256];
  __shared__ int si[256];
  int r = blockIdx.x;
  const float* x = logits + (size_t)r * ldl;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += 256) {
    float y = x[v];
    if (y > bv || (y == bv && v < bi)) { bv = y; bi = v; }
  }
  sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      float ov = sv[threadIdx.x + o]; int oi = si[threadIdx.x + o];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
        sv[threadIdx.x] = ov; si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int64_t w = si[0] == 0x7fffffff ? 0 : si[0];
    tokens[(size_t)r * tok_stride] = w;
    if (tokens2) tokens2[r] = w;
  }
}

// 'sampling' strategy of the reference (evaluate.py:84-90: np.random.choice(V, p=softmax(logits))):
// tokens(r) = the smallest v with cumsum_softmax(r, v) > u(r), u from the counter-based generator
// (seed, stream 30, index step*R_total + row) -- the inverse-CDF draw numpy makes, with our own uniforms.
// One block per row: softmax statistics, then every thread owns a contiguous segment of the
// vocabulary; a block scan of the segment masses finds the segment, a short serial scan the token.
__device__ __forceinline__ float sample_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + 30ull * 0x9E3779B97F4A7C15ull + idx * 0xD1342543DE82EF95ull;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
__global__ void __launch_bounds__(256) sample_rows_kernel(const float* __restrict__ logits, int ldl, int64_t* __restrict__ tokens,
                                                          long long tok_stride, int64_t* __restrict__ tokens2, int R, int V,
                                                          unsigned long long seed, unsigned long long idx_base) {
  __shared__ float sh[256];
  __shared__ float s_m, s_target;
  __shared__ int s_seg;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + (size_t)r * ldl;
  const int seg = (V + 255) / 256, v0 = tid * seg, v1 = min(V, v0 + seg);
  float m = -INFINITY;
  for (int v = v0; v < v1; ++v) m = fmaxf(m, x[v]);
  sh[tid] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) sh[tid] = fmaxf(sh[tid], sh[tid + o]);
    __syncthreads();
  }
  if (tid == 0) s_m = sh[0];
  __syncthreads();
  m = s_m;
  float mass = 0.f;
  for (int v = v0; v < v1; ++v) mass += expf(x[v] - m);
  __syncthreads();
  sh[tid] = mass;
  __syncthreads();
  if (tid == 0) {      // serial scan of 256 segment masses: deterministic, and tiny next to the passes over V
    float tot = 0.f;
    for (int i = 0; i < 256; ++i) tot += sh[i];
    const float target = sample_uniform(seed, idx_base + r) * tot;
    float acc = 0.f;
    int k = 0;
    for (; k < 255; ++k) {
      if (acc + sh[k] > target) break;
      acc += sh[k];
    }
    s_seg = k;
    s_target = target - acc;
  }
  __syncthreads();
  if (tid == s_seg) {
    float acc = 0.f;
    int w = v1 > v0 ? v1 - 1 : 0;
    for (int v = v0; v < v1; ++v) {
      acc += expf(x[v] - m);
      if (acc > s_target) { w = v; break; }
    }
    tokens[(size_t)r * tok_stride] = w;
    if (tokens2) tokens2[r] = w;
  }
}
__global__ void sample_uniform_kernel(float* out, int n, unsigned long long seed, unsigned long long base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sample_uniform(seed, base + i);
}

// out(n) = beta*out(n) + sum_m X(m,n).  Block = 32 columns x 8 row lanes.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int ldx, float* __restrict__ out,
                                                     float* __restrict__ out2, int M, int N, float beta) {
  __shared__ float sh[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int m = ty; m < M; m += 8) s += X[(size_t)m * ldx + n];
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][tx];
    t += (beta != 0.f ? beta * out[n] : 0.f);
    out[n] = t;
    if (out2) out2[n] = t;
  }
}

// y(n) = a(n) + b(n)
__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

// Time-major token index arrays from the batch-major inputs (reference train.py:164-175):
//   idx_ctx(t*B+b) = context(b,t);  idx_dec(t*B+b) = t==0 ? <start> : target(b,t-1);  tgt_tm(t*B+b) = target(b,t)
__global__ void build_indices_kernel(const int64_t* __restrict__ ctx, const int64_t* __restrict__ tgt,
                                     int64_t* __restrict__ idx_ctx, int64_t* __restrict__ idx_dec,
                                     int64_t* __restrict__ tgt_tm, int B, int T_t, int T_q,
                                     const int* __restrict__ shift_t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * T_t) {
    int t = i / B, b = i % B;
    // variable lengths: sample b is right-aligned in time (its first token sits at t = shift_t[b]);
    // the steps before it are masked in the recurrent kernels, any valid index will do there
    const int sh = shift_t ? shift_t[b] : 0;
    idx_ctx[i] = t >= sh ? ctx[(size_t)b * T_t + t - sh] : 0;
  }
  if (tgt && i < B * T_q) {
    int t = i / B, b = i % B;
    idx_dec[i] = t == 0 ? 1 : tgt[(size_t)b * T_q + t - 1];
    tgt_tm[i] = tgt[(size_t)b * T_q + t];
  }
}

__global__ void fill_i64_kernel(int64_t* p, int n, int64_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// loss = scale * sum_r nll(r), single block, fixed order (deterministic).
__global__ void __launch_bounds__(256) sum_scale_kernel(const float* __restrict__ x, int n, float scale,
                                                        float* __restrict__ out) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += x[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// ---------------------------------------------------------------------------- host wrappers
int lstm_pointwise_fwd(float* gates, int ldg, const float* c_prev, int ldcp, float* c_out, int ldc, float* h_out,
                       int ldh, float* h2, int ldh2, int B, int H, cudaStream_t st, PreSpec ps, LenSpec len) {
  MMQG_REQUIRE(gates && c_out && h_out && B > 0 && H > 0, "lstm_pointwise_fwd: bad args");
  int n = B * H;
  MMQG_PROBE(KC_POINTWISE, 0, 4.0 * n * (c_prev ? 11 : 10) + (h2 ? 4.0 * n : 0));
  MMQG_CUDA(launch_k(lstm_pointwise_fwd_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, st, gates, ldg, c_prev, ldcp, c_out, ldc, h_out, ldh, h2,
                     ldh2, B, H, ps, len));
  MMQG_LAUNCH_CHECK();
  return 0;
}

int lstm_pointwise_bwd(float* acts, int ldg, const float* c_prev, int ldcp, const float* c_new, int ldc,
                       const float* dh0, int ldh0, int n0, long long s0, const float* dh1, int ldh1, int n1, long long s1,
                       const float* dh2, int ldh2, float* dc, int lddc, int dc_is_zero, int B, int H,
                       cudaStream_t st, void* dg_split, LenSpec len) {
  MMQG_REQUIRE(acts && c_new && dc && B > 0 && H > 0, "lstm_pointwise_bwd: bad args");
  int n = B * H;
  MMQG_PROBE(KC_POINTWISE, 0, 4.0 * n * (10 + (c_prev ? 1 : 0) + (dh0 ? n0 : 0) + (dh1 ? n1 : 0) + (dh2 ? 1 : 0)));
  MMQG_CUDA(launch_k(lstm_pointwise_bwd_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, st, acts, ldg, c_prev, ldcp, c_new, ldc, dh0, ldh0, n0, s0,
                     dh1, ldh1, n1, s1, dh2, ldh2, dc, lddc, dc_is_zero, B, H, reinterpret_cast<__nv_bfloat16*>(dg_split), len));
  MMQG_LAUNCH_CHECK();
  return 0;
}

int embedding_gather(const float* emb, const int64_t* idx, float* out, int ldo, int N, int E, int V, cudaStream_t st) {
  MMQG_REQUIRE(emb && idx && out && N > 0 && E > 0 && V > 0 && ldo >= E, "embedding_gather: bad args");
  MMQG_PROBE(KC_EMBED, 0, 8.0 * N * E);
  embedding_gather_kernel<<<N, 128, 0, st>>>(emb, idx, out, ldo, N, E, V);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int embedding_scatter_add(float* demb, const int64_t* idx, const float* dx, int N, int E, int V, cudaStream_t st) {
  MMQG_REQUIRE(demb && idx && dx && N > 0 && E > 0 && V > 0, "embedding_scatter_add: bad args");
  MMQG_PROBE(KC_EMBED, 0, 12.0 * N * E);
  embedding_scatter_add_kernel<<<N, 128, 0, st>>>(demb, idx, dx, N, E, V);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int nll_rows(float* logits, int ldl, const int64_t* targets, long long tgt_stride, float* nll, int R, int V,
             float scale, cudaStream_t st, const float* row_w) {
  MMQG_REQUIRE(logits && targets && nll && R > 0 && V > 0, "nll_rows: bad args");
  MMQG_PROBE(KC_LOSS, 0, 4.0 * R * V * (scale != 0.f ? 2 : 1));
  nll_rows_kernel<<<R, 256, 0, st>>>(logits, ldl, targets, tgt_stride, nll, R, V, scale, row_w);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int argmax_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int64_t* tokens2, int R, int V,
                cudaStream_t st) {
  MMQG_REQUIRE(logits && tokens && R > 0 && V > 0, "argmax_rows: bad args");
  argmax_rows_kernel<<<R, 256, 0, st>>>(logits, ldl, tokens, tok_stride, tokens2, R, V);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int sample_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int64_t* tokens2, int R, int V,
                unsigned long long seed, unsigned long long step, int row0, int R_total, cudaStream_t st) {
  MMQG_REQUIRE(logits && tokens && R > 0 && V > 0, "sample_rows: bad args");
  // the uniform of row r at step t is indexed t*R_total + row0 + r, whatever the row chunking
  sample_rows_kernel<<<R, 256, 0, st>>>(logits, ldl, tokens, tok_stride, tokens2, R, V, seed,
                                         step * (unsigned long long)R_total + (unsigned long long)row0);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int colsum(const float* X, int ldx, float* out, float* out2, int M, int N, float beta, cudaStream_t st) {
  MMQG_REQUIRE(X && out && M > 0 && N > 0, "colsum: bad args");
  colsum_kernel<<<ceil_div(N, 32), 256, 0, st>>>(X, ldx, out, out2, M, N, beta);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int add2(const float* a, const float* b, float* y, int n, cudaStream_t st) {
  add2_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a, b, y, n);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int build_indices(const int64_t* ctx, const int64_t* tgt, int64_t* idx_ctx, int64_t* idx_dec, int64_t* tgt_tm, int B,
                  int T_t, int T_q, cudaStream_t st, const int* shift_t) {
  int n = B * (T_t > T_q ? T_t : T_q);
  build_indices_kernel<<<ceil_div(n, 256), 256, 0, st>>>(ctx, tgt, idx_ctx, idx_dec, tgt_tm, B, T_t, T_q, shift_t);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int fill_i64(int64_t* p, int n, int64_t v, cudaStream_t st) {
  fill_i64_kernel<<<ceil_div(n, 256), 256, 0, st>>>(p, n, v);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int sum_scale(const float* x, int n, float scale, float* out, cudaStream_t st) {
  sum_scale_kernel<<<1, 256, 0, st>>>(x, n, scale, out);
  MMQG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmqg

using namespace mmqg;

extern "C" {

int mmqg_abi_version(void) { return MMQG_ABI_VERSION; }
const char* mmqg_last_error(void) { return g_err; }
unsigned long long mmqg_launch_count(void) { return g_launches.load(); }

int mmqg_device_ok(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return set_err(MMQG_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
  if (prop.major != 10) return set_err(MMQG_ERR_ARCH, "device %d is sm_%d%d; libmmqg is built for sm_100a only", dev,
                                        prop.major, prop.minor);
  return 0;
}

int mmqg_embedding_gather(const float* emb, const int64_t* idx, float* out, int N, int E, int V, void* stream) {
  return embedding_gather(emb, idx, out, E, N, E, V, as_stream(stream));
}
int mmqg_embedding_scatter_add(float* demb, const int64_t* idx, const float* dx, int N, int E, int V, void* stream) {
  return embedding_scatter_add(demb, idx, dx, N, E, V, as_stream(stream));
}
int mmqg_lstm_pointwise_fwd(float* gates, int ldg, const float* c_prev, int ldcp, float* c_out, int ldc, float* h_out,
                            int ldh, float* h2, int ldh2, int B, int H, void* stream) {
  return lstm_pointwise_fwd(gates, ldg, c_prev, ldcp, c_out, ldc, h_out, ldh, h2, ldh2, B, H, as_stream(stream), PreSpec());
}
int mmqg_lstm_pointwise_bwd(float* acts, int ldg, const float* c_prev, int ldcp, const float* c_new, int ldc,
                            const float* dh0, int ldh0, int n0, long long s0, const float* dh1, int ldh1, int n1, long long s1,
                            const float* dh2, int ldh2, float* dc, int lddc, int dc_is_zero, int B, int H,
                            void* stream) {
  return lstm_pointwise_bwd(acts, ldg, c_prev, ldcp, c_new, ldc, dh0, ldh0, n0, s0, dh1, ldh1, n1, s1, dh2, ldh2, dc, lddc,
                            dc_is_zero, B, H, as_stream(stream));
}
int mmqg_nll_rows(float* logits, int ldl, const int64_t* targets, long long tgt_stride, float* nll, int R, int V,
                  float dlogits_scale, void* stream) {
  return nll_rows(logits, ldl, targets, tgt_stride, nll, R, V, dlogits_scale, as_stream(stream));
}
int mmqg_argmax_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int R, int V, void* stream) {
  return argmax_rows(logits, ldl, tokens, tok_stride, nullptr, R, V, as_stream(stream));
}
int mmqg_sample_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int R, int V,
                     unsigned long long seed, unsigned long long step, void* stream) {
  return sample_rows(logits, ldl, tokens, tok_stride, nullptr, R, V, seed, step, 0, R, as_stream(stream));
}
int mmqg_sample_uniform(float* out, int n, unsigned long long seed, unsigned long long step, void* stream) {
  if (!out || n <= 0) return set_err(MMQG_ERR_BAD_ARG, "sample_uniform: bad args");
  cudaStream_t st = as_stream(stream);
  sample_uniform_kernel<<<ceil_div(n, 256), 256, 0, st>>>(out, n, seed, step * (unsigned long long)n);
  MMQG_LAUNCH_CHECK();
  return 0;
}
int mmqg_colsum(const float* X, int ldx, float* out, int M, int N, float beta, void* stream) {
  return colsum(X, ldx, out, nullptr, M, N, beta, as_stream(stream));
}

}  // extern "C"

// ---- fused multi-tensor Adam (reference train.py:179-181, 265-267) -------------
// One launch updates every parameter of the model: parameters, gradients and both moments live in
// flat fp32 buffers with identical layout.  torch.optim.Adam semantics (no weight decay, no
// amsgrad): m += (1-b1)(g-m); v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
// Elements in [rep_lo, rep_hi) are stepped TWICE: the reference registers the shared embedding
// with two optimisers (train.py:236,245,255,266-267), whose moments are
// identical by construction, so the second update equals the first.
// The step count lives in device memory (state[0]) and is advanced by the prep kernel, so a
// captured CUDA graph replays correct bias corrections.  HBM-bound: 28 bytes per element.
namespace mmqg {
__global__ void adam_prep_kernel(float* state, float lr, float beta1, float beta2) {
  int t = __float_as_int(state[0]) + 1;
  state[0] = __int_as_float(t);
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  state[1] = (float)((double)lr / bc1);        // step size
  state[2] = (float)sqrt(bc2);                 // sqrt of the second bias correction
}
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, long long rep_lo, long long rep_hi,
                                                        const float* __restrict__ state, float beta1, float beta2, float eps) {
  const float step_size = state[1], bc2_sqrt = state[2];
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mm[k] = mm[k] + (1.0f - beta1) * (gg[k] - mm[k]);
      vv[k] = beta2 * vv[k] + (1.0f - beta2) * gg[k] * gg[k];
      const float upd = step_size * (mm[k] / (sqrtf(vv[k]) / bc2_sqrt + eps));
      pp[k] -= upd;
      const long long e = 4 * i + k;
      if (e >= rep_lo && e < rep_hi) pp[k] -= upd;
    }
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
  }
}
}  // namespace mmqg

extern "C" int mmqg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                              long long rep_lo, long long rep_hi, float lr, float beta1, float beta2, float eps,
                              float* state, void* stream) {
  using namespace mmqg;
  MMQG_REQUIRE(params && grads && exp_avg && exp_avg_sq && state && n > 0, "adam_step: null pointer");
  MMQG_REQUIRE(n % 4 == 0, "adam_step: n=%lld must be a multiple of 4 (pad the flat buffers)", n);
  MMQG_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  adam_prep_kernel<<<1, 1, 0, st>>>(state, lr, beta1, beta2);
  MMQG_LAUNCH_CHECK();
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;      // grid-stride: 16 resident CTAs of 256 threads on each of the 148 SMs
  MMQG_PROBE(KC_OTHER, 0, 28.0 * n);
  adam_step_kernel<<<(unsigned)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, rep_lo, rep_hi, state, beta1, beta2, eps);
  MMQG_LAUNCH_CHECK();
  return 0;
}


// Fused location-attention step (reference model/decoder.py:78-97): three softmaxes over
// the slot axis (warp-shuffle reductions) and the three context products, one CTA per
// sample, memory rows read with coalesced (128-bit when aligned) loads.  HBM/L2-bound:
// per step and sample it reads T_t*H + T_v*(H_a+H_v) memory elements once.
//
// Slot packing inside a score row: [ text TM | audio AM | video AM ] -- the order the
// reference returns the weights (decoder.py:107); context packing [c_txt|c_aud|c_vid]
// is the decoder-input order of decoder.py:99.
//
// Q1: the reference's length mask is a no-op, so each softmax runs over
// ALL TM / AM slots; only the context sums are bounded by T_t / T_v (padded rows are 0).
#include <cuda_bf16.h>
#include "kernels.h"

namespace mmqg {

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

template <bool VEC>
__global__ void __launch_bounds__(256) attn_fwd_kernel(float* __restrict__ scores, int lds,
                                                       const float* __restrict__ M_txt,
                                                       const float* __restrict__ M_aud,
                                                       const float* __restrict__ M_vid, float* __restrict__ ctx,
                                                       int ldctx, AttnShape s) {
  extern __shared__ float a[];          // S softmax weights
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = s.TM + 2 * s.AM;
  float* sc = scores + (size_t)b * lds;
  for (int j = tid; j < S; j += 256) a[j] = sc[j];
  __syncthreads();
  if (warp < 3) {
    const int off = warp == 0 ? 0 : (warp == 1 ? s.TM : s.TM + s.AM);
    const int len = warp == 0 ? s.TM : s.AM;
    float m = -INFINITY;
    for (int j = lane; j < len; j += 32) m = fmaxf(m, a[off + j]);
    m = wmax(m);
    float z = 0.f;
    for (int j = lane; j < len; j += 32) {
      float e = expf(a[off + j] - m);
      a[off + j] = e;
      z += e;
    }
    z = wsum(z);
    float inv = 1.0f / z;
    for (int j = lane; j < len; j += 32) {
      float p = a[off + j] * inv;
      a[off + j] = p;
      sc[off + j] = p;
    }
  }
  __syncthreads();
  const float* mt = M_txt + (size_t)b * s.TM * s.H;
  const float* ma = M_aud + (size_t)b * s.AM * s.H_a;
  const float* mv = M_vid + (size_t)b * s.AM * s.H_v;
  float* out = ctx + (size_t)b * ldctx;
  if (VEC) {
    const int q_t = s.H / 4, q_a = s.H_a / 4, q_v = s.H_v / 4;
    for (int o = tid; o < q_t + q_a + q_v; o += 256) {
      const float* base; const float* w; int n, ld, h4, dst;
      if (o < q_t) { base = mt; w = a; n = s.T_t; ld = s.H; h4 = o; dst = 4 * o; }
      else if (o < q_t + q_a) { base = ma; w = a + s.TM; n = s.T_v; ld = s.H_a; h4 = o - q_t; dst = s.H + 4 * h4; }
      else { base = mv; w = a + s.TM + s.AM; n = s.T_v; ld = s.H_v; h4 = o - q_t - q_a; dst = s.H + s.H_a + 4 * h4; }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        float4 v = *reinterpret_cast<const float4*>(base + (size_t)j * ld + 4 * h4);
        float p = w[j];
        acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y);
        acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
      }
      *reinterpret_cast<float4*>(out + dst) = acc;
      if (s.ctx16) {
        __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(s.ctx16) + (size_t)b * s.ldctx16 + dst;
        o16[0] = __float2bfloat16_rn(acc.x); o16[1] = __float2bfloat16_rn(acc.y);
        o16[2] = __float2bfloat16_rn(acc.z); o16[3] = __float2bfloat16_rn(acc.w);
      }
    }
  } else {
    for (int o = tid; o < s.H + s.H_a + s.H_v; o += 256) {
      const float* base; const float* w; int n, ld, h;
      if (o < s.H) { base = mt; w = a; n = s.T_t; ld = s.H; h = o; }
      else if (o < s.H + s.H_a) { base = ma; w = a + s.TM; n = s.T_v; ld = s.H_a; h = o - s.H; }
      else { base = mv; w = a + s.TM + s.AM; n = s.T_v; ld = s.H_v; h = o - s.H - s.H_a; }
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(w[j], base[(size_t)j * ld + h], acc);
      out[o] = acc;
      if (s.ctx16) reinterpret_cast<__nv_bfloat16*>(s.ctx16)[(size_t)b * s.ldctx16 + o] = __float2bfloat16_rn(acc);
    }
  }
}

__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* attn, float* ds_out, int lds,
                                                       const float* __restrict__ dctx, int lddctx,
                                                       const float* __restrict__ M_txt,
                                                       const float* __restrict__ M_aud,
                                                       const float* __restrict__ M_vid, float* __restrict__ dM_txt,
                                                       float* __restrict__ dM_vid, AttnShape s) {
  extern __shared__ float sm[];
  const int S = s.TM + 2 * s.AM, C = s.H + s.H_a + s.H_v;
  float* a = sm;            // S
  float* da = sm + S;       // S
  float* dc = sm + 2 * S;   // C
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* at = attn + (size_t)b * lds;
  float* ds = ds_out + (size_t)b * lds;
  for (int j = tid; j < S; j += 256) { a[j] = at[j]; da[j] = 0.f; }
  for (int o = tid; o < C; o += 256) {
    float v = dctx[(size_t)b * lddctx + o];
    for (int k = 1; k < s.dctx_parts; ++k) v += dctx[(size_t)k * s.dctx_part_stride + (size_t)b * lddctx + o];
    dc[o] = v;
    if (s.dctx_sum) s.dctx_sum[(size_t)b * s.lddsum + o] = v;
  }
  __syncthreads();
  const float* mt = M_txt + (size_t)b * s.TM * s.H;
  const float* ma = M_aud + (size_t)b * s.AM * s.H_a;
  const float* mv = M_vid + (size_t)b * s.AM * s.H_v;
  // da(j) = <dctx_head, M(b,j,:)> for the real rows; padded rows are zero so da stays 0.
  for (int j = warp; j < s.T_t + 2 * s.T_v; j += 8) {
    const float* row; const float* g; int n, slot;
    if (j < s.T_t) { row = mt + (size_t)j * s.H; g = dc; n = s.H; slot = j; }
    else if (j < s.T_t + s.T_v) { int k = j - s.T_t; row = ma + (size_t)k * s.H_a; g = dc + s.H; n = s.H_a; slot = s.TM + k; }
    else { int k = j - s.T_t - s.T_v; row = mv + (size_t)k * s.H_v; g = dc + s.H + s.H_a; n = s.H_v; slot = s.TM + s.AM + k; }
    float acc = 0.f;
    for (int h = lane; h < n; h += 32) acc = fmaf(g[h], row[h], acc);
    acc = wsum(acc);
    if (lane == 0) da[slot] = acc;
  }
  __syncthreads();
  if (warp < 3) {
    const int off = warp == 0 ? 0 : (warp == 1 ? s.TM : s.TM + s.AM);
    const int len = warp == 0 ? s.TM : s.AM;
    float dot = 0.f;
    for (int j = lane; j < len; j += 32) dot = fmaf(a[off + j], da[off + j], dot);
    dot = wsum(dot);
    for (int j = lane; j < len; j += 32) {
      float v = a[off + j] * (da[off + j] - dot);
      ds[off + j] = v;
      if (s.ds16) reinterpret_cast<__nv_bfloat16*>(s.ds16)[(size_t)b * s.ldds16 + off + j] = __float2bfloat16_rn(v);
    }
  }
  if (dM_txt) {
    float* d = dM_txt + (size_t)b * s.TM * s.H;
    for (int h = tid; h < s.H; h += 256) {
      float g = dc[h];
      for (int j = 0; j < s.T_t; ++j) d[(size_t)j * s.H + h] += a[j] * g;
    }
  }
  if (dM_vid) {
    float* d = dM_vid + (size_t)b * s.AM * s.H_v;
    for (int h = tid; h < s.H_v; h += 256) {
      float g = dc[s.H + s.H_a + h];
      for (int j = 0; j < s.T_v; ++j) d[(size_t)j * s.H_v + h] += a[s.TM + s.AM + j] * g;
    }
  }
}

// grid (B, ceil((H+H_v)/128)); block 256 = 128 columns x 2 row lanes.
__global__ void __launch_bounds__(256) attn_dmem_kernel(const float* __restrict__ attn_all, int lds,
                                                        const float* __restrict__ dctx_all, int lddctx,
                                                        float* __restrict__ dM_txt, float* __restrict__ dM_vid, int T_q,
                                                        AttnShape s) {
  extern __shared__ float sm[];
  const int NS = s.T_t + s.T_v;
  float* w = sm;                 // T_q x NS : attention weights of the real text / video rows
  float* g = sm + T_q * NS;      // T_q x 128: dctx column slice
  const int b = blockIdx.x, tid = threadIdx.x;
  const int col0 = blockIdx.y * 128;
  for (int i = tid; i < T_q * NS; i += 256) {
    int t = i / NS, j = i % NS;
    int slot = j < s.T_t ? j : s.TM + s.AM + (j - s.T_t);
    w[i] = attn_all[((size_t)t * s.B + b) * lds + slot];
  }
  for (int i = tid; i < T_q * 128; i += 256) {
    int t = i / 128, c = col0 + (i % 128);
    float v = 0.f;
    if (c < s.H + s.H_v) {
      int src = c < s.H ? c : s.H + s.H_a + (c - s.H);     // skip the audio part of dctx
      v = dctx_all[((size_t)t * s.B + b) * lddctx + src];
    }
    g[i] = v;
  }
  __syncthreads();
  const int cl = tid & 127, par = tid >> 7;
  const int c = col0 + cl;
  if (c >= s.H + s.H_v) return;
  const bool is_txt = c < s.H;
  // A 128-column slice may straddle the text/video boundary; each thread handles its own side.
  const int n = is_txt ? s.T_t : s.T_v;
  const int woff = is_txt ? 0 : s.T_t;
  float* dst = is_txt ? dM_txt + (size_t)b * s.TM * s.H + c : dM_vid + (size_t)b * s.AM * s.H_v + (c - s.H);
  const int ld = is_txt ? s.H : s.H_v;
  for (int j = par; j < n; j += 2) {
    float acc = 0.f;
    for (int t = 0; t < T_q; ++t) acc = fmaf(w[t * NS + woff + j], g[t * 128 + cl], acc);
    dst[(size_t)j * ld] = acc;
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int attn_fwd(float* scores, int lds, const float* M_txt, const float* M_aud, const float* M_vid, float* ctx, int ldctx,
             const AttnShape& s, cudaStream_t st) {
  MMQG_REQUIRE(scores && M_txt && M_aud && M_vid && ctx, "attn_fwd: null pointer");
  MMQG_REQUIRE(s.B > 0 && s.T_t <= s.TM && s.T_v <= s.AM, "attn_fwd: bad shape");
  if (s.m_txt16 && s.m_vid16 && attn_fast_ok(s, s.m_txt16, M_aud, s.m_vid16))
    return attn_fwd_fast(scores, lds, s.m_txt16, M_aud, s.m_vid16, true, ctx, ldctx, s, st);
  if (attn_fast_ok(s, M_txt, M_aud, M_vid)) return attn_fwd_fast(scores, lds, M_txt, M_aud, M_vid, false, ctx, ldctx, s, st);
  const int S = s.TM + 2 * s.AM;
  const size_t smem = (size_t)S * sizeof(float);
  MMQG_REQUIRE(smem <= 48 * 1024, "attn_fwd: %d attention slots exceed the 48 KB staging buffer", S);
  bool vec = s.H % 4 == 0 && s.H_a % 4 == 0 && s.H_v % 4 == 0 && ldctx % 4 == 0 && al16(M_txt) && al16(M_aud) &&
             al16(M_vid) && al16(ctx);
  MMQG_PROBE(KC_ATTN, 2.0 * s.B * ((double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v)),
             4.0 * s.B * ((double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v) + 2.0 * S + s.H + s.H_a + s.H_v));
  if (vec) attn_fwd_kernel<true><<<s.B, 256, smem, st>>>(scores, lds, M_txt, M_aud, M_vid, ctx, ldctx, s);
  else attn_fwd_kernel<false><<<s.B, 256, smem, st>>>(scores, lds, M_txt, M_aud, M_vid, ctx, ldctx, s);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int attn_bwd(const float* attn, float* ds_out, int lds, const float* dctx, int lddctx, const float* M_txt,
             const float* M_aud, const float* M_vid, float* dM_txt, float* dM_vid, const AttnShape& s,
             cudaStream_t st) {
  MMQG_REQUIRE(attn && ds_out && dctx && M_txt && M_aud && M_vid, "attn_bwd: null pointer");
  if (!dM_txt && !dM_vid) {
    if (s.m_txt16 && s.m_vid16 && attn_fast_ok(s, s.m_txt16, M_aud, s.m_vid16))
      return attn_bwd_fast(attn, ds_out, lds, dctx, lddctx, s.m_txt16, M_aud, s.m_vid16, true, s, st);
    if (attn_fast_ok(s, M_txt, M_aud, M_vid))
      return attn_bwd_fast(attn, ds_out, lds, dctx, lddctx, M_txt, M_aud, M_vid, false, s, st);
  }
  const int S = s.TM + 2 * s.AM, C = s.H + s.H_a + s.H_v;
  const size_t smem = (size_t)(2 * S + C) * sizeof(float);
  MMQG_REQUIRE(smem <= 48 * 1024, "attn_bwd: shape exceeds the 48 KB staging buffer");
  MMQG_PROBE(KC_ATTN, 2.0 * s.B * ((double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v)),
             4.0 * s.B * ((double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v) + 2.0 * S + C));
  attn_bwd_kernel<<<s.B, 256, smem, st>>>(attn, ds_out, lds, dctx, lddctx, M_txt, M_aud, M_vid, dM_txt, dM_vid, s);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int attn_dmem(const float* attn_all, int lds, const float* dctx_all, int lddctx, float* dM_txt, float* dM_vid, int T_q,
              const AttnShape& s, cudaStream_t st) {
  MMQG_REQUIRE(attn_all && dctx_all && dM_txt && dM_vid && T_q > 0, "attn_dmem: bad args");
  const size_t smem = (size_t)T_q * (s.T_t + s.T_v + 128) * sizeof(float);
  MMQG_REQUIRE(smem <= 200 * 1024, "attn_dmem: T_q*(T_t+T_v+128) floats exceed shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    MMQG_CUDA(cudaFuncSetAttribute(attn_dmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(s.B, ceil_div(s.H + s.H_v, 128));
  attn_dmem_kernel<<<grid, 256, smem, st>>>(attn_all, lds, dctx_all, lddctx, dM_txt, dM_vid, T_q, s);
  MMQG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmqg

using namespace mmqg;

extern "C" {

int mmqg_attn_fwd(float* scores, int lds, const float* M_txt, const float* M_aud, const float* M_vid, float* ctx,
                  int ldctx, int B, int TM, int AM, int H, int H_a, int H_v, int T_t, int T_v, void* stream) {
  AttnShape s{B, TM, AM, H, H_a, H_v, T_t, T_v};
  return attn_fwd(scores, lds, M_txt, M_aud, M_vid, ctx, ldctx, s, as_stream(stream));
}

int mmqg_attn_bwd(float* attn, int lds, const float* dctx, int lddctx, const float* M_txt, const float* M_aud,
                  const float* M_vid, float* dM_txt, float* dM_vid, int B, int TM, int AM, int H, int H_a, int H_v,
                  int T_t, int T_v, void* stream) {
  AttnShape s{B, TM, AM, H, H_a, H_v, T_t, T_v};
  return attn_bwd(attn, attn, lds, dctx, lddctx, M_txt, M_aud, M_vid, dM_txt, dM_vid, s, as_stream(stream));
}

}  // extern "C"

// Wide variants of the fused attention step (reference decoder.py:78-97) for H, H_v multiples
// of 8: 512 threads per sample, the memory rows split over 8 row groups so that every thread
// keeps several 16-byte loads in flight, and (bf16 mode) bf16 attention memories, which are
// bit-identical to the fp32 ones there (h is bf16-rounded already) at half the bytes.
// Per step and sample the kernels read T_t*H + T_v*(H_a+H_v) memory elements once: HBM/L2-bound.
#include <cuda_bf16.h>
#include "kernels.h"

namespace mmqg {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float fw_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float fw_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 consecutive elements of a memory row as floats
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

template <typename MT>
__global__ void __launch_bounds__(512) attn_fwd_fast_kernel(float* __restrict__ scores, int lds, const MT* __restrict__ M_txt,
                                                            const float* __restrict__ M_aud, const MT* __restrict__ M_vid,
                                                            float* __restrict__ ctx, int ldctx, AttnShape s) {
  extern __shared__ float sm[];
  const int S = s.TM + 2 * s.AM;
  const int Hmax = s.H > s.H_v ? s.H : s.H_v;
  float* a = sm;                 // S softmax weights
  float* part = sm + S;          // 8 x Hmax partial context sums
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();
  float* sc = scores + (size_t)b * lds;
  for (int j = tid; j < S; j += 512) a[j] = sc[j];
  __syncthreads();
  if (warp < 3) {
    const int off = warp == 0 ? 0 : (warp == 1 ? s.TM : s.TM + s.AM);
    const int len = warp == 0 ? s.TM : s.AM;
    float m = -INFINITY;
    for (int j = lane; j < len; j += 32) m = fmaxf(m, a[off + j]);
    m = fw_max(m);
    float z = 0.f;
    for (int j = lane; j < len; j += 32) {
      const float e = expf(a[off + j] - m);
      a[off + j] = e;
      z += e;
    }
    z = fw_sum(z);
    const float inv = 1.0f / z;
    for (int j = lane; j < len; j += 32) {
      const float p = a[off + j] * inv;
      a[off + j] = p;
      sc[off + j] = p;
    }
  }
  __syncthreads();
  float* out = ctx + (size_t)b * ldctx;
  bf16* out16 = s.ctx16 ? reinterpret_cast<bf16*>(s.ctx16) + (size_t)b * s.ldctx16 : nullptr;
  // ---- text and video contexts: 8 row groups x (H/8) column units ----
#pragma unroll 1
  for (int head = 0; head < 2; ++head) {
    const MT* base = head == 0 ? M_txt + (size_t)b * s.TM * s.H : M_vid + (size_t)b * s.AM * s.H_v;
    const float* w = head == 0 ? a : a + s.TM + s.AM;
    const int n = head == 0 ? s.T_t : s.T_v, Hh = head == 0 ? s.H : s.H_v;
    const int units = Hh / 8;
    const int dst0 = head == 0 ? 0 : s.H + s.H_a;
    for (int item = tid; item < units * 8; item += 512) {
      const int cu = item % units, rg = item / units;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      int j = rg;
      for (; j + 24 < n; j += 32) {                       // 4 rows in flight
        float v0[8], v1[8], v2[8], v3[8];
        load8(base + (size_t)j * Hh + 8 * cu, v0);
        load8(base + (size_t)(j + 8) * Hh + 8 * cu, v1);
        load8(base + (size_t)(j + 16) * Hh + 8 * cu, v2);
        load8(base + (size_t)(j + 24) * Hh + 8 * cu, v3);
        const float p0 = w[j], p1 = w[j + 8], p2 = w[j + 16], p3 = w[j + 24];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p3, v3[i], fmaf(p2, v2[i], fmaf(p1, v1[i], fmaf(p0, v0[i], acc[i]))));
      }
      for (; j < n; j += 8) {
        float v0[8];
        load8(base + (size_t)j * Hh + 8 * cu, v0);
        const float p0 = w[j];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p0, v0[i], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) part[rg * Hmax + 8 * cu + i] = acc[i];
    }
    __syncthreads();
    for (int h = tid; h < Hh; h += 512) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += part[r * Hmax + h];
      out[dst0 + h] = t;
      if (out16) out16[dst0 + h] = __float2bfloat16_rn(t);
    }
    __syncthreads();
  }
  // ---- audio context (fp32 features, T_v rows x H_a) ----
  {
    const float* ma = M_aud + (size_t)b * s.AM * s.H_a;
    const float* w = a + s.TM;
    for (int h = tid; h < s.H_a; h += 512) {
      float t = 0.f;
      for (int j = 0; j < s.T_v; ++j) t = fmaf(w[j], ma[(size_t)j * s.H_a + h], t);
      out[s.H + h] = t;
      if (out16) out16[s.H + h] = __float2bfloat16_rn(t);
    }
  }
}

template <typename MT>
__global__ void __launch_bounds__(512) attn_bwd_fast_kernel(const float* attn, float* ds_out, int lds,
                                                            const float* __restrict__ dctx, int lddctx,
                                                            const MT* __restrict__ M_txt, const float* __restrict__ M_aud,
                                                            const MT* __restrict__ M_vid, AttnShape s) {
  extern __shared__ float sm[];
  const int S = s.TM + 2 * s.AM, C = s.H + s.H_a + s.H_v;
  float* a = sm;
  float* da = sm + S;
  float* dc = sm + ((2 * S + 3) & ~3);      // 16-byte aligned: read as float4 below
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const float* at = attn + (size_t)b * lds;
  float* ds = ds_out + (size_t)b * lds;
  for (int j = tid; j < S; j += 512) { a[j] = at[j]; da[j] = 0.f; }
  for (int o = tid; o < C; o += 512) {
    float v = dctx[(size_t)b * lddctx + o];
    for (int k = 1; k < s.dctx_parts; ++k) v += dctx[(size_t)k * s.dctx_part_stride + (size_t)b * lddctx + o];
    dc[o] = v;
    if (s.dctx_sum) s.dctx_sum[(size_t)b * s.lddsum + o] = v;
  }
  __syncthreads();
  // da(j) = <dctx_head, M(b,j,:)> over the real rows.  Text rows: each of the 16 warps takes rows
  // j0, j0+16, j0+32, j0+48 together, so four independent 16-byte loads per lane are in flight
  // (the kernel is bound by load latency, not bytes: the rows of one sample are only ~100 KB).
  for (int j0 = warp; j0 < s.T_t; j0 += 64) {
    float acc4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int h = 8 * lane; h < s.H; h += 256) {
      float v[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = j0 + 16 * r;
        if (j < s.T_t) load8(M_txt + ((size_t)b * s.TM + j) * s.H + h, v[r]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
        }
      }
      const float4 d0 = *reinterpret_cast<const float4*>(dc + h), d1 = *reinterpret_cast<const float4*>(dc + h + 4);
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc4[r] = fmaf(dv[i], v[r][i], acc4[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float t = fw_sum(acc4[r]);
      if (lane == 0 && j0 + 16 * r < s.T_t) da[j0 + 16 * r] = t;
    }
  }
  // audio / video rows: one warp per row
  const int n_rows = s.T_t + 2 * s.T_v;
  for (int j = s.T_t + warp; j < n_rows; j += 16) {
    float acc = 0.f;
    int slot;
    if (j < s.T_t + s.T_v) {
      const int k = j - s.T_t;
      const float* row = M_aud + ((size_t)b * s.AM + k) * s.H_a;
      for (int h = lane; h < s.H_a; h += 32) acc = fmaf(dc[s.H + h], row[h], acc);
      slot = s.TM + k;
    } else {
      const int k = j - s.T_t - s.T_v;
      const MT* row = M_vid + ((size_t)b * s.AM + k) * s.H_v;
      for (int h = 8 * lane; h < s.H_v; h += 256) {
        float v[8];
        load8(row + h, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(dc[s.H + s.H_a + h + i], v[i], acc);
      }
      slot = s.TM + s.AM + k;
    }
    acc = fw_sum(acc);
    if (lane == 0) da[slot] = acc;
  }
  __syncthreads();
  if (warp < 3) {
    const int off = warp == 0 ? 0 : (warp == 1 ? s.TM : s.TM + s.AM);
    const int len = warp == 0 ? s.TM : s.AM;
    float dot = 0.f;
    for (int j = lane; j < len; j += 32) dot = fmaf(a[off + j], da[off + j], dot);
    dot = fw_sum(dot);
    bf16* ds16 = s.ds16 ? reinterpret_cast<bf16*>(s.ds16) + (size_t)b * s.ldds16 : nullptr;
    for (int j = lane; j < len; j += 32) {
      const float v = a[off + j] * (da[off + j] - dot);
      ds[off + j] = v;
      if (ds16) ds16[off + j] = __float2bfloat16_rn(v);
    }
  }
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool attn_fast_ok(const AttnShape& s, const void* M_txt, const void* M_aud, const void* M_vid) {
  // both kernels stage one sample in at most 48 KB of shared memory; larger shapes take the generic kernels
  const int S = s.TM + 2 * s.AM, Hmax = s.H > s.H_v ? s.H : s.H_v, C = s.H + s.H_a + s.H_v;
  const bool fits = (size_t)(S + 8 * Hmax) * sizeof(float) <= 48 * 1024 && (size_t)(2 * S + 4 + C) * sizeof(float) <= 48 * 1024;
  return fits && s.H % 8 == 0 && s.H_v % 8 == 0 && al16p(M_txt) && al16p(M_vid) && M_aud != nullptr;
}

int attn_fwd_fast(float* scores, int lds, const void* M_txt, const float* M_aud, const void* M_vid, bool mem_bf16, float* ctx,
                  int ldctx, const AttnShape& s, cudaStream_t st) {
  const int S = s.TM + 2 * s.AM, Hmax = s.H > s.H_v ? s.H : s.H_v;
  const size_t smem = (size_t)(S + 8 * Hmax) * sizeof(float);
  MMQG_REQUIRE(smem <= 48 * 1024, "attn_fwd_fast: shape exceeds the 48 KB staging buffer");
  const double esz = mem_bf16 ? 2.0 : 4.0;
  MMQG_PROBE(KC_ATTN, 2.0 * s.B * ((double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v)),
             s.B * (esz * ((double)s.T_t * s.H + (double)s.T_v * s.H_v) + 4.0 * s.T_v * s.H_a + 8.0 * S + 4.0 * (s.H + s.H_a + s.H_v)));
  if (mem_bf16)
    MMQG_CUDA(launch_k(attn_fwd_fast_kernel<bf16>, dim3(s.B), dim3(512), smem, st, scores, lds, reinterpret_cast<const bf16*>(M_txt),
                       M_aud, reinterpret_cast<const bf16*>(M_vid), ctx, ldctx, s));
  else
    MMQG_CUDA(launch_k(attn_fwd_fast_kernel<float>, dim3(s.B), dim3(512), smem, st, scores, lds, reinterpret_cast<const float*>(M_txt),
                       M_aud, reinterpret_cast<const float*>(M_vid), ctx, ldctx, s));
  MMQG_LAUNCH_CHECK();
  return 0;
}

int attn_bwd_fast(const float* attn, float* ds_out, int lds, const float* dctx, int lddctx, const void* M_txt, const float* M_aud,
                  const void* M_vid, bool mem_bf16, const AttnShape& s, cudaStream_t st) {
  const int S = s.TM + 2 * s.AM, C = s.H + s.H_a + s.H_v;
  const size_t smem = (size_t)(2 * S + 4 + C) * sizeof(float);
  MMQG_REQUIRE(smem <= 48 * 1024, "attn_bwd_fast: shape exceeds the 48 KB staging buffer");
  const double esz = mem_bf16 ? 2.0 : 4.0;
  MMQG_PROBE(KC_ATTN, 2.0 * s.B * ((double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v)),
             s.B * (esz * ((double)s.T_t * s.H + (double)s.T_v * s.H_v) + 4.0 * s.T_v * s.H_a + 8.0 * S + 4.0 * C));
  if (mem_bf16)
    MMQG_CUDA(launch_k(attn_bwd_fast_kernel<bf16>, dim3(s.B), dim3(512), smem, st, attn, ds_out, lds, dctx, lddctx,
                       reinterpret_cast<const bf16*>(M_txt), M_aud, reinterpret_cast<const bf16*>(M_vid), s));
  else
    MMQG_CUDA(launch_k(attn_bwd_fast_kernel<float>, dim3(s.B), dim3(512), smem, st, attn, ds_out, lds, dctx, lddctx,
                       reinterpret_cast<const float*>(M_txt), M_aud, reinterpret_cast<const float*>(M_vid), s));
  MMQG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmqg

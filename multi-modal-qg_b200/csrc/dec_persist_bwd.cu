// Persistent decoder-step kernel, backward through time: ONE cooperative launch runs the BPTT of a whole run of
// AttnDecoder steps (autograd twin of reference model/decoder.py:74-107 under train.py:177) instead of ~8 dependent
// launches per step (3 x (cell gradient + product), attention backward, query product).
//
// Decomposition (mirror of dec_persist.cu).  A group of ROWS samples is served by H/16 CTAs; CTA s owns the hidden
// units [16 s, 16 s + 16) of EVERY decoder layer -- i.e. 64 columns of each dG_l = d loss / d pre-activations --,
// 48 columns of the context gradient (s < ceil(C/48)) and ceil(ROWS / (H/16)) samples for the attention heads.
// Per step t (descending), warp-specialised:
//   tensor-core side (TMA warp + MMA warp, one smem ring, accumulators in tensor memory).  dG_l(t) is streamed ONCE
//   (ROWS x 4H bf16) and feeds two products at the same time, the weights arriving in the same ring stage:
//     S_l(t), l = L-1 .. 0 :  D_r[l]  = dG_l(t) . W_hh_l[:, units]         -> recurrent gradient of step t-1
//                             D_x[l-1] = dG_l(t) . W_ih_l[:, units]   (l>0) -> gradient w.r.t. the (dropped) h_{l-1}(t)
//                             D_ctx    = dG_0(t) . W_ih_l0[:, ctx columns]  -> d contexts (l = 0)
//     S_s(t)               :  D_r[L-1] += ds(t) . Wa_h[:, units]            -> query path of the attention Linears
//   worker side (4 warps):
//     cell l   : dh = D_r[l] (from step t+1) + mask . D_x[l] (+ d h_top from the loss head), LSTM cell gradient with
//                dc in registers, dG_l(t) published (bf16: also the operand of the hoisted weight-gradient products)
//     contexts : D_ctx -> fp32 rows of dctx_all (read by the hoisted memory-gradient kernel as well)
//     attention: per sample da_j = <dctx_head, M_j>, ds = a . (da - <a, da>) for the three heads, memory rows streamed
//                into shared memory by a loader thread (cp.async.bulk), warp-shuffle reductions
// Hand-over between the CTAs of a group: global memory (L2) + arrival counters, acquire polls, fence.proxy.async in
// front of TMA reads -- exactly as in the forward kernel.  All CTAs of a launch must be co-resident (cooperative).
#include <cuda_bf16.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "dec_common.cuh"
#include "dropout.cuh"

namespace mmqg {

using namespace tc;
typedef __nv_bfloat16 bf16;

namespace dpb {

using namespace dp;

static constexpr int NCTX = 48;                 // context-gradient columns per CTA (UMMA N = 48)

struct Maps {
  CUtensorMap dg[MAXL];     // dG_l: (T_q*B, 4H) bf16, box ROWS x 64
  CUtensorMap ds;           // d scores: (T_q*B, Sp) bf16, box ROWS x 64 (columns beyond Sp read as zero)
  // decoder weights, transposed (row = output column, K = gate index) and packed per slice so that ONE box per ring stage
  // brings both weight operands: layer l >= 1: rows [32 s, 32 s + 32) = 16 rows of W_hh_l^T | 16 rows of W_ih_l^T;
  // layer 0: rows [64 s, 64 s + 64) = 16 rows of W_hh_0^T | 48 rows of W_ih_l0[:, ctx]^T (zero rows beyond C)
  CUtensorMap wT[MAXL];     // box 32 x 64 (l >= 1) / 64 x 64 (l = 0)
  CUtensorMap w0T_rec;      // layer 0, slices without context columns: box 16 x 64 on the same tensor
  CUtensorMap wahT;         // attention Linears, state columns, transposed: (H, Spp) bf16, box 16 x 64
};

struct P {
  int t_lo, t_hi, Tq, B, H, C, Sp, TM, AM, T_t, T_v, H_a, H_v, L;
  int n_slices, n_ct, KBg, KBs, g0;
  const float* attn_all;           // (Tq*B, Sp) softmax weights of the forward pass
  float* ds_all; bf16* ds16;       // (Tq*B, Sp) out (padding columns stay as the caller zeroed them)
  float* dctx_all;                 // (Tq*B, C) out
  const float* dhtop;              // (Tq*B, H) d loss / d h_top from the loss head
  const float* acts[MAXL]; const float* cs[MAXL]; bf16* dg[MAXL];
  float* dc[MAXL];                 // (B, H): in (t_hi < Tq: state handed over by the launch of the later steps) / out
  float* dh_rec[MAXL];             // (B, H): in (t_hi < Tq) recurrent gradient for step t_hi-1 / out: for step t_lo-1
  const bf16* m_txt16; const bf16* m_vid16; const float* m_aud;
  uint32_t* flags; int fstride;    // per group: F_g[l][0..Tq) | F_c[0..Tq) | F_s[0..Tq)
  float drop_p; unsigned long long seed; const unsigned long long* ctr; int sid0;
  long long* trace;                // debug: 24 %globaltimer stamps per step written by CTA 0 (mmqg_debug_decb_trace), nullable
};

__device__ __forceinline__ uint32_t* flag_g(const P& p, int g, int l, int t) { return p.flags + (size_t)g * p.fstride + l * p.Tq + t; }
__device__ __forceinline__ uint32_t* flag_c(const P& p, int g, int t) { return p.flags + (size_t)g * p.fstride + MAXL * p.Tq + t; }
__device__ __forceinline__ uint32_t* flag_s(const P& p, int g, int t) { return p.flags + (size_t)g * p.fstride + (MAXL + 1) * p.Tq + t; }

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.z));
  const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.w));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// da[r] = <row r of a bf16 chunk in shared memory, d> for the rows of one chunk.  One worker warp per SM sub-partition:
// nothing hides latency but the warp's own ILP, so a warp takes 6 rows at a time, branch-free (row index clamped, result
// discarded), 12 independent 16-byte loads and FMA chains in flight, then six interleaved shuffle reductions.
// Lanes take 8 columns per 256; d holds the lane's columns in registers.
__device__ __forceinline__ void rows_dot_bf16(const uint8_t* sl, int nr, int ncols, const float (&d)[16], float* da_out, int warp,
                                              int lane) {
  const uint4* base = reinterpret_cast<const uint4*>(sl);
  const int pitch = ncols >> 3;                       // uint4 per row
  const bool second = 8 * lane + 256 < ncols;
  const bool first = 8 * lane < ncols;
  for (int rb = 0; rb < nr; rb += 24) {
    uint4 u[6][2];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int r = min(rb + warp + 4 * k, nr - 1);
      u[k][0] = first ? base[r * pitch + lane] : make_uint4(0u, 0u, 0u, 0u);
      u[k][1] = second ? base[r * pitch + lane + 32] : make_uint4(0u, 0u, 0u, 0u);
    }
    float acc[6][2];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float v[8];
        unpack8(u[k][i], v);
        float x = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) x = fmaf(v[e], d[8 * i + e], x);
        acc[k][i] = x;
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k][0] += acc[k][1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[k][0] += __shfl_xor_sync(0xffffffffu, acc[k][0], o);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int r = rb + warp + 4 * k;
      if (lane == k && r < nr) da_out[r] = acc[k][0];
    }
  }
}

template <int ROWS>
__global__ void __launch_bounds__(224, 1)
dec_seq_bwd_kernel(const __grid_constant__ Maps maps, const P p) {
  constexpr int A_BYTES = ROWS * 128, B_BYTES = 64 * 128, STAGE = A_BYTES + B_BYTES;
  // ONE 192 KB pool serves both the operand ring of the streams (16 / 24 KB stages) and the chunks of the attention
  // memories (24 KB slots): the two are never live at the same time (streams S_{L-1}..S_0 -> attention heads -> S_s and
  // the next step's streams), and each is bound by the bytes it can keep in flight (TMA round trip ~1.5 us at
  // ~100 GB/s per SM).  Hand-over: the loader waits for the last stream of the step (xdone[0]) before its first copy,
  // the TMA producer for the arrival counter of the d-scores (which this CTA bumps after its last sample).
  constexpr int POOL = 192 * 1024;
  constexpr int NSTAGE = POOL / STAGE;
  constexpr int ASTAGES = POOL / ASLOT;
  constexpr int EW = ROWS / 32;                  // epilogue warps (thread = batch row of the group)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* aring = smem;
  float* stg_all = reinterpret_cast<float*>(smem + POOL);
  float* a_sm = stg_all + 4 * STG_WARP;          // 2 x 512: softmax weights of the two samples in flight
  float* da_sm = a_sm + 1024;                    // 2 x 512: d loss / d (softmax output)
  float* dc_sm = da_sm + 1024;                   // 2 x 1536: d context
  __shared__ uint64_t full[NSTAGE], empty[NSTAGE], xdone[MAXL], topdone, afull[ASTAGES], aempty[ASTAGES];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = p.g0 + blockIdx.x / p.n_slices, s = blockIdx.x % p.n_slices;      // group, unit slice
  const int B = p.B, H = p.H, G = 4 * p.H, L = p.L, C = p.C;
  const int row_g0 = g * ROWS;
  const int rows_grp = min(ROWS, B - row_g0);
  const bool has_ctx = s < p.n_ct;
  const bool drop = p.drop_p > 0.f;
  const bool cont = p.t_hi < p.Tq;               // state handed over from the launch of the later steps

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < MAXL; ++i) mbar_init(&xdone[i], 1);
    mbar_init(&topdone, 1);
    for (int i = 0; i < ASTAGES; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  // Accumulator block of stream S_l: columns [0, 16) = D_r[l], [16, 32) = D_x[l-1] (l >= 1) or [16, 64) = D_ctx (l = 0), so
  // that ONE MMA per k-step (N = 32 / 64) reads the A tile from shared memory once for both products -- with N = 16 the
  // tensor core is bound by the 4 KB A-tile fetch of every instruction, not by its math.
  auto blk = [](int l) -> uint32_t { return l == 0 ? 0u : 64u + 32u * (uint32_t)(l - 1); };

  if (warp == 4) {
    // ---------------- TMA producer of the tensor-core side ----------------
    if (elect_one()) {
      int i = 0;
      long long* trp = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
      // One stream: nkb ring stages of [A box | weight boxes].  The weights do not depend on the arrival counter, so
      // those of the first stages are issued BEFORE the poll (their round trip overlaps the wait), the A boxes after it.
      auto stream = [&](const uint32_t* flag, uint32_t target, const CUtensorMap* mA, int rowA, int nkb, const CUtensorMap* mW,
                        int rowW, uint32_t w_bytes, bool early_w, long long* stamp) {
        const uint32_t bytes = A_BYTES + w_bytes;
        const int pre = !early_w ? 0 : (nkb < NSTAGE ? nkb : NSTAGE);
        const int i0 = i;
        for (int kb = 0; kb < pre; ++kb) {
          const int st = (i0 + kb) % NSTAGE, ph = ((i0 + kb) / NSTAGE) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect_tx(&full[st], bytes);
          tma_load_2d(ring + st * STAGE + A_BYTES, mW, &full[st], kb * 64, rowW);
        }
        wait_count(flag, target);
        fence_proxy_async();
        if (stamp) *stamp = gtime();
        for (int kb = 0; kb < pre; ++kb) tma_load_2d(ring + ((i0 + kb) % NSTAGE) * STAGE, mA, &full[(i0 + kb) % NSTAGE], kb * 64, rowA);
        i = i0 + pre;
        for (int kb = pre; kb < nkb; ++kb, ++i) {
          const int st = i % NSTAGE, ph = (i / NSTAGE) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect_tx(&full[st], bytes);
          uint8_t* a = ring + st * STAGE;
          tma_load_2d(a, mA, &full[st], kb * 64, rowA);
          tma_load_2d(a + A_BYTES, mW, &full[st], kb * 64, rowW);
        }
      };
      for (int t = p.t_hi - 1; t >= p.t_lo; --t) {
        const int rowA = t * B + row_g0;
        for (int l = L - 1; l >= 0; --l) {
          const CUtensorMap* mW = l > 0 ? &maps.wT[l] : (has_ctx ? &maps.wT[0] : &maps.w0T_rec);
          stream(flag_g(p, g, l, t), (uint32_t)p.n_slices, &maps.dg[l], rowA, p.KBg, mW, l > 0 ? s * 32 : s * 64,
                 l > 0 ? 4096u : (has_ctx ? 8192u : 2048u), true, (trp && l == L - 1) ? &trp[t * 24 + 12] : nullptr);
          if (trp && l == L - 1) trp[t * 24 + 13] = gtime();
        }
        stream(flag_s(p, g, t), (uint32_t)rows_grp, &maps.ds, rowA, p.KBs, &maps.wahT, s * 16, 2048u, false, trp ? &trp[t * 24 + 14] : nullptr);
        if (trp) trp[t * 24 + 15] = gtime();
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer (same job list) ----------------
    if (elect_one()) {
      constexpr int MM = ROWS == 64 ? 128 : ROWS;      // 64-row groups run M = 128 with stale upper rows
      constexpr uint32_t idesc16 = umma_idesc_bf16(MM, 16, 0, 0), idesc32 = umma_idesc_bf16(MM, 32, 0, 0),
                         idesc64 = umma_idesc_bf16(MM, 16 + NCTX, 0, 0);
      int i = 0;
      for (int t = p.t_hi - 1; t >= p.t_lo; --t) {
        for (int l = L - 1; l >= 0; --l) {
          const uint32_t d_blk = tmem_base + blk(l);
          const uint32_t idesc = l > 0 ? idesc32 : (has_ctx ? idesc64 : idesc16);
          for (int kb = 0; kb < p.KBg; ++kb, ++i) {
            const int st = i % NSTAGE, ph = (i / NSTAGE) & 1;
            mbar_wait(&full[st], ph);
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(ring + st * STAGE), b_addr = a_addr + A_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_blk, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[st]);
          }
          umma_commit(&xdone[l]);
        }
        for (int kb = 0; kb < p.KBs; ++kb, ++i) {
          const int st = i % NSTAGE, ph = (i / NSTAGE) & 1;
          mbar_wait(&full[st], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(ring + st * STAGE), b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + blk(L - 1), umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc16, 1u);
          umma_commit(&empty[st]);
        }
        umma_commit(&topdone);
      }
    }
  } else if (warp == 6) {
    // ---------------- loader of the attention memories ----------------
    if (elect_one()) {
      const Chunker ck(p.H, p.H_a, p.H_v, p.T_t, p.T_v);
      int i = 0, it = 0;
      for (int t = p.t_hi - 1; t >= p.t_lo; --t, ++it) {
        mbar_wait(&xdone[0], it & 1);            // the pool is free: every stream stage of this step has been consumed
        for (int b = row_g0 + s; b < row_g0 + rows_grp; b += p.n_slices) {
          for (int c = 0; c < ck.count(); ++c, ++i) {
            const int st = i % ASTAGES, ph = (i / ASTAGES) & 1;
            mbar_wait(&aempty[st], ph ^ 1);
            const void* src;
            uint32_t bytes;
            if (c < ck.n_t) {
              const int r0 = c * ck.cr_t, nr = min(ck.cr_t, p.T_t - r0);
              src = p.m_txt16 + ((size_t)b * p.TM + r0) * H; bytes = (uint32_t)nr * H * 2;
            } else if (c < ck.n_t + ck.n_a) {
              const int r0 = (c - ck.n_t) * ck.cr_a, nr = min(ck.cr_a, p.T_v - r0);
              src = p.m_aud + ((size_t)b * p.AM + r0) * p.H_a; bytes = (uint32_t)nr * p.H_a * 4;
            } else {
              const int r0 = (c - ck.n_t - ck.n_a) * ck.cr_v, nr = min(ck.cr_v, p.T_v - r0);
              src = p.m_vid16 + ((size_t)b * p.AM + r0) * p.H_v; bytes = (uint32_t)nr * p.H_v * 2;
            }
            mbar_expect_tx(&afull[st], bytes);
            bulk_g2s(aring + st * ASLOT, src, bytes, &afull[st]);
          }
        }
      }
    }
  } else {
    // ---------------- workers: cell gradients, context rows, attention heads ----------------
    const int tid = threadIdx.x;                     // 0..127
    const bool epi = warp < EW;
    const int m0w = row_g0 + warp * 32;
    const int m = m0w + lane;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = s * 16;
    float* stg = stg_all + warp * STG_WARP;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(32 * warp) << 16);
    const Chunker ck(p.H, p.H_a, p.H_v, p.T_t, p.T_v);
    const int off_a = p.TM, off_v = p.TM + p.AM;
    const float ik = drop ? 1.0f / (1.0f - p.drop_p) : 1.f;
    const unsigned long long sd = p.seed + ((drop && p.ctr) ? *p.ctr : 0ull);
    float dcv[MAXL][16];
#pragma unroll
    for (int l = 0; l < MAXL; ++l) {
#pragma unroll
      for (int u = 0; u < 16; ++u) dcv[l][u] = 0.f;
    }
    if (epi && cont) {
#pragma unroll
      for (int l = 0; l < MAXL; ++l)
        if (l < L) {
          float4 c4[4];
          coop_ldg(p.dc[l] + (size_t)m0w * H + j0, H, rows_valid, lane, c4);
          coop_to_row(stg, lane, c4, dcv[l]);
        }
    }
    int ai = 0, it = 0;
    long long* tr = (p.trace && blockIdx.x == 0 && tid == 0) ? p.trace : nullptr;
    for (int t = p.t_hi - 1; t >= p.t_lo; --t, ++it) {
      if (tr) tr[t * 24 + 0] = gtime();
      if (epi) {
#pragma unroll
        for (int l = MAXL - 1; l >= 0; --l) {
          if (l >= L) continue;
          // operands of this layer's cell gradient (independent of the products in flight)
          float a[64], cn[16], cp[16], dh[16];
          {
            const float* abase = p.acts[l] + ((size_t)t * B + m0w) * G + j0;
            float4 a4[4][4], cn4[4], cp4[4], t4[4];      // all loads in flight before the first transpose
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) coop_ldg(abase + gg * H, G, rows_valid, lane, a4[gg]);
            coop_ldg(p.cs[l] + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid, lane, cn4);
            coop_ldg(p.cs[l] + ((size_t)t * B + m0w) * H + j0, H, rows_valid, lane, cp4);
            if (l == L - 1) coop_ldg(p.dhtop + ((size_t)t * B + m0w) * H + j0, H, rows_valid, lane, t4);
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) coop_to_row(stg, lane, a4[gg], a + gg * 16);
            coop_to_row(stg, lane, cn4, cn);
            coop_to_row(stg, lane, cp4, cp);
            if (l == L - 1) {
              coop_to_row(stg, lane, t4, dh);
            } else {
#pragma unroll
              for (int u = 0; u < 16; ++u) dh[u] = 0.f;
            }
            if (it == 0 && cont) {
              float r[16];
              coop_ldg(p.dh_rec[l] + (size_t)m0w * H + j0, H, rows_valid, lane, t4);
              coop_to_row(stg, lane, t4, r);
#pragma unroll
              for (int u = 0; u < 16; ++u) dh[u] += r[u];
            }
          }
          if (tr && l == L - 1) tr[t * 24 + 1] = gtime();
          if (l == L - 1) {
            if (it > 0) {
              mbar_wait(&topdone, (it - 1) & 1);
              tc_fence_after_sync();
              float r[16];
              tmem_ld_32x16(tbase + blk(l), r);
              tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; ++u) dh[u] += r[u];
            }
          } else {
            mbar_wait(&xdone[l + 1], it & 1);
            tc_fence_after_sync();
            float x[16], r[16];
            tmem_ld_32x16(tbase + blk(l + 1) + 16, x);
            if (it > 0) tmem_ld_32x16(tbase + blk(l), r);
            tmem_ld_wait();
            if (drop) {       // x is d/d(dropped h_l(t)): back through the mask of layer l's output
              const unsigned long long e0 = ((unsigned long long)t * B + m) * H + j0;
#pragma unroll
              for (int u = 0; u < 16; ++u) x[u] *= drop_scale(sd, p.sid0 + l, e0 + u, p.drop_p, ik);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) dh[u] += x[u] + (it > 0 ? r[u] : 0.f);
          }
          tc_fence_before_sync();
          if (tr) tr[t * 24 + 2 + 2 * (L - 1 - l)] = gtime();
          uint32_t w8[4][8];
#pragma unroll
          for (int u = 0; u < 16; u += 2) {
            float dgv[4][2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float ig = a[u + e], fg = a[16 + u + e], gt = a[32 + u + e], og = a[48 + u + e];
              const float d = dh[u + e];
              const float tcn = tanh_fast(cn[u + e]);
              const float dct = dcv[l][u + e] + d * og * (1.f - tcn * tcn);
              dgv[0][e] = dct * gt * ig * (1.f - ig);
              dgv[1][e] = dct * cp[u + e] * fg * (1.f - fg);
              dgv[2][e] = dct * ig * (1.f - gt * gt);
              dgv[3][e] = d * tcn * og * (1.f - og);
              dcv[l][u + e] = dct * fg;
            }
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              __nv_bfloat162 t2 = __floats2bfloat162_rn(dgv[gg][0], dgv[gg][1]);
              w8[gg][u >> 1] = *reinterpret_cast<uint32_t*>(&t2);
            }
          }
          bf16* dbase = p.dg[l] + ((size_t)t * B + m0w) * G + j0;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, w8[gg], dbase + gg * H, G, rows_valid);
          bar_epi(ROWS);
          if (tid == 0) {
            __threadfence();
            red_relaxed_gpu_add(flag_g(p, g, l, t), 1u);
            if (tr) tr[t * 24 + 3 + 2 * (L - 1 - l)] = gtime();
          }
        }
        // ---- context gradient of this CTA's 48 columns ----
        mbar_wait(&xdone[0], it & 1);          // every epilogue thread observes it: D_r[0] is read at the next step
        if (tr) tr[t * 24 + 8] = gtime();
        if (has_ctx) {
          tc_fence_after_sync();
          float cx[NCTX];
#pragma unroll
          for (int q = 0; q < 3; ++q) tmem_ld_32x16(tbase + blk(0) + 16 + 16 * q, cx + 16 * q);
          tmem_ld_wait();
          tc_fence_before_sync();
          float* cbase = p.dctx_all + ((size_t)t * B + m0w) * C + s * NCTX;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float4 tmp[4];
            row_to_coop(stg, lane, cx + 16 * q, tmp);
            if (s * NCTX + 16 * q + 4 * (lane & 3) < C) coop_stg(cbase + 16 * q, C, rows_valid, lane, tmp);
          }
          bar_epi(ROWS);
          if (tid == 0) {
            __threadfence();
            red_relaxed_gpu_add(flag_c(p, g, t), 1u);
            if (tr) tr[t * 24 + 9] = gtime();
          }
        }
      }
      // ---- attention heads of this CTA's samples, backward ----
      if (tid == 0) wait_count(flag_c(p, g, t), (uint32_t)p.n_ct);
      if (tr) tr[t * 24 + 10] = gtime();
      bar_workers();
      int n_done = 0;
      for (int b0 = row_g0 + s; b0 < row_g0 + rows_grp; b0 += 2 * p.n_slices) {
        // operands of up to two samples in one trip to L2
        const int nb = (b0 + p.n_slices < row_g0 + rows_grp) ? 2 : 1;
        for (int q = 0; q < nb; ++q) {
          const int b = b0 + q * p.n_slices;
          const float* arow = p.attn_all + ((size_t)t * B + b) * p.Sp;
          const float* drow = p.dctx_all + ((size_t)t * B + b) * C;
          for (int j = tid; j < p.Sp; j += 128) { a_sm[q * 512 + j] = __ldcg(arow + j); da_sm[q * 512 + j] = 0.f; }
          for (int j = tid; j < C; j += 128) dc_sm[q * 1536 + j] = __ldcg(drow + j);
        }
        bar_workers();
        if (tr && n_done == 0) tr[t * 24 + 16] = gtime();
        for (int q = 0; q < nb; ++q, ++n_done) {
          const int b = b0 + q * p.n_slices;
          const float* as = a_sm + q * 512;
          float* das = da_sm + q * 512;
          const float* dcs = dc_sm + q * 1536;
          const bool tr0 = tr && n_done == 0;
          float dt[16], dv[16];      // this lane's columns of the text / video head's d context
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c0 = 8 * lane + 256 * i;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              dt[8 * i + k] = c0 < H ? dcs[c0 + k] : 0.f;
              dv[8 * i + k] = c0 < p.H_v ? dcs[H + p.H_a + c0 + k] : 0.f;
            }
          }
          for (int c = 0; c < ck.count(); ++c, ++ai) {
            const int st = ai % ASTAGES, ph = (ai / ASTAGES) & 1;
            mbar_wait(&afull[st], ph);
            const uint8_t* sl = aring + st * ASLOT;
            if (c < ck.n_t) {
              const int r0 = c * ck.cr_t, nr = min(ck.cr_t, p.T_t - r0);
              rows_dot_bf16(sl, nr, H, dt, das + r0, warp, lane);
            } else if (c < ck.n_t + ck.n_a) {
              const int r0 = (c - ck.n_t) * ck.cr_a, nr = min(ck.cr_a, p.T_v - r0);
              for (int r = warp; r < nr; r += 4) {
                float acc = 0.f;
                for (int c0 = 4 * lane; c0 < p.H_a; c0 += 128) {
                  const float4 v = *reinterpret_cast<const float4*>(sl + ((size_t)r * p.H_a + c0) * 4);
                  const float4 d4 = *reinterpret_cast<const float4*>(dcs + H + c0);
                  acc = fmaf(v.x, d4.x, acc); acc = fmaf(v.y, d4.y, acc); acc = fmaf(v.z, d4.z, acc); acc = fmaf(v.w, d4.w, acc);
                }
                acc = wsum(acc);
                if (lane == 0) das[off_a + r0 + r] = acc;
              }
            } else {
              const int r0 = (c - ck.n_t - ck.n_a) * ck.cr_v, nr = min(ck.cr_v, p.T_v - r0);
              rows_dot_bf16(sl, nr, p.H_v, dv, das + off_v + r0, warp, lane);
            }
            mbar_arrive(&aempty[st]);
          }
          if (tr0) tr[t * 24 + 17] = gtime();
          bar_workers();                                // d (softmax output) of this sample is complete
          // softmax backward per head (all slots: the reference's length mask is a no-op); the fourth warp
          // goes straight on to the next sample, which has its own buffers
          if (warp < 3) {
            const int off = warp == 0 ? 0 : (warp == 1 ? off_a : off_v);
            const int len = warp == 0 ? p.TM : p.AM;
            float dot = 0.f;
            for (int j = lane; j < len; j += 32) dot = fmaf(as[off + j], das[off + j], dot);
            dot = wsum(dot);
            float* dsr = p.ds_all + ((size_t)t * B + b) * p.Sp;
            bf16* ds16r = p.ds16 + ((size_t)t * B + b) * p.Sp;
            for (int j = lane; j < len; j += 32) {
              const float v = as[off + j] * (das[off + j] - dot);
              dsr[off + j] = v;
              ds16r[off + j] = __float2bfloat16_rn(v);
            }
          }
          if (tr0) tr[t * 24 + 20] = gtime();
        }
        bar_workers();                                  // the buffers are rewritten by the next round; covers the ds stores
      }
      if (tid == 0 && n_done > 0) {
        __threadfence();
        red_relaxed_gpu_add(flag_s(p, g, t), (uint32_t)n_done);
      }
      if (tr) tr[t * 24 + 11] = gtime();
    }
    // ---- hand-over: recurrent gradients for step t_lo - 1 (the encoder's final state when t_lo = 0) and d c ----
    if (epi && it > 0) {
      mbar_wait(&topdone, (it - 1) & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int l = 0; l < MAXL; ++l) {
        if (l >= L) continue;
        float r[16];
        tmem_ld_32x16(tbase + blk(l), r);
        tmem_ld_wait();
        float4 tmp[4];
        row_to_coop(stg, lane, r, tmp);
        coop_stg(p.dh_rec[l] + (size_t)m0w * H + j0, H, rows_valid, lane, tmp);
        row_to_coop(stg, lane, dcv[l], tmp);
        coop_stg(p.dc[l] + (size_t)m0w * H + j0, H, rows_valid, lane, tmp);
      }
      tc_fence_before_sync();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}

// dst[c * ld_dst + r] = bf16(src[r * ld_src + c]) for r < R, c < Cc (32 x 32 tiles through shared memory)
__global__ void transpose_f32_bf16_kernel(const float* __restrict__ src, long long ld_src, int R, int Cc, bf16* __restrict__ dst,
                                          long long ld_dst) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? src[(size_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) dst[(size_t)c * ld_dst + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// Packed transposed weights of one decoder layer for the BPTT kernel: out row (slice * (16 + n_in) + r) holds, over
// k = gate row 0 .. 4H-1,  W_hh[k][16 slice + r] for r < 16 and W_in[k][n_in slice + (r - 16)] for r >= 16 (zero when
// that column is >= n_in_total).  Tiles of 32 out rows x 32 k through shared memory; both sides 64-byte coalesced or better.
__global__ void pack_decb_weights_kernel(const float* __restrict__ w_hh, int H, const float* __restrict__ w_in, long long ld_in,
                                         int n_in, int n_in_total, bf16* __restrict__ out, int n_rows) {
  __shared__ float tile[32][33];
  const int G = 4 * H, per = 16 + n_in;
  const int row0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int row = row0 + threadIdx.x;
  const int sl = row / per, r = row % per;
  const float* src = nullptr;
  long long ld = 0;
  if (row < n_rows) {
    if (r < 16) { src = w_hh + 16 * sl + r; ld = H; }
    else if (n_in * sl + (r - 16) < n_in_total) { src = w_in + n_in * sl + (r - 16); ld = ld_in; }
  }
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = (src && k0 + i < G) ? src[(size_t)(k0 + i) * ld] : 0.f;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    if (row0 + i < n_rows && k0 + threadIdx.x < G) out[(size_t)(row0 + i) * G + k0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x][i]);
}

}  // namespace dpb

// ---- host ---------------------------------------------------------------------------------------
int transpose_f32_bf16(const float* src, long long ld_src, int R, int Cc, void* dst, long long ld_dst, cudaStream_t st) {
  MMQG_REQUIRE(src && dst && R > 0 && Cc > 0, "transpose_f32_bf16: bad args");
  dpb::transpose_f32_bf16_kernel<<<dim3(ceil_div(R, 32), ceil_div(Cc, 32)), dim3(32, 8), 0, st>>>(src, ld_src, R, Cc,
                                                                                                 reinterpret_cast<bf16*>(dst), ld_dst);
  MMQG_LAUNCH_CHECK();
  return 0;
}

// rows of the packed weight tensor of layer l (see pack_decb_weights_kernel): (H/16) * (16 + n_in), n_in = 16 (l >= 1) / 48 (l = 0)
size_t dec_bwd_weight_rows(int H, int l) { return (size_t)(H / 16) * (l == 0 ? 16 + dpb::NCTX : 32); }
int pack_decb_weights(const float* w_hh, int H, const float* w_in, long long ld_in, int n_in_total, int l, void* out, cudaStream_t st) {
  MMQG_REQUIRE(w_hh && w_in && out && H % 16 == 0, "pack_decb_weights: bad args");
  const int n_rows = (int)dec_bwd_weight_rows(H, l);
  dpb::pack_decb_weights_kernel<<<dim3(ceil_div(n_rows, 32), ceil_div(4 * H, 32)), dim3(32, 8), 0, st>>>(
      w_hh, H, w_in, ld_in, l == 0 ? dpb::NCTX : 16, n_in_total, reinterpret_cast<bf16*>(out), n_rows);
  MMQG_LAUNCH_CHECK();
  return 0;
}

static long long* g_decb_trace = nullptr;      // debug hook, see mmqg_debug_decb_trace()
static int decb_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// same group geometry as the forward kernel (dec_persist_rows); additionally the context columns must fit the slices
bool dec_bwd_persist_ok(const DecPersistShape& s) {
  if (!dec_persist_ok(s)) return false;
  if (ceil_div(s.C, dpb::NCTX) > s.H / 16 || s.C > 1536 || s.H_a % 4 != 0) return false;
  return true;
}

size_t dec_bwd_persist_flag_words(const DecPersistShape& s, int Tq) {
  const int rows = dec_persist_rows(s.B, s.H);
  if (!rows) return 0;
  return (size_t)ceil_div(s.B, rows) * (dp::MAXL + 2) * Tq;
}

template <int ROWS>
static int launch_decb(const dpb::Maps& maps, const dpb::P& p, int n_grp, cudaStream_t st) {
  const size_t smem = 192 * 1024 + 4 * dp::STG_WARP * sizeof(float) + 2 * (512 + 512 + 1536) * sizeof(float) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(dpb::dec_seq_bwd_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_grp * p.n_slices);
  cfg.blockDim = dim3(224);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  MMQG_CUDA(cudaLaunchKernelEx(&cfg, dpb::dec_seq_bwd_kernel<ROWS>, maps, p));
  return 0;
}

// BPTT of the decoder steps [t_lo, t_hi), last step first.  t_hi < Tq continues from the state (dc, dh_rec) an earlier
// call over the later steps left behind; the arrival counters must have been zeroed before the first call of a sequence.
int dec_seq_bwd_persist(const DecPersistBwdArgs& a, int t_lo, int t_hi, cudaStream_t st) {
  const DecPersistShape& s = a.shape;
  MMQG_REQUIRE(dec_bwd_persist_ok(s), "dec_seq_bwd_persist: shape not supported");
  MMQG_REQUIRE(t_lo >= 0 && t_lo < t_hi && t_hi <= a.Tq, "dec_seq_bwd_persist: steps [%d, %d) outside [0, %d)", t_lo, t_hi, a.Tq);
  const int rows = dec_persist_rows(s.B, s.H);
  const int n_grp = ceil_div(s.B, rows);
  dpb::P p{};
  p.t_lo = t_lo; p.t_hi = t_hi; p.Tq = a.Tq; p.B = s.B; p.H = s.H; p.C = s.C; p.Sp = s.Sp; p.TM = s.TM; p.AM = s.AM;
  p.T_t = s.T_t; p.T_v = s.T_v; p.H_a = s.H_a; p.H_v = s.H_v; p.L = s.L;
  p.n_slices = s.H / 16; p.n_ct = ceil_div(s.C, dpb::NCTX); p.KBg = 4 * s.H / 64; p.KBs = ceil_div(s.Sp, 64);
  p.attn_all = a.attn_all; p.ds_all = a.ds_all; p.ds16 = reinterpret_cast<bf16*>(a.ds16); p.dctx_all = a.dctx_all; p.dhtop = a.dhtop;
  p.m_txt16 = reinterpret_cast<const bf16*>(a.m_txt16); p.m_vid16 = reinterpret_cast<const bf16*>(a.m_vid16); p.m_aud = a.m_aud;
  p.flags = a.flags; p.fstride = (dp::MAXL + 2) * a.Tq;
  p.drop_p = a.drop_p; p.seed = a.seed; p.ctr = a.ctr; p.sid0 = a.sid0;
  p.trace = g_decb_trace;
  dpb::Maps maps;
  const uint64_t rows_step = (uint64_t)a.Tq * s.B, G = 4 * (uint64_t)s.H;
  for (int l = 0; l < dp::MAXL; ++l) {
    const int ll = l < s.L ? l : 0;
    p.acts[l] = a.acts[ll]; p.cs[l] = a.cs[ll]; p.dg[l] = reinterpret_cast<bf16*>(a.dg[ll]); p.dc[l] = a.dc[ll]; p.dh_rec[l] = a.dh_rec[ll];
    MMQG_TRY(make_tmap_bf16_2d(&maps.dg[l], a.dg[ll], rows_step, G, G, rows, 64));
    if (ll == 0) MMQG_TRY(make_tmap_bf16_2d(&maps.wT[l], a.wT[0], 64 * (uint64_t)p.n_slices, G, G, 64, 64));
    else MMQG_TRY(make_tmap_bf16_2d(&maps.wT[l], a.wT[ll], 32 * (uint64_t)p.n_slices, G, G, 32, 64));
  }
  MMQG_TRY(make_tmap_bf16_2d(&maps.w0T_rec, a.wT[0], 64 * (uint64_t)p.n_slices, G, G, 16, 64));
  MMQG_TRY(make_tmap_bf16_2d(&maps.ds, a.ds16, rows_step, s.Sp, s.Sp, rows, 64));
  MMQG_TRY(make_tmap_bf16_2d(&maps.wahT, a.wahT, s.H, a.Spp, a.Spp, 16, 64));
  const int T = t_hi - t_lo;
  const double fl = 2.0 * T * s.B * (4.0 * s.H * ((double)s.C + s.H + (s.L - 1) * 2.0 * s.H) + (double)s.Sp * s.H +
                                     (double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v));
  const int per_wave = decb_sms() / p.n_slices;
  for (int g0 = 0; g0 < n_grp; g0 += per_wave) {
    p.g0 = g0;
    const int ng = n_grp - g0 < per_wave ? n_grp - g0 : per_wave;
    MMQG_PROBE(KC_GEMM_STEP, fl * ng / n_grp, 0);
    if (rows == 64) MMQG_TRY(launch_decb<64>(maps, p, ng, st));
    else MMQG_TRY(launch_decb<128>(maps, p, ng, st));
    MMQG_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace mmqg

// Debug hook (not part of the product path): device buffer of 24 * T_q int64 that CTA 0 of the next persistent decoder
// BPTT launches fills with %globaltimer stamps per step (see tools/decb_trace.py for the columns).  NULL switches it off.
extern "C" void mmqg_debug_decb_trace(void* dev_buf) { mmqg::g_decb_trace = reinterpret_cast<long long*>(dev_buf); }

// Persistent decoder-step kernel, forward (teacher forcing): ONE cooperative launch runs a whole run of
// AttnDecoder steps (reference model/decoder.py:74-107 driven by train.py:171-175) instead of ~8 dependent
// launches per step (score product, attention, 3 x (product + cell update)).
//
// Decomposition.  The batch is cut into groups of ROWS samples (64 or 128); a group is served by H/16 CTAs
// (32 at H = 512), CTA s owning the 16 hidden units [16 s, 16 s + 16) of EVERY decoder LSTM layer, the score
// columns [64 s, 64 s + 64) (s < ceil(S/64)) and ceil(ROWS / (H/16)) of the group's samples for the attention
// heads.  Groups never talk to each other.  Per step t every CTA runs, warp-specialised:
//   tensor-core side (TMA warp + MMA warp, one 6-stage smem ring, all accumulators in tensor memory):
//     P1   scores tile   D3[ROWS x 64]  = h_top(t-1) . Wa_h[64 s.., :]^T
//     R_l  recurrent halves D_l[ROWS x 64] = h_l(t-1) . W_hh_l[slice]^T   (l = 0..L-1; these only need LAST step's
//          state, so they stream in while the attention heads of this step are still being evaluated)
//     IN_0 D_0 += ctx(t) . W_ih_l0[slice, ctx columns]^T      (waits for the group's contexts)
//     IN_l D_l += x_l(t) . W_ih_l[slice]^T, x_l = (dropped) h_{l-1}(t)   (waits for layer l-1 of THIS step)
//   worker side (4 warps):
//     P1 epilogue : adds the hoisted embedding part of the scores (decoder.py:78,84,92: Linear([emb ; h_top]))
//     attention   : 3 softmaxes over ALL slots (the reference's length mask is a no-op) with warp
//                   shuffles, then the three context sums a . M (decoder.py:81,87,95) over memory rows that a loader
//                   thread streams into shared memory with bulk async copies (cp.async.bulk, 24 KB chunks, 4-slot
//                   ring); the memories do not change during the decode, so the loader runs ahead of the scores
//     cell l      : TMEM -> registers, + hoisted embedding product / bias, LSTM cell update (decoder.py:104),
//                   h_l(t) published first (bf16, + its dropped copy), then the activations saved for BPTT
// Hand-over between CTAs of a group goes through global memory (L2) and arrival counters, exactly like the
// persistent encoder kernels (lstm_persist.cu): writers st.global -> CTA barrier -> one fence + red.add, readers
// poll with ld.relaxed.gpu + fence.acq_rel.gpu (+ fence.proxy.async in front of TMA reads).
// All CTAs of a launch must be co-resident: cooperative launch, grid <= #SMs.
#include <cuda_bf16.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "dec_common.cuh"
#include "dropout.cuh"

namespace mmqg {

using namespace tc;
typedef __nv_bfloat16 bf16;

namespace dp {

struct Maps {
  CUtensorMap hs[MAXL];     // state sequences h_l: ((T_q+1)*B, H) bf16, box ROWS x 64
  CUtensorMap hd[MAXL];     // dropped layer outputs (T_q*B, H) bf16 (entries l < L-1, only with dropout)
  CUtensorMap ctx;          // contexts (T_q*B, C) bf16
  CUtensorMap wa;           // attention Linears, state columns: (Sp, H) bf16, box 64 x 64
  CUtensorMap whh[MAXL];    // W_hh_l, 16-unit gate slices: (4H, H)
  CUtensorMap win[MAXL];    // W_ih_l (layer 0: context columns), same row order: (4H, C | H)
};

struct P {
  int T, t0, Tq, B, H, C, Sp, TM, AM, T_t, T_v, H_a, H_v, L;
  int n_slices, n_nt, KBh, KBc;
  int g0;                          // first group of this launch (batches with more groups than fit on the SMs run in waves)
  float* attn_all;                 // (Tq*B, Sp): in = hoisted embedding part + bias, out = softmax weights
  bf16* ctx16;                     // (Tq*B, C)
  float* acts[MAXL];               // (Tq*B, 4H): layer 0 in = hoisted embedding product + bias; out = activated gates
  float* cs[MAXL];                 // ((Tq+1)*B, H)
  bf16* hs[MAXL];                  // ((Tq+1)*B, H)
  bf16* hdrop[MAXL];               // (Tq*B, H) or null
  const float* bias[MAXL];         // (4H) b_ih + b_hh of layers >= 1 (layer 0's is inside the hoisted product)
  const bf16* m_txt16; const bf16* m_vid16; const float* m_aud;
  uint32_t* flags; int fstride;    // per group: F_h[l][0..Tq] | F_s[0..Tq) | F_c[0..Tq)
  float drop_p; unsigned long long seed; const unsigned long long* ctr; int sid0;
  long long* trace;                // debug: 8 %globaltimer stamps per step written by CTA 0 (mmqg_debug_dec_trace), nullable
};

__device__ __forceinline__ uint32_t* flag_h(const P& p, int g, int l, int slab) { return p.flags + (size_t)g * p.fstride + l * (p.Tq + 1) + slab; }
__device__ __forceinline__ uint32_t* flag_s(const P& p, int g, int ta) { return p.flags + (size_t)g * p.fstride + MAXL * (p.Tq + 1) + ta; }
__device__ __forceinline__ uint32_t* flag_c(const P& p, int g, int ta) { return p.flags + (size_t)g * p.fstride + MAXL * (p.Tq + 1) + p.Tq + ta; }

template <int ROWS>
__global__ void __launch_bounds__(224, 1)
dec_seq_fwd_kernel(const __grid_constant__ Maps maps, const P p) {
  constexpr int A_BYTES = ROWS * 128, W_BYTES = 64 * 128, STAGE = A_BYTES + W_BYTES;
  constexpr int NSTAGE = ROWS == 64 ? 6 : 5;
  constexpr int ASTAGES = ROWS == 64 ? 4 : 3;
  constexpr int EW = ROWS / 32;                  // epilogue warps (thread = batch row of the group)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* aring = ring + NSTAGE * STAGE;
  float* stg_all = reinterpret_cast<float*>(aring + ASTAGES * ASLOT);
  float* a_sm = stg_all + 4 * STG_WARP;          // 2 x 512: softmax weights of the two samples in flight
  float* bias_sm = a_sm + 1024;                  // MAXL x 64: bias slices of this CTA's units (layers >= 1)
  __shared__ uint64_t full[NSTAGE], empty[NSTAGE], acc_done[MAXL + 1], acc_free[MAXL + 1], afull[ASTAGES], aempty[ASTAGES];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = p.g0 + blockIdx.x / p.n_slices, s = blockIdx.x % p.n_slices;      // group, unit slice
  const int B = p.B, H = p.H, G = 4 * p.H, L = p.L;
  const int row_g0 = g * ROWS;                                 // first batch row of the group
  const int rows_grp = min(ROWS, B - row_g0);
  const bool has_p1 = s < p.n_nt;
  const bool drop = p.drop_p > 0.f;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i <= MAXL; ++i) { mbar_init(&acc_done[i], 1); mbar_init(&acc_free[i], ROWS); }
    for (int i = 0; i < ASTAGES; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&tmem_slot, 256);
  if (warp < 4) {
    for (int i = threadIdx.x; i < MAXL * 64; i += 128) {
      const int l = i / 64, gg = (i % 64) / 16, u = i % 16;
      bias_sm[i] = (l >= 1 && l < L && p.bias[l]) ? p.bias[l][gg * H + s * 16 + u] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    // ---------------- TMA producer of the tensor-core side ----------------
    if (elect_one()) {
      int i = 0;
      auto job = [&](const CUtensorMap* mA, int rowA, int nkb, const CUtensorMap* mW) {
        for (int kb = 0; kb < nkb; ++kb, ++i) {
          const int st = i % NSTAGE, ph = (i / NSTAGE) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect_tx(&full[st], STAGE);
          uint8_t* a = ring + st * STAGE;
          tma_load_2d(a, mA, &full[st], kb * 64, rowA);
          tma_load_2d(a + A_BYTES, mW, &full[st], kb * 64, s * 64);
        }
      };
      // Job order (the MMA warp walks the same list).  R_l(t) = h_l(t-1) W_hh_l^T only needs layer l of the PREVIOUS
      // step, which is complete long before that step ends, so R_0 and R_1 of the next step are slotted into the
      // current step right behind the input products that wait for the same arrival counter; their 128 KB each stream
      // in while the latency-bound hand-overs of layers 1 and 2 are in flight.  Only R_{L-1} (needs the top layer, i.e.
      // the end of the previous step) stays at the head of the step, behind the score tile.
      //   prologue: R_0 .. R_{L-2} of the first step
      //   step t  : P1(t), R_{L-1}(t), IN_0(t), { IN_l(t), R_{l-1}(t+1) } for l = 1 .. L-1
      auto rec_job = [&](int l, int ta, bool wait) {
        if (wait) { wait_count(flag_h(p, g, l, ta), p.n_slices); fence_proxy_async(); }
        job(&maps.hs[l], ta * B + row_g0, p.KBh, &maps.whh[l]);
      };
      for (int l = 0; l + 1 < L; ++l) rec_job(l, p.t0, false);
      for (int t = 0; t < p.T; ++t) {
        const int ta = p.t0 + t;
        const int rprev = ta * B + row_g0;            // rows of slab ta (state before this step) / of step ta
        if (has_p1) {
          if (t > 0) { wait_count(flag_h(p, g, L - 1, ta), p.n_slices); fence_proxy_async(); }
          job(&maps.hs[L - 1], rprev, p.KBh, &maps.wa);
        }
        rec_job(L - 1, ta, t > 0);
        wait_count(flag_c(p, g, ta), (uint32_t)rows_grp);
        fence_proxy_async();
        job(&maps.ctx, rprev, p.KBc, &maps.win[0]);
        for (int l = 1; l < L; ++l) {
          wait_count(flag_h(p, g, l - 1, ta + 1), p.n_slices);
          fence_proxy_async();
          if (drop) job(&maps.hd[l - 1], rprev, p.KBh, &maps.win[l]);
          else job(&maps.hs[l - 1], (ta + 1) * B + row_g0, p.KBh, &maps.win[l]);
          if (t + 1 < p.T) rec_job(l - 1, ta + 1, false);      // the counter waited for just above covers it
        }
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(ROWS == 64 ? 128 : ROWS, 64, 0, 0);     // 64-row groups run M = 128 with stale upper rows
      int i = 0;
      auto mma_job = [&](uint32_t d_tmem, int nkb, bool fresh) {
        for (int kb = 0; kb < nkb; ++kb, ++i) {
          const int st = i % NSTAGE, ph = (i / NSTAGE) & 1;
          mbar_wait(&full[st], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(ring + st * STAGE), b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                      (!fresh || kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[st]);
        }
      };
      // fresh accumulation into D_l overwrites what the cell epilogue of the previous step read: wait for its release
      auto rec_mma = [&](int l, int t_step) {
        if (t_step > 0) mbar_wait(&acc_free[l], (t_step - 1) & 1);
        tc_fence_after_sync();
        mma_job(tmem_base + l * 64, p.KBh, true);
      };
      for (int l = 0; l + 1 < L; ++l) rec_mma(l, 0);
      for (int t = 0; t < p.T; ++t) {
        if (has_p1) {
          if (t > 0) mbar_wait(&acc_free[MAXL], (t - 1) & 1);
          tc_fence_after_sync();
          mma_job(tmem_base + MAXL * 64, p.KBh, true);
          umma_commit(&acc_done[MAXL]);
        }
        rec_mma(L - 1, t);
        mma_job(tmem_base, p.KBc, false);
        umma_commit(&acc_done[0]);
        for (int l = 1; l < L; ++l) {
          mma_job(tmem_base + l * 64, p.KBh, false);
          umma_commit(&acc_done[l]);
          if (t + 1 < p.T) rec_mma(l - 1, t + 1);
        }
      }
    }
  } else if (warp == 6) {
    // ---------------- loader of the attention memories (they do not change during the decode) ----------------
    if (elect_one()) {
      const Chunker ck(p.H, p.H_a, p.H_v, p.T_t, p.T_v);
      int i = 0;
      for (int t = 0; t < p.T; ++t) {
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
    // ---------------- workers: score epilogue, attention heads, cell updates ----------------
    const int tid = threadIdx.x;                     // 0..127
    const bool epi = warp < EW;                      // owns batch row (row_g0 + tid) of the group in the cell updates
    const int m0w = row_g0 + warp * 32;
    const int m = m0w + lane;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = s * 16;
    float* stg = stg_all + warp * STG_WARP;
    const Chunker ck(p.H, p.H_a, p.H_v, p.T_t, p.T_v);
    int ai = 0;                                      // attention ring position
    const int off_a = p.TM, off_v = p.TM + p.AM;     // slots: [text | audio | video]
    long long* tr = (p.trace && blockIdx.x == 0 && tid == 0) ? p.trace : nullptr;
    for (int t = 0; t < p.T; ++t) {
      const int ta = p.t0 + t;
      if (tr) tr[ta * 12 + 0] = gtime();
      // ---- P1 epilogue: raw scores = tensor-core part + hoisted embedding part (already in attn_all) ----
      if (has_p1 && epi) {
        float* abase = p.attn_all + ((size_t)ta * B + m0w) * p.Sp + s * 64;
        float4 pre[4][4];          // hoisted embedding part + bias of this tile (written before the launch): fetched while the MMAs run
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = s * 64 + q * 16 + 4 * (lane & 3);
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            const int r = 8 * i2 + (lane >> 2);
            pre[q][i2] = (col < p.Sp && r < rows_valid) ? __ldcg(reinterpret_cast<const float4*>(abase + (size_t)r * p.Sp + q * 16 + 4 * (lane & 3)))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        mbar_wait(&acc_done[MAXL], t & 1);
        tc_fence_after_sync();
        float sc[64];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + MAXL * 64 + q * 16, sc + q * 16);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(&acc_free[MAXL]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 tmp[4];
          row_to_coop(stg, lane, sc + q * 16, tmp);
          const int col = s * 64 + q * 16 + 4 * (lane & 3);
          if (col < p.Sp) {
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
              const int r = 8 * i2 + (lane >> 2);
              if (r < rows_valid) {
                const float4 o = pre[q][i2];
                *reinterpret_cast<float4*>(abase + (size_t)r * p.Sp + q * 16 + 4 * (lane & 3)) =
                    make_float4(o.x + tmp[i2].x, o.y + tmp[i2].y, o.z + tmp[i2].z, o.w + tmp[i2].w);
              }
            }
          }
        }
        bar_epi(ROWS);
        if (tid == 0) {
          __threadfence();
          red_relaxed_gpu_add(flag_s(p, g, ta), 1u);
        }
      }
      if (tr) tr[ta * 12 + 1] = gtime();
      // ---- attention heads of this CTA's samples ----
      if (tid == 0) wait_count(flag_s(p, g, ta), (uint32_t)p.n_nt);
      if (tr) tr[ta * 12 + 2] = gtime();
      bar_workers();
      int n_done = 0;
      for (int b0 = row_g0 + s; b0 < row_g0 + rows_grp; b0 += 2 * p.n_slices) {
        // up to two samples per round: one trip to L2 for both score rows, the six softmax heads spread over the 4 warps
        const int nb = (b0 + p.n_slices < row_g0 + rows_grp) ? 2 : 1;
        for (int j = tid; j < nb * p.Sp; j += 128) {
          const int q = j >= p.Sp ? 1 : 0, jj = j - q * p.Sp;
          a_sm[q * 512 + jj] = __ldcg(p.attn_all + ((size_t)ta * B + b0 + q * p.n_slices) * p.Sp + jj);
        }
        bar_workers();
        if (tr && b0 == row_g0 + s) tr[ta * 12 + 8] = gtime();
        for (int task = warp; task < 3 * nb; task += 4) {
          const int q = task / 3, head = task % 3;
          const int off = head == 0 ? 0 : (head == 1 ? off_a : off_v);
          const int len = head == 0 ? p.TM : p.AM;
          float* as = a_sm + q * 512;
          float* arow = p.attn_all + ((size_t)ta * B + b0 + q * p.n_slices) * p.Sp;
          float mx = -INFINITY;
          for (int j = lane; j < len; j += 32) mx = fmaxf(mx, as[off + j]);
          mx = wmax(mx);
          float z = 0.f;
          for (int j = lane; j < len; j += 32) {
            const float e = __expf(as[off + j] - mx);
            as[off + j] = e;
            z += e;
          }
          z = wsum(z);
          const float inv = 1.0f / z;
          for (int j = lane; j < len; j += 32) {
            const float pr = as[off + j] * inv;
            as[off + j] = pr;
            arow[off + j] = pr;                       // softmax weights: what the backward pass (and callers) read
          }
        }
        bar_workers();
        if (tr && b0 == row_g0 + s) tr[ta * 12 + 9] = gtime();
        long long waited = 0;
        for (int q = 0; q < nb; ++q, ++n_done) {
          const int b = b0 + q * p.n_slices;
          if (tr && q == 1 && b0 == row_g0 + s) tr[ta * 12 + 10] = gtime();
          const float* as = a_sm + q * 512;
          float ct[4] = {0.f, 0.f, 0.f, 0.f}, ca[4] = {0.f, 0.f, 0.f, 0.f}, cv[4] = {0.f, 0.f, 0.f, 0.f};
          for (int c = 0; c < ck.count(); ++c, ++ai) {
            const int st = ai % ASTAGES, ph = (ai / ASTAGES) & 1;
            const long long w0 = tr ? gtime() : 0;
            mbar_wait(&afull[st], ph);
            if (tr) waited += gtime() - w0;
            const uint8_t* sl = aring + st * ASLOT;
            if (c < ck.n_t) {
              const int r0 = c * ck.cr_t, nr = min(ck.cr_t, p.T_t - r0);
              if (4 * tid < H) {
                const uint2* col = reinterpret_cast<const uint2*>(sl) + tid;
#pragma unroll 4
                for (int r = 0; r < nr; ++r) {
                  const float w = as[r0 + r];
                  const uint2 u = col[(size_t)r * (H / 4)];
                  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
                  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
                  ct[0] = fmaf(w, lo.x, ct[0]); ct[1] = fmaf(w, lo.y, ct[1]); ct[2] = fmaf(w, hi.x, ct[2]); ct[3] = fmaf(w, hi.y, ct[3]);
                }
              }
            } else if (c < ck.n_t + ck.n_a) {
              const int r0 = (c - ck.n_t) * ck.cr_a, nr = min(ck.cr_a, p.T_v - r0);
              if (4 * tid < p.H_a) {
                const float4* col = reinterpret_cast<const float4*>(sl) + tid;
                for (int r = 0; r < nr; ++r) {
                  const float w = as[off_a + r0 + r];
                  const float4 v = col[(size_t)r * (p.H_a / 4)];
                  ca[0] = fmaf(w, v.x, ca[0]); ca[1] = fmaf(w, v.y, ca[1]); ca[2] = fmaf(w, v.z, ca[2]); ca[3] = fmaf(w, v.w, ca[3]);
                }
              }
            } else {
              const int r0 = (c - ck.n_t - ck.n_a) * ck.cr_v, nr = min(ck.cr_v, p.T_v - r0);
              if (4 * tid < p.H_v) {
                const uint2* col = reinterpret_cast<const uint2*>(sl) + tid;
                for (int r = 0; r < nr; ++r) {
                  const float w = as[off_v + r0 + r];
                  const uint2 u = col[(size_t)r * (p.H_v / 4)];
                  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
                  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
                  cv[0] = fmaf(w, lo.x, cv[0]); cv[1] = fmaf(w, lo.y, cv[1]); cv[2] = fmaf(w, hi.x, cv[2]); cv[3] = fmaf(w, hi.y, cv[3]);
                }
              }
            }
            mbar_arrive(&aempty[st]);
          }
          // contexts of this sample, bf16, decoder.py:99 order [text | audio | video]
          bf16* crow = p.ctx16 + ((size_t)ta * B + b) * p.C;
          auto put4 = [&](bf16* dst, const float* v) {
            __nv_bfloat162 x = __floats2bfloat162_rn(v[0], v[1]), y = __floats2bfloat162_rn(v[2], v[3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&x); u.y = *reinterpret_cast<uint32_t*>(&y);
            *reinterpret_cast<uint2*>(dst) = u;
          };
          if (4 * tid < H) put4(crow + 4 * tid, ct);
          if (4 * tid < p.H_a) put4(crow + H + 4 * tid, ca);
          if (4 * tid < p.H_v) put4(crow + H + p.H_a + 4 * tid, cv);
        }
        bar_workers();                                  // a_sm is rewritten by the next round; covers the context stores
        if (tr && b0 == row_g0 + s) tr[ta * 12 + 11] = waited;
      }
      if (tid == 0 && n_done > 0) {
        __threadfence();                                // after the barrier above: covers every worker's context stores
        red_relaxed_gpu_add(flag_c(p, g, ta), (uint32_t)n_done);
      }
      if (tr) tr[ta * 12 + 3] = gtime();
      // ---- LSTM cells of this CTA's 16 units, layer by layer ----
      if (epi) {
        for (int l = 0; l < L; ++l) {
          float* gbase = p.acts[l] + ((size_t)ta * B + m0w) * G + j0;
          float gxr[64], cp[16];
          __syncwarp();          // c_{t-1} below was stored by other lanes of this warp at the end of the previous step
          if (l == 0) {
            float4 gxc[4][4];
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) coop_ldg(gbase + gg * H, G, rows_valid, lane, gxc[gg]);
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) coop_to_row(stg, lane, gxc[gg], gxr + gg * 16);
          } else {
#pragma unroll
            for (int u = 0; u < 64; ++u) gxr[u] = bias_sm[l * 64 + u];
          }
          {
            float4 c4[4];
            coop_ldg(p.cs[l] + ((size_t)ta * B + m0w) * H + j0, H, rows_valid, lane, c4);
            coop_to_row(stg, lane, c4, cp);
          }
          mbar_wait(&acc_done[l], t & 1);
          if (tr) tr[ta * 12 + 4 + l] = gtime();
          tc_fence_after_sync();
          float acc[64];
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + l * 64 + gg * 16, acc + gg * 16);
          tmem_ld_wait();
          tc_fence_before_sync();
          mbar_arrive(&acc_free[l]);
          float hv[16], cn[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float ig = sigm_fast(acc[u] + gxr[u]);
            const float fg = sigm_fast(acc[16 + u] + gxr[16 + u]);
            const float gt = tanh_fast(acc[32 + u] + gxr[32 + u]);
            const float og = sigm_fast(acc[48 + u] + gxr[48 + u]);
            cn[u] = fmaf(fg, cp[u], ig * gt);
            hv[u] = og * tanh_fast(cn[u]);
            acc[u] = ig; acc[16 + u] = fg; acc[32 + u] = gt; acc[48 + u] = og;
          }
          uint32_t hp[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            __nv_bfloat162 t2 = __floats2bfloat162_rn(hv[2 * v], hv[2 * v + 1]);
            hp[v] = *reinterpret_cast<uint32_t*>(&t2);
          }
          row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, hp, p.hs[l] + ((size_t)(ta + 1) * B + m0w) * H + j0, H, rows_valid);
          if (drop && l + 1 < L) {                     // inter-layer dropout: the next layer reads this copy
            const float ik = 1.0f / (1.0f - p.drop_p);
            const unsigned long long sd = p.seed + (p.ctr ? *p.ctr : 0ull);
            const unsigned long long e0 = ((unsigned long long)ta * B + m) * H + j0;
            uint32_t dpk[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              __nv_bfloat162 t2 = __floats2bfloat162_rn(hv[2 * v] * drop_scale(sd, p.sid0 + l, e0 + 2 * v, p.drop_p, ik),
                                                        hv[2 * v + 1] * drop_scale(sd, p.sid0 + l, e0 + 2 * v + 1, p.drop_p, ik));
              dpk[v] = *reinterpret_cast<uint32_t*>(&t2);
            }
            row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, dpk, p.hdrop[l] + ((size_t)ta * B + m0w) * H + j0, H, rows_valid);
          }
          bar_epi(ROWS);
          if (tid == 0) {
            __threadfence();
            red_relaxed_gpu_add(flag_h(p, g, l, ta + 1), 1u);
          }
          // saved for the backward pass (off the critical path)
          float4 tmp[4];
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            row_to_coop(stg, lane, acc + gg * 16, tmp);
            coop_stg(gbase + gg * H, G, rows_valid, lane, tmp);
          }
          row_to_coop(stg, lane, cn, tmp);
          coop_stg(p.cs[l] + ((size_t)(ta + 1) * B + m0w) * H + j0, H, rows_valid, lane, tmp);
        }
      }
      if (tr) tr[ta * 12 + 7] = gtime();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

// 16-unit gate-slice row order for a (4H, K) weight block read with row pitch ld: out row (s*64 + g*16 + u) =
// w[(g*H + s*16 + u), 0..K) as bf16, padded with zeros to Kp columns
__global__ void pack_rows_gate16_kernel(const float* __restrict__ w, int ld, int K, int Kp, bf16* __restrict__ out, int H) {
  const int r = blockIdx.x;
  const int sl = r / 64, gg = (r % 64) / 16, u = r % 16;
  const float* src = w + (size_t)(gg * H + sl * 16 + u) * ld;
  bf16* dst = out + (size_t)r * Kp;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) dst[k] = __float2bfloat16_rn(k < K ? src[k] : 0.f);
}

}  // namespace dp

// ---- host ---------------------------------------------------------------------------------------
static long long* g_dec_trace = nullptr;      // debug hook, see mmqg_debug_dec_trace()
static int dec_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// Rows per group: 64 (shorter steps: half the attention samples per CTA) when all groups of the batch fit on the
// SMs at once, else 128; batches with more 128-row groups than fit run in waves of co-resident groups.
int dec_persist_rows(int B, int H) {
  static const int want = []() { const char* e = getenv("MMQG_DEC_ROWS"); return e ? atoi(e) : 64; }();
  if (want == 0 || H / 16 > dec_sms()) return 0;
  if (want != 128 && ceil_div(B, 64) * (H / 16) <= dec_sms()) return 64;
  return 128;
}

bool dec_persist_ok(const DecPersistShape& s) {
  if (s.L < 1 || s.L > dp::MAXL) return false;
  if (s.H % 64 != 0 || s.H > 512 || s.H_v % 8 != 0 || s.H_v > 512 || s.H_a % 4 != 0 || s.H_a > 512 || s.C % 8 != 0) return false;
  if (s.Sp > 512 || s.Sp % 8 != 0 || ceil_div(s.Sp, 64) > s.H / 16) return false;
  if (s.H * 2 > dp::ASLOT || s.H_v * 2 > dp::ASLOT || s.H_a * 4 > dp::ASLOT) return false;
  return dec_persist_rows(s.B, s.H) != 0;
}

int pack_rows_gate16(const float* w, int ld, int K, int Kp, void* out, int H, cudaStream_t st) {
  MMQG_REQUIRE(w && out && H % 16 == 0 && Kp >= K && Kp % 8 == 0, "pack_rows_gate16: bad args");
  dp::pack_rows_gate16_kernel<<<4 * H, 128, 0, st>>>(w, ld, K, Kp, reinterpret_cast<bf16*>(out), H);
  MMQG_LAUNCH_CHECK();
  return 0;
}

// launches one call needs: groups beyond the co-resident ones run in further waves
int dec_persist_waves(const DecPersistShape& s) {
  const int rows = dec_persist_rows(s.B, s.H);
  if (!rows) return 0;
  return ceil_div(ceil_div(s.B, rows), dec_sms() / (s.H / 16));
}

size_t dec_persist_flag_words(const DecPersistShape& s, int Tq) {
  const int rows = dec_persist_rows(s.B, s.H);
  if (!rows) return 0;
  return (size_t)ceil_div(s.B, rows) * (dp::MAXL * (Tq + 1) + 2 * Tq);
}

template <int ROWS>
static int launch_dec(const dp::Maps& maps, const dp::P& p, int n_grp, cudaStream_t st) {
  constexpr int STAGE = ROWS * 128 + 8192;
  constexpr int NSTAGE = ROWS == 64 ? 6 : 5, ASTAGES = ROWS == 64 ? 4 : 3;
  const size_t smem = (size_t)NSTAGE * STAGE + (size_t)ASTAGES * dp::ASLOT + 4 * dp::STG_WARP * sizeof(float) + 1024 * sizeof(float) +
                      dp::MAXL * 64 * sizeof(float) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(dp::dec_seq_fwd_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
  MMQG_CUDA(cudaLaunchKernelEx(&cfg, dp::dec_seq_fwd_kernel<ROWS>, maps, p));
  return 0;
}

int dec_seq_fwd_persist(const DecPersistArgs& a, int t0, int T, cudaStream_t st) {
  const DecPersistShape& s = a.shape;
  MMQG_REQUIRE(dec_persist_ok(s), "dec_seq_fwd_persist: shape not supported");
  MMQG_REQUIRE(t0 >= 0 && T >= 1 && t0 + T <= a.Tq, "dec_seq_fwd_persist: steps [%d, %d) outside [0, %d)", t0, t0 + T, a.Tq);
  const int rows = dec_persist_rows(s.B, s.H);
  const int n_grp = ceil_div(s.B, rows);
  dp::P p{};
  p.T = T; p.t0 = t0; p.Tq = a.Tq; p.B = s.B; p.H = s.H; p.C = s.C; p.Sp = s.Sp; p.TM = s.TM; p.AM = s.AM; p.T_t = s.T_t; p.T_v = s.T_v;
  p.H_a = s.H_a; p.H_v = s.H_v; p.L = s.L;
  p.n_slices = s.H / 16; p.n_nt = ceil_div(s.Sp, 64); p.KBh = s.H / 64; p.KBc = ceil_div(s.C, 64);
  p.attn_all = a.attn_all; p.ctx16 = reinterpret_cast<bf16*>(a.ctx16);
  p.m_txt16 = reinterpret_cast<const bf16*>(a.m_txt16); p.m_vid16 = reinterpret_cast<const bf16*>(a.m_vid16); p.m_aud = a.m_aud;
  p.flags = a.flags; p.fstride = dp::MAXL * (a.Tq + 1) + 2 * a.Tq;
  p.drop_p = a.drop_p; p.seed = a.seed; p.ctr = a.ctr; p.sid0 = a.sid0;
  p.trace = g_dec_trace;
  dp::Maps maps;
  const uint64_t rows_state = (uint64_t)(a.Tq + 1) * s.B, rows_step = (uint64_t)a.Tq * s.B;
  for (int l = 0; l < dp::MAXL; ++l) {
    const int ll = l < s.L ? l : 0;            // unused entries alias layer 0 (valid descriptors, never dereferenced)
    p.acts[l] = a.acts[ll]; p.cs[l] = a.cs[ll]; p.hs[l] = reinterpret_cast<bf16*>(a.hs[ll]);
    p.hdrop[l] = reinterpret_cast<bf16*>(a.hdrop[ll]); p.bias[l] = l < s.L ? a.bias[l] : nullptr;
    MMQG_TRY(make_tmap_bf16_2d(&maps.hs[l], a.hs[ll], rows_state, s.H, s.H, rows, 64));
    const void* hd = (a.drop_p > 0.f && ll + 1 < s.L && a.hdrop[ll]) ? a.hdrop[ll] : a.hs[ll];
    MMQG_TRY(make_tmap_bf16_2d(&maps.hd[l], hd, hd == a.hs[ll] ? rows_state : rows_step, s.H, s.H, rows, 64));
    MMQG_TRY(make_tmap_bf16_2d(&maps.whh[l], a.w_hh[ll], 4 * (uint64_t)s.H, s.H, s.H, 64, 64));
    const int Kin = ll == 0 ? s.C : s.H;
    MMQG_TRY(make_tmap_bf16_2d(&maps.win[l], a.w_in[ll], 4 * (uint64_t)s.H, Kin, Kin, 64, 64));
  }
  MMQG_TRY(make_tmap_bf16_2d(&maps.ctx, a.ctx16, rows_step, s.C, s.C, rows, 64));
  MMQG_TRY(make_tmap_bf16_2d(&maps.wa, a.wa_h, s.Sp, s.H, s.H, 64, 64));
  const double fl = 2.0 * T * s.B * ((double)s.Sp * s.H + 4.0 * s.H * ((double)s.C + s.H + (s.L - 1) * 2.0 * s.H) +
                                     (double)s.T_t * s.H + (double)s.T_v * (s.H_a + s.H_v));
  MMQG_PROBE(KC_GEMM_STEP, fl, 0);
  const int per_wave = dec_sms() / p.n_slices;       // co-resident groups
  for (int g0 = 0; g0 < n_grp; g0 += per_wave) {
    p.g0 = g0;
    const int ng = n_grp - g0 < per_wave ? n_grp - g0 : per_wave;
    if (rows == 64) MMQG_TRY(launch_dec<64>(maps, p, ng, st));
    else MMQG_TRY(launch_dec<128>(maps, p, ng, st));
    MMQG_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace mmqg

// Debug hook (not part of the product path): device buffer of 12 * T_q int64 that CTA 0 of the next persistent decoder
// launches fills with %globaltimer stamps per step: 0 step start, 1 score epilogue done, 2 scores of the group seen,
// 3 contexts published, 4..6 accumulator of layer 0..2 complete, 7 step end.  NULL switches it off.
extern "C" void mmqg_debug_dec_trace(void* dev_buf) { mmqg::g_dec_trace = reinterpret_cast<long long*>(dev_buf); }

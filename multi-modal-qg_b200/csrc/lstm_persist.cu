// Persistent recurrent-cell kernels (bf16 mode): ONE launch runs all T timesteps of one LSTM
// layer, forward or backward-through-time, instead of 2 launches per step.
// Replaces the per-token aten::lstm calls of the reference (encoder.py:69,98 driven by
// train.py:164-166) and their autograd twins; the input projection x W_ih^T + b is hoisted
// into one tensor-core GEMM over the whole sequence (gemm_tc.cu) and arrives here as `gx`.
//
// Decomposition.  The hidden units are cut into slices of 16; CTA (s, m) owns slice s for the
// 128 batch rows of m-tile m and keeps ITS weights resident in shared memory for the whole
// sequence, so per step only activations move:
//   forward : D[128 x 64] = h_{t-1}[128 x H] . Wslice^T     Wslice = the 4 gate rows of the 16
//             units (64 x H bf16, 64 KB at H=512), h_{t-1} streamed by TMA (128 KB),
//             accumulator in TMEM; the epilogue warps add gx, apply the cell update with the
//             cell state held in registers across all T steps, and publish h_t.
//   backward: D[128 x 16] = dG_{t+1}[128 x 4H] . W_hh[:, slice]   (W^T slice 16 x 4H, 64 KB,
//             resident; dG_{t+1} streamed through a 6-stage TMA ring), then the pointwise cell
//             gradient with dc in registers, publishing dG_t (bf16, the operand of the
//             hoisted weight-gradient GEMMs).
// Cross-CTA exchange of h_t / dG_t goes through global memory (L2) with one arrival counter
// per (timestep, m-tile): writers st.global -> fence -> bar -> red.release.gpu, the reader's
// TMA thread spins with ld.acquire.gpu and issues fence.proxy.async before the bulk loads.
// All CTAs must be co-resident (grid <= #SMs, 1 CTA/SM): launched cooperatively.
#include <cuda_bf16.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "dropout.cuh"

namespace mmqg {

using namespace tc;
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }


// ---- warp-cooperative tile movers ---------------------------------------------------------
// A warp owns 32 consecutive batch rows and per row only touches short segments (16 fp32 = 64 B
// of a row that is 2-8 KB long).  Thread-per-row global accesses would touch 32 different lines
// per instruction and flood the LSU (measured: ~5.6k L1 wavefronts per step, which also delayed
// the flag polls by ~3 us).  Instead 4 lanes share one row segment (8 rows per instruction) and
// a padded per-warp smem tile transposes between that layout and the thread-per-row layout the
// TMEM accumulator arrives in.
static constexpr int STG_LD = 20;            // floats per staged row (16 + 4 pad, keeps 16 B alignment)
static constexpr int STG_WARP = 32 * STG_LD; // floats per warp

__device__ __forceinline__ void coop_ldg(const float* base, size_t row_stride, int rows_valid, int lane, float4 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    v[i] = r < rows_valid ? *reinterpret_cast<const float4*>(base + (size_t)r * row_stride + 4 * (lane & 3))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void coop_stg(float* base, size_t row_stride, int rows_valid, int lane, const float4 (&v)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    if (r < rows_valid) *reinterpret_cast<float4*>(base + (size_t)r * row_stride + 4 * (lane & 3)) = v[i];
  }
}
// cooperative registers -> staging -> this thread's row (16 floats)
__device__ __forceinline__ void coop_to_row(float* stg, int lane, const float4 (&v)[4], float* mine) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(stg + (8 * i + (lane >> 2)) * STG_LD + 4 * (lane & 3)) = v[i];
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(stg + lane * STG_LD + 4 * q);
    mine[4 * q] = x.x; mine[4 * q + 1] = x.y; mine[4 * q + 2] = x.z; mine[4 * q + 3] = x.w;
  }
}
// this thread's row (16 floats) -> staging -> cooperative registers
__device__ __forceinline__ void row_to_coop(float* stg, int lane, const float* mine, float4 (&v)[4]) {
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<float4*>(stg + lane * STG_LD + 4 * q) = make_float4(mine[4 * q], mine[4 * q + 1], mine[4 * q + 2], mine[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(stg + (8 * i + (lane >> 2)) * STG_LD + 4 * (lane & 3));
}
// 16 bf16 per row (packed as 8 words): 2 lanes share a row segment, 16 rows per instruction
__device__ __forceinline__ void row_bf16_to_global(uint32_t* stg, int lane, const uint32_t (&w8)[8], bf16* base, size_t row_stride,
                                                   int rows_valid) {
  __syncwarp();
  *reinterpret_cast<uint4*>(stg + lane * 12) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
  *reinterpret_cast<uint4*>(stg + lane * 12 + 4) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = 16 * i + (lane >> 1), hsel = lane & 1;
    const uint4 x = *reinterpret_cast<const uint4*>(stg + r * 12 + 4 * hsel);
    if (r < rows_valid) *reinterpret_cast<uint4*>(base + (size_t)r * row_stride + 8 * hsel) = x;
  }
}

// kernel timeline record: buf[0] = atomic slot counter, then (tag, start, end) triples
__device__ __forceinline__ int ktrace_begin(long long* buf, int tag) {
  const int slot = (int)atomicAdd(reinterpret_cast<unsigned long long*>(buf), 1ull);
  buf[1 + 3 * slot] = tag;
  buf[2 + 3 * slot] = gtime();
  return slot;
}
__device__ __forceinline__ void ktrace_end(long long* buf, int slot) { buf[3 + 3 * slot] = gtime(); }

// ---- per-row shifted variants for variable-length batches: row r of the warp belongs to a sample whose
// shift `sh` lives in lane r; element rows are addressed by (time - shift), masked rows are skipped / zero.
__device__ __forceinline__ void coop_ldg_shift(const float* base, size_t row_stride, int rows_valid, int lane, float4 (&v)[4], int sh,
                                               int tg, long long ts) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    const int shr = __shfl_sync(0xffffffffu, sh, r);
    v[i] = (r < rows_valid && tg >= shr)
               ? *reinterpret_cast<const float4*>(base + (size_t)r * row_stride - (long long)shr * ts + 4 * (lane & 3))
               : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void coop_stg_shift(float* base, size_t row_stride, int rows_valid, int lane, const float4 (&v)[4], int sh,
                                               int tg, long long ts) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    const int shr = __shfl_sync(0xffffffffu, sh, r);
    if (r < rows_valid && tg >= shr) *reinterpret_cast<float4*>(base + (size_t)r * row_stride - (long long)shr * ts + 4 * (lane & 3)) = v[i];
  }
}
__device__ __forceinline__ void row_bf16_to_global_shift(uint32_t* stg, int lane, const uint32_t (&w8)[8], bf16* base, size_t row_stride,
                                                         int rows_valid, int sh, int tg, long long ts) {
  __syncwarp();
  *reinterpret_cast<uint4*>(stg + lane * 12) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
  *reinterpret_cast<uint4*>(stg + lane * 12 + 4) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = 16 * i + (lane >> 1), hsel = lane & 1;
    const int shr = __shfl_sync(0xffffffffu, sh, r);
    const uint4 x = *reinterpret_cast<const uint4*>(stg + r * 12 + 4 * hsel);
    if (r < rows_valid && tg >= shr) *reinterpret_cast<uint4*>(base + (size_t)r * row_stride - (long long)shr * ts + 8 * hsel) = x;
  }
}

struct LstmFwdP {
  float* gates;        // (T*B, 4H) fp32: in = x W_ih^T + b (hoisted), out = activated gates i,f,g,o
  float* cs;           // ((T+1)*B, H) fp32; slab t+1 receives c_t (slab 0 is not read: c_{-1} = 0)
  bf16* hs;            // ((T+1)*B, H) bf16; slab 0 = h_{-1} (zeros), slab t+1 receives h_t
  float* mem;          // optional batch-major fp32 copy of h_t: mem[b*mem_ld + t*H + j]
  bf16* mem16;         // optional batch-major bf16 copy (same indexing)
  long long mem_ld;
  uint32_t* flags;     // ((T+1) * n_mt) arrival counters, zeroed before launch
  int T, B, H, n_mt, n_slices, KB;
  int load_c0;         // 1: c_{-1} is read from slab 0 of cs (continuing a sequence chunk), 0: zeros
  long long* trace;    // debug: per-step clock64 stamps of CTA (0,0), 8 per step (nullable)
  long long* ktrace; int ktag;   // debug: kernel-level timeline (see mmqg_debug_ktrace)
  DropSpec dr;         // dr.out: optional dropped bf16 copy of h_t, (T*B, H) rows t*B+b (input of the next layer)
  LenSpec len;
  int dbg_nosave;      // timing experiments only (MMQG_DEBUG_NOSAVE=1): the activated gates / c_t are not saved for the backward pass
};

// CL = 2 (MMQG_FWD_MC=1): the two CTAs of a (2,1,1) cluster -- neighbouring unit slices of one m-tile, which read the SAME
// 128 KB h_{t-1} tile -- each issue half of its eight 16 KB boxes as a cluster-multicast TMA load, so the tile leaves L2 once
// per pair instead of once per CTA.  A box may be written into the peer's ring slot as soon as the issuing CTA has seen the
// step's arrival counter complete: the peer publishes h_{t-1} only after its step-(t-1) MMAs (the last readers of that slot)
// have committed.  Every CTA arms its own eight mbarriers for the full tile; a complete_tx that lands before the expect_tx
// leaves the transaction count negative until it is armed.
template <int CL>
__global__ void __launch_bounds__(160, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, LstmFwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                       // KB x (64 rows x 128 B)
  uint8_t* sA = smem + p.KB * 8192;         // KB x (128 rows x 128 B)
  __shared__ uint64_t w_full, a_full[8], mma_done, tmem_free;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, mt = blockIdx.y;
  const int H = p.H, B = p.B, G = 4 * p.H;

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int k = 0; k < 8; ++k) mbar_init(&a_full[k], 1);
    mbar_init(&mma_done, 1);
    mbar_init(&tmem_free, 128);
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmH);
  }
  if (warp == 4) tmem_alloc(&tmem_slot, 64);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  if (CL > 1) cluster_sync_all();             // the peer's mbarriers exist before a multicast load can signal them

  if (warp == 4) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      long long* tr = p.trace;
      mbar_expect_tx(&w_full, p.KB * 8192);
      for (int kb = 0; kb < p.KB; ++kb) tma_load_2d(sW + kb * 8192, &tmW, &w_full, kb * 64, slice * 64);
      for (int t = 0; t < p.T; ++t) {
        if (t > 0) {
          mbar_wait(&mma_done, (t - 1) & 1);                       // sA is free again
          const uint32_t* f = p.flags + (size_t)t * p.n_mt + mt;    // h_{t-1} complete for this m-tile?
          // relaxed poll + one acquire load (measured 0.4 us per step cheaper than a poll + fence.acq_rel.gpu, which costs a MEMBAR)
          wait_counter_acquire(f, (uint32_t)p.n_slices);
          fence_proxy_async();      // hands the ordering of the counter observation to the async proxy (TMA)
        }
        if (tr) { const long long g = gtime(); atomicMin((unsigned long long*)&tr[t * 8 + 0], (unsigned long long)g); atomicMax((unsigned long long*)&tr[t * 8 + 1], (unsigned long long)g); }
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_expect_tx(&a_full[kb], 16384);
          if (CL == 1) tma_load_2d(sA + kb * 16384, &tmH, &a_full[kb], kb * 64, t * B + mt * 128);
          else if ((uint32_t)(kb & (CL - 1)) == crank)
            tma_load_2d_mc(sA + kb * 16384, &tmH, &a_full[kb], kb * 64, t * B + mt * 128, (uint16_t)((1u << CL) - 1u));
        }
        if (t == 0) mbar_wait(&w_full, 0);
        else mbar_wait(&tmem_free, (t - 1) & 1);                    // epilogue has drained the accumulator
        tc_fence_after_sync();
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&a_full[kb], t & 1);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + kb * 16384), b_addr = smem_u32(sW + kb * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        if (tr) atomicMax((unsigned long long*)&tr[t * 8 + 7], (unsigned long long)gtime());
        umma_commit(&mma_done);
      }
    }
  } else {
    // ---- epilogue warps: thread = batch row, 16 hidden units x 4 gates ----
    const int row = warp * 32 + lane;
    const int m0w = mt * 128 + warp * 32;                 // first batch row of this warp
    const int m = m0w + lane;
    const bool valid = m < B;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = slice * 16;
    float* stg = reinterpret_cast<float*>(sA + p.KB * 16384) + warp * STG_WARP;
    const bool kt = p.ktrace && row == 0 && slice == 0 && mt == 0;
    const int kslot = kt ? ktrace_begin(p.ktrace, p.ktag) : 0;
    const int sh = (p.len.shift && valid) ? p.len.shift[m] : 0;       // variable lengths: first real step of this row
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    if (p.load_c0) {
      float4 c4[4];
      coop_ldg(p.cs + (size_t)m0w * H + j0, H, rows_valid, lane, c4);
      coop_to_row(stg, lane, c4, c);
    }
    for (int t = 0; t < p.T; ++t) {
      // prefetch the hoisted pre-gates of this warp's 32 rows (independent of the recurrence)
      float* gbase = p.gates + ((size_t)t * B + m0w) * G + j0;
      float gxr[64];
      {
        float4 gxc[4][4];
#pragma unroll
        for (int g = 0; g < 4; ++g) coop_ldg(gbase + g * H, G, rows_valid, lane, gxc[g]);
        // transpose to thread-per-row while the tensor core is still busy with this step
#pragma unroll
        for (int g = 0; g < 4; ++g) coop_to_row(stg, lane, gxc[g], gxr + g * 16);
      }
      mbar_wait(&mma_done, t & 1);
      long long* tre = (p.trace && row == 0) ? p.trace : nullptr;
      if (tre) { const long long g = gtime(); atomicMin((unsigned long long*)&tre[t * 8 + 2], (unsigned long long)g); atomicMax((unsigned long long*)&tre[t * 8 + 3], (unsigned long long)g); }
      tc_fence_after_sync();
      float acc[64];
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + g * 16, acc + g * 16);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&tmem_free);
#pragma unroll
      for (int u = 0; u < 64; ++u) acc[u] += gxr[u];
      float hv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float ig = sigm_fast(acc[u]);
        const float fg = sigm_fast(acc[16 + u]);
        const float gg = tanh_fast(acc[32 + u]);
        const float og = sigm_fast(acc[48 + u]);
        c[u] = fmaf(fg, c[u], ig * gg);
        hv[u] = og * tanh_fast(c[u]);
        acc[u] = ig; acc[16 + u] = fg; acc[32 + u] = gg; acc[48 + u] = og;
      }
      const int tg = p.len.t_base + t;
      if (tg < sh) {          // before the sample's first token: the state stays at its zero initial value
#pragma unroll
        for (int u = 0; u < 16; ++u) { c[u] = 0.f; hv[u] = 0.f; }
      }
      // h_t first: it is the only thing the other CTAs are waiting for
      uint32_t hp[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        __nv_bfloat162 t2 = __floats2bfloat162_rn(hv[2 * v], hv[2 * v + 1]);
        hp[v] = *reinterpret_cast<uint32_t*>(&t2);
      }
      row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, hp, p.hs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid);
      if (tre) atomicMax((unsigned long long*)&tre[t * 8 + 6], (unsigned long long)gtime());
      // publish: CTA barrier, then ONE thread fences (cumulativity covers the CTA's h stores)
      // and bumps the arrival counter
      epi_bar_sync();
      if (row == 0) {
        __threadfence();
        red_relaxed_gpu_add(p.flags + (size_t)(t + 1) * p.n_mt + mt, 1u);   // the fence above is the release
        if (tre) { const long long g = gtime(); atomicMin((unsigned long long*)&tre[t * 8 + 4], (unsigned long long)g); atomicMax((unsigned long long*)&tre[t * 8 + 5], (unsigned long long)g); }
      }
      // everything else (saved activations for the backward pass, c_t, the attention-memory
      // copy) is off the critical path and overlaps the wait for the next step
      {
        float4 tmp[4];
        if (!p.dbg_nosave) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            row_to_coop(stg, lane, acc + g * 16, tmp);
            coop_stg(gbase + g * H, G, rows_valid, lane, tmp);
          }
        }
        if (!p.dbg_nosave || t == p.T - 1) {
          row_to_coop(stg, lane, c, tmp);
          coop_stg(p.cs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid, lane, tmp);
        }
        const bool shifted = p.len.shift && p.len.mem_shift;      // memory row = the sample's own position
        if (p.mem) {
          row_to_coop(stg, lane, hv, tmp);
          if (shifted) coop_stg_shift(p.mem + (size_t)m0w * p.mem_ld + (size_t)t * H + j0, (size_t)p.mem_ld, rows_valid, lane, tmp, sh, tg, H);
          else coop_stg(p.mem + (size_t)m0w * p.mem_ld + (size_t)t * H + j0, (size_t)p.mem_ld, rows_valid, lane, tmp);
        }
        if (p.mem16) {
          if (shifted)
            row_bf16_to_global_shift(reinterpret_cast<uint32_t*>(stg), lane, hp, p.mem16 + (size_t)m0w * p.mem_ld + (size_t)t * H + j0,
                                     (size_t)p.mem_ld, rows_valid, sh, tg, H);
          else
            row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, hp, p.mem16 + (size_t)m0w * p.mem_ld + (size_t)t * H + j0,
                               (size_t)p.mem_ld, rows_valid);
        }
        if (p.dr.out) {       // inter-layer dropout: the next layer reads this copy
          const float ik = 1.0f / (1.0f - p.dr.p);
          const unsigned long long sd = p.dr.seed + (p.dr.ctr ? *p.dr.ctr : 0ull);
          const unsigned long long e0 = p.dr.base + ((unsigned long long)t * B + m) * H + j0;
          uint32_t dp[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            __nv_bfloat162 t2 = __floats2bfloat162_rn(hv[2 * v] * drop_scale(sd, p.dr.sid, e0 + 2 * v, p.dr.p, ik),
                                                      hv[2 * v + 1] * drop_scale(sd, p.dr.sid, e0 + 2 * v + 1, p.dr.p, ik));
            dp[v] = *reinterpret_cast<uint32_t*>(&t2);
          }
          row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, dp, reinterpret_cast<bf16*>(p.dr.out) + ((size_t)t * B + m0w) * p.dr.ld + j0,
                             (size_t)p.dr.ld, rows_valid);
        }
      }
      (void)valid;
    }
    if (kt) ktrace_end(p.ktrace, kslot);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 64);
  }
  if (CL > 1) cluster_sync_all();             // nobody leaves while a peer's multicast load may still target its ring
}

// ---------------------------------------------------------------------------------------------
// Two m-tiles per CTA.  The recurrences of different batch rows are independent chains, and one step of a chain
// is a sequence of latencies (counter poll -> TMA -> MMA -> TMEM load -> cell math -> store -> fence -> counter)
// that leaves the SM idle most of the time.  Here CTA (s, pair) serves m-tiles 2*pair and 2*pair+1 with ONE
// resident weight slice: the TMA and MMA threads alternate between the two chains through one 8-stage operand
// ring, each chain has its own accumulator columns in tensor memory, its own 4 epilogue warps and its own
// barriers, so the operand ingest of one chain runs under the hand-over latency of the other.  Same per-step
// time with half the CTAs per layer (32 at H = 512, B = 256): three or four layer/chunk kernels are co-resident
// on the 148 SMs instead of two.  Arrival counters and data layouts are those of lstm_seq_fwd_kernel.
static constexpr int FWD2_STAGES = 8;

__global__ void __launch_bounds__(320, 1)
lstm_seq_fwd2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, LstmFwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                              // KB x (64 rows x 128 B), resident
  uint8_t* sA = smem + p.KB * 8192;                // FWD2_STAGES x (128 rows x 128 B) ring shared by both chains
  float* stg_base = reinterpret_cast<float*>(sA + FWD2_STAGES * 16384);
  __shared__ uint64_t w_full, full[FWD2_STAGES], empty[FWD2_STAGES], mma_done[2], tmem_free[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const int H = p.H, B = p.B, G = 4 * p.H;
  const int mt0 = 2 * blockIdx.y;
  const int n_chain = min(2, p.n_mt - mt0);        // 1 when the number of m-tiles is odd and this is the last CTA row

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int s = 0; s < FWD2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(&mma_done[q], 1); mbar_init(&tmem_free[q], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmH);
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 8) {
    // ---- TMA producer: chains alternate step by step ----
    if (elect_one()) {
      long long* trp = (p.trace && slice == 0 && blockIdx.y == 0) ? p.trace : nullptr;   // debug stamps, see tools/trace_lstm2.py
      mbar_expect_tx(&w_full, p.KB * 8192);
      for (int kb = 0; kb < p.KB; ++kb) tma_load_2d(sW + kb * 8192, &tmW, &w_full, kb * 64, slice * 64);
      int i = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int q = 0; q < n_chain; ++q) {
          const int mt = mt0 + q;
          if (t > 0) {
            const uint32_t* f = p.flags + (size_t)t * p.n_mt + mt;    // h_{t-1} complete for this m-tile?
            if (trp) trp[t * 16 + q * 8 + 7] = gtime();
            wait_counter_acquire(f, (uint32_t)p.n_slices);
            fence_proxy_async();
          }
          if (trp) trp[t * 16 + q * 8 + 0] = gtime();
          for (int kb = 0; kb < p.KB; ++kb, ++i) {
            const int st = i % FWD2_STAGES, ph = (i / FWD2_STAGES) & 1;
            mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], 16384);
            tma_load_2d(sA + st * 16384, &tmH, &full[st], kb * 64, t * B + mt * 128);
          }
          if (trp) trp[t * 16 + q * 8 + 1] = gtime();
        }
      }
    }
  } else if (warp == 9) {
    // ---- MMA issuer: same job order as the producer ----
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      long long* trm = (p.trace && slice == 0 && blockIdx.y == 0) ? p.trace : nullptr;
      mbar_wait(&w_full, 0);
      int i = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int q = 0; q < n_chain; ++q) {
          if (t > 0) mbar_wait(&tmem_free[q], (t - 1) & 1);           // the chain's epilogue has drained its accumulator
          tc_fence_after_sync();
          for (int kb = 0; kb < p.KB; ++kb, ++i) {
            const int st = i % FWD2_STAGES, ph = (i / FWD2_STAGES) & 1;
            mbar_wait(&full[st], ph);
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(sA + st * 16384), b_addr = smem_u32(sW + kb * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + q * 64, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[st]);
          }
          umma_commit(&mma_done[q]);
          if (trm) trm[t * 16 + q * 8 + 2] = gtime();
        }
      }
    }
  } else if ((warp >> 2) < n_chain) {
    // ---- epilogue warps of chain q: thread = batch row, 16 hidden units x 4 gates ----
    const int q = warp >> 2, ws = warp & 3;
    const int mt = mt0 + q;
    const int row = ws * 32 + lane;
    const int m0w = mt * 128 + ws * 32;
    const int m = m0w + lane;
    const bool valid = m < B;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = slice * 16;
    float* stg = stg_base + warp * STG_WARP;
    const uint32_t tacc = tmem_base + (static_cast<uint32_t>(32 * ws) << 16) + q * 64;
    const bool kt = p.ktrace && row == 0 && slice == 0 && mt == 0;
    const int kslot = kt ? ktrace_begin(p.ktrace, p.ktag) : 0;
    const int sh = (p.len.shift && valid) ? p.len.shift[m] : 0;
    long long* tre = (p.trace && row == 0 && slice == 0 && blockIdx.y == 0) ? p.trace : nullptr;
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    if (p.load_c0) {
      float4 c4[4];
      coop_ldg(p.cs + (size_t)m0w * H + j0, H, rows_valid, lane, c4);
      coop_to_row(stg, lane, c4, c);
    }
    for (int t = 0; t < p.T; ++t) {
      float* gbase = p.gates + ((size_t)t * B + m0w) * G + j0;
      float gxr[64];
      {
        float4 gxc[4][4];
#pragma unroll
        for (int g = 0; g < 4; ++g) coop_ldg(gbase + g * H, G, rows_valid, lane, gxc[g]);
#pragma unroll
        for (int g = 0; g < 4; ++g) coop_to_row(stg, lane, gxc[g], gxr + g * 16);
      }
      mbar_wait(&mma_done[q], t & 1);
      if (tre) tre[t * 16 + q * 8 + 3] = gtime();
      tc_fence_after_sync();
      float acc[64];
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_ld_32x16(tacc + g * 16, acc + g * 16);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&tmem_free[q]);
#pragma unroll
      for (int u = 0; u < 64; ++u) acc[u] += gxr[u];
      float hv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float ig = sigm_fast(acc[u]);
        const float fg = sigm_fast(acc[16 + u]);
        const float gg = tanh_fast(acc[32 + u]);
        const float og = sigm_fast(acc[48 + u]);
        c[u] = fmaf(fg, c[u], ig * gg);
        hv[u] = og * tanh_fast(c[u]);
        acc[u] = ig; acc[16 + u] = fg; acc[32 + u] = gg; acc[48 + u] = og;
      }
      const int tg = p.len.t_base + t;
      if (tg < sh) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { c[u] = 0.f; hv[u] = 0.f; }
      }
      uint32_t hp[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        __nv_bfloat162 t2 = __floats2bfloat162_rn(hv[2 * v], hv[2 * v + 1]);
        hp[v] = *reinterpret_cast<uint32_t*>(&t2);
      }
      row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, hp, p.hs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid);
      // publish: barrier over the chain's 4 warps, then ONE thread fences and bumps the m-tile's counter
      if (q == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (row == 0) {
        if (tre) tre[t * 16 + q * 8 + 4] = gtime();
        __threadfence();
        red_relaxed_gpu_add(p.flags + (size_t)(t + 1) * p.n_mt + mt, 1u);
        if (tre) tre[t * 16 + q * 8 + 5] = gtime();
      }
      // saved activations, c_t, attention-memory copy, dropped copy: off the critical path
      {
        float4 tmp[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          row_to_coop(stg, lane, acc + g * 16, tmp);
          coop_stg(gbase + g * H, G, rows_valid, lane, tmp);
        }
        row_to_coop(stg, lane, c, tmp);
        coop_stg(p.cs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid, lane, tmp);
        const bool shifted = p.len.shift && p.len.mem_shift;
        if (p.mem) {
          row_to_coop(stg, lane, hv, tmp);
          if (shifted) coop_stg_shift(p.mem + (size_t)m0w * p.mem_ld + (size_t)t * H + j0, (size_t)p.mem_ld, rows_valid, lane, tmp, sh, tg, H);
          else coop_stg(p.mem + (size_t)m0w * p.mem_ld + (size_t)t * H + j0, (size_t)p.mem_ld, rows_valid, lane, tmp);
        }
        if (p.mem16) {
          if (shifted)
            row_bf16_to_global_shift(reinterpret_cast<uint32_t*>(stg), lane, hp, p.mem16 + (size_t)m0w * p.mem_ld + (size_t)t * H + j0,
                                     (size_t)p.mem_ld, rows_valid, sh, tg, H);
          else
            row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, hp, p.mem16 + (size_t)m0w * p.mem_ld + (size_t)t * H + j0,
                               (size_t)p.mem_ld, rows_valid);
        }
        if (p.dr.out) {
          const float ik = 1.0f / (1.0f - p.dr.p);
          const unsigned long long sd = p.dr.seed + (p.dr.ctr ? *p.dr.ctr : 0ull);
          const unsigned long long e0 = p.dr.base + ((unsigned long long)t * B + m) * H + j0;
          uint32_t dp[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            __nv_bfloat162 t2 = __floats2bfloat162_rn(hv[2 * v] * drop_scale(sd, p.dr.sid, e0 + 2 * v, p.dr.p, ik),
                                                      hv[2 * v + 1] * drop_scale(sd, p.dr.sid, e0 + 2 * v + 1, p.dr.p, ik));
            dp[v] = *reinterpret_cast<uint32_t*>(&t2);
          }
          row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, dp, reinterpret_cast<bf16*>(p.dr.out) + ((size_t)t * B + m0w) * p.dr.ld + j0,
                             (size_t)p.dr.ld, rows_valid);
        }
      }
      if (tre) tre[t * 16 + q * 8 + 6] = gtime();
    }
    if (kt) ktrace_end(p.ktrace, kslot);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}

// ---------------------------------------------------------------------------------------------
struct LstmBwdP {
  const float* acts;   // (T*B, 4H) activated gates from the forward
  const float* cs;     // ((T+1)*B, H) cell states (slab 0 must hold c_{-1}: zeros for the encoders)
  bf16* dg;            // (T*B, 4H) bf16 out: d loss / d pre-activations (also streamed back through tmG)
  const float* dh_ext; // external d loss / d h_t per step: dh_ext[t*ext_ts + b*ext_ld + j]  (nullable)
  long long ext_ts, ext_ld;
  const float* dh_last; // (B,H) fp32 added to d h_{T-1} (nullable)
  const float* dc_last; // (B,H) fp32 d loss / d c_{T-1} from the consumer of the final state (nullable)
  uint32_t* flags;      // (T * n_mt), zeroed before launch
  int T, B, H, n_mt, n_slices, NKB;   // NKB = 4H/64
  int has_next;         // 1: dg slab T (first step of the following chunk) exists and feeds step T-1
  float* dc_out;        // (B,H) receives d loss / d c_{-1} at the end (nullable)
  long long* ktrace; int ktag;
  DropSpec dr;          // dr.p > 0: dh_ext is the gradient w.r.t. the DROPPED h_t -> multiplied by the mask here
  LenSpec len;
  long long* trace;     // debug: per-step %globaltimer stamps of CTA (0,0) of lstm_seq_bwd4_kernel, 16 per step (tools/trace_lstm_bwd.py)
};

static constexpr int BWD_STAGES = 6;     // 144 KB in flight: the streaming rate is ring bytes / TMA round trip

__global__ void __launch_bounds__(192, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, LstmBwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                        // NKB x (16 rows x 128 B)
  uint8_t* sA = smem + p.NKB * 2048;         // BWD_STAGES x (128 rows x 128 B)
  __shared__ uint64_t w_full, full[BWD_STAGES], empty[BWD_STAGES], mma_done, tmem_free;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, mt = blockIdx.y;
  const int H = p.H, B = p.B, G = 4 * p.H, T = p.T;

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&mma_done, 1);
    mbar_init(&tmem_free, 128);
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 5) tmem_alloc(&tmem_slot, 32);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(&w_full, p.NKB * 2048);
      for (int kb = 0; kb < p.NKB; ++kb) tma_load_2d(sW + kb * 2048, &tmW, &w_full, kb * 64, slice * 16);
      int i = 0;
      for (int t = T - 1 - (p.has_next ? 0 : 1); t >= 0; --t) {   // step t consumes dG_{t+1}
        if (t + 1 < T) {                           // slab T comes from an earlier launch: already complete
          const uint32_t* f = p.flags + (size_t)(t + 1) * p.n_mt + mt;
          wait_counter_acquire(f, (uint32_t)p.n_slices);
          fence_proxy_async();
        }
        for (int kb = 0; kb < p.NKB; ++kb, ++i) {
          const int s = i % BWD_STAGES, ph = (i / BWD_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], 16384);
          tma_load_2d(sA + s * 16384, &tmG, &full[s], kb * 64, (t + 1) * B + mt * 128);
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 16, 0, 0);
      mbar_wait(&w_full, 0);
      int i = 0, it = 0;
      for (int t = T - 1 - (p.has_next ? 0 : 1); t >= 0; --t, ++it) {
        if (it > 0) mbar_wait(&tmem_free, (it - 1) & 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < p.NKB; ++kb, ++i) {
          const int s = i % BWD_STAGES, ph = (i / BWD_STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * 16384), b_addr = smem_u32(sW + kb * 2048);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&mma_done);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const int m0w = mt * 128 + warp * 32;
    const int m = m0w + lane;
    const bool valid = m < B;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = slice * 16;
    float* stg = reinterpret_cast<float*>(sA + BWD_STAGES * 16384) + warp * STG_WARP;
    const bool kt = p.ktrace && row == 0 && slice == 0 && mt == 0;
    const int kslot = kt ? ktrace_begin(p.ktrace, p.ktag) : 0;
    const int sh = (p.len.shift && valid) ? p.len.shift[m] : 0;
    float dc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) dc[u] = (p.dc_last && valid) ? p.dc_last[(size_t)m * H + j0 + u] : 0.f;
    int it = 0;
    for (int t = T - 1; t >= 0; --t) {
      // cooperative prefetch (4 lanes per 64-byte row segment) of everything that does not
      // depend on the recurrence
      float4 a4[4][4], cn4[4], cp4[4], ex4[4];
      const float* abase = p.acts + ((size_t)t * B + m0w) * G + j0;
#pragma unroll
      for (int g = 0; g < 4; ++g) coop_ldg(abase + g * H, G, rows_valid, lane, a4[g]);
      coop_ldg(p.cs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid, lane, cn4);
      coop_ldg(p.cs + ((size_t)t * B + m0w) * H + j0, H, rows_valid, lane, cp4);
      const int tg = p.len.t_base + t;
      if (p.dh_ext && p.len.shift && p.len.mem_shift)
        coop_ldg_shift(p.dh_ext + (size_t)t * p.ext_ts + (size_t)m0w * p.ext_ld + j0, (size_t)p.ext_ld, rows_valid, lane, ex4, sh, tg, p.ext_ts);
      else if (p.dh_ext) coop_ldg(p.dh_ext + (size_t)t * p.ext_ts + (size_t)m0w * p.ext_ld + j0, (size_t)p.ext_ld, rows_valid, lane, ex4);
      else {
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) ex4[i2] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float a[64], cn[16], cp[16], ex[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) coop_to_row(stg, lane, a4[g], a + g * 16);
      coop_to_row(stg, lane, cn4, cn);
      coop_to_row(stg, lane, cp4, cp);
      coop_to_row(stg, lane, ex4, ex);
      if (p.dr.p > 0.f) {     // ext is d/d(dropped h_t): back through this layer's output mask
        const float ik = 1.0f / (1.0f - p.dr.p);
        const unsigned long long sd = p.dr.seed + (p.dr.ctr ? *p.dr.ctr : 0ull);
        const unsigned long long e0 = p.dr.base + ((unsigned long long)t * B + m) * H + j0;
#pragma unroll
        for (int u = 0; u < 16; ++u) ex[u] *= drop_scale(sd, p.dr.sid, e0 + u, p.dr.p, ik);
      }
      float dh[16];
      if (t == T - 1 && !p.has_next) {
#pragma unroll
        for (int u = 0; u < 16; ++u) dh[u] = (p.dh_last && valid) ? p.dh_last[(size_t)m * H + j0 + u] : 0.f;
      } else {
        mbar_wait(&mma_done, it & 1);
        tc_fence_after_sync();
        tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * warp) << 16), dh);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(&tmem_free);
        ++it;
      }
      {
        float dgv[64];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float ig = a[u], fg = a[16 + u], gg = a[32 + u], og = a[48 + u];
          const float d = dh[u] + ex[u];
          const float tc_ = tanh_fast(cn[u]);
          const float dct = dc[u] + d * og * (1.f - tc_ * tc_);
          dgv[u] = dct * gg * ig * (1.f - ig);
          dgv[16 + u] = dct * cp[u] * fg * (1.f - fg);
          dgv[32 + u] = dct * ig * (1.f - gg * gg);
          dgv[48 + u] = d * tc_ * og * (1.f - og);
          dc[u] = dct * fg;
        }
        if (tg < sh) {        // masked step: no gradient reaches the weights or anything earlier
#pragma unroll
          for (int u = 0; u < 64; ++u) dgv[u] = 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) dc[u] = 0.f;
        }
        bf16* dbase = p.dg + ((size_t)t * B + m0w) * G + j0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w8[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            __nv_bfloat162 t2 = __floats2bfloat162_rn(dgv[g * 16 + 2 * v], dgv[g * 16 + 2 * v + 1]);
            w8[v] = *reinterpret_cast<uint32_t*>(&t2);
          }
          row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, w8, dbase + g * H, G, rows_valid);
        }
      }
      epi_bar_sync();
      if (row == 0) {
        __threadfence();
        red_relaxed_gpu_add(p.flags + (size_t)t * p.n_mt + mt, 1u);
      }
    }
    if (p.dc_out) {
      float4 tmp[4];
      row_to_coop(stg, lane, dc, tmp);
      coop_stg(p.dc_out + (size_t)m0w * H + j0, H, rows_valid, lane, tmp);
    }
    if (kt) ktrace_end(p.ktrace, kslot);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 32);
  }
}

// ---------------------------------------------------------------------------------------------
// BPTT kernel, split-K over a 4-CTA cluster (round 2).
//
// lstm_seq_bwd_kernel above gives every CTA the output slice D[128 x 16] = dG_{t+1}[128 x 4H] . W_hh[:, 16 units]:
// each CTA streams the WHOLE 512 KB dG tile of its m-tile every step, and at ~100 GB/s of L2->SM ingest
// per SM that is 5 of the step's 9.7 us (profiles/r01_ncu_full_prof_r1d: tensor pipe 5.7 %).
// Here the 4 CTAs of a cluster own 64 hidden units together and split the K = 4H reduction by GATE:
// CTA rank r keeps W_hh[gate r rows, 64 units]^T (64 x H bf16, 64 KB at H=512) resident, streams only the
// gate-r quarter of dG_{t+1} (128 x H bf16 = 128 KB per step) and forms the partial product
// P_r[128 x 64] in TMEM.  The partials are exchanged through distributed shared memory: every thread
// (= batch row) stores the 16-column blocks that belong to the three peer CTAs straight into their
// exchange buffers with st.async, which reports the bytes to the receiving warp's mbarrier (complete_tx):
// no fence and no separate arrive; after its own barrier completes (3 peers x 32 rows x 64 B) the thread sums
// the four partials of its 16 units and continues with the pointwise cell gradient exactly like the
// kernel above.  Exchange buffers are double-buffered by step parity: a peer can only write buffer b
// again after it has received this CTA's partial of the step in between, which this CTA sends after it
// has consumed buffer b -- no "buffer free" message is needed.
// Cross-cluster hand-over of dG_t is unchanged (global memory + one arrival counter per (t, m-tile)).
struct LstmBwd4Smem {
  static constexpr int STAGES = 6;
  static constexpr int XROW = 64;                   // bytes per exchanged row (16 fp32)
  static constexpr int XSRC = 128 * XROW;           // one source CTA's block: 128 rows
  static constexpr int XBUF = 3 * XSRC;             // three peers
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// asynchronous remote store that reports its 16 bytes to an mbarrier in the destination CTA (complete_tx)
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t cluster_bar, float a, float b, float c, float d) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%2, %3, %4, %5}, [%1];"
               ::"r"(cluster_addr), "r"(cluster_bar), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

__global__ void __launch_bounds__(192, 1)
lstm_seq_bwd4_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, LstmBwdP p) {
  constexpr int STAGES = LstmBwd4Smem::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int KB = p.H / 64;                    // 64-wide k-blocks of one gate
  uint8_t* sW = smem;                         // KB x (64 rows x 128 B): W_hh[gate r, 64 units]^T, resident
  uint8_t* sA = sW + KB * 8192;               // STAGES x (128 rows x 128 B): streamed quarter of dG_{t+1}
  uint8_t* sX = sA + STAGES * 16384;          // 2 x 3 x (128 rows x 64 B): partials received from the peers
  float* stg_base = reinterpret_cast<float*>(sX + 2 * LstmBwd4Smem::XBUF);
  __shared__ uint64_t w_full, full[STAGES], empty[STAGES], mma_done, tmem_free, xfull[2][4];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, mt = blockIdx.y;
  const int rank = (int)cluster_ctarank();    // == slice & 3: gate (K quarter) of this CTA
  const int unit0 = (slice >> 2) * 64;        // first hidden unit of the cluster
  const int H = p.H, B = p.B, G = 4 * p.H, T = p.T;

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&mma_done, 1);
    mbar_init(&tmem_free, 128);
    for (int b2 = 0; b2 < 2; ++b2)            // one per (buffer, epilogue warp): completes on 3 peers x 32 rows x 64 bytes
      for (int w2 = 0; w2 < 4; ++w2) mbar_init(&xfull[b2][w2], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 5) tmem_alloc(&tmem_slot, 64);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  cluster_sync_all();                         // peers' mbarriers exist before anyone arrives on them
  const uint32_t tmem_base = tmem_slot;
  const int t_first = T - 1 - (p.has_next ? 0 : 1);     // first step that has a matrix product

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(&w_full, KB * 8192);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * 8192, &tmW, &w_full, rank * H + kb * 64, unit0);
      int i = 0;
      long long* trp = (p.trace && slice == 0 && mt == 0) ? p.trace : nullptr;
      for (int t = t_first; t >= 0; --t) {          // step t consumes dG_{t+1}
        if (trp) trp[t * 16 + 0] = gtime();
        if (t + 1 < T) {                            // slab T comes from an earlier launch: already complete
          const uint32_t* f = p.flags + (size_t)(t + 1) * p.n_mt + mt;
          wait_counter_acquire(f, (uint32_t)p.n_slices);
          fence_proxy_async();
        }
        if (trp) trp[t * 16 + 1] = gtime();
        if (p.trace && mt == 0) {      // spread over the CTAs of m-tile 0
          const unsigned long long g = (unsigned long long)gtime();
          atomicMax((unsigned long long*)&p.trace[t * 16 + 12], g);
          atomicMin((unsigned long long*)&p.trace[t * 16 + 13], g);
        }
        for (int kb = 0; kb < KB; ++kb, ++i) {
          const int s = i % STAGES, ph = (i / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], 16384);
          tma_load_2d(sA + s * 16384, &tmG, &full[s], rank * H + kb * 64, (t + 1) * B + mt * 128);
        }
        if (trp) trp[t * 16 + 2] = gtime();
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      mbar_wait(&w_full, 0);
      int i = 0, it = 0;
      for (int t = t_first; t >= 0; --t, ++it) {
        if (it > 0) mbar_wait(&tmem_free, (it - 1) & 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < KB; ++kb, ++i) {
          const int s = i % STAGES, ph = (i / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * 16384), b_addr = smem_u32(sW + kb * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&mma_done);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const int m0w = mt * 128 + warp * 32;
    const int m = m0w + lane;
    const bool valid = m < B;
    const int rows_valid = max(0, min(32, B - m0w));
    const int j0 = slice * 16;                 // == unit0 + 16 * rank
    float* stg = stg_base + warp * STG_WARP;
    const bool kt = p.ktrace && row == 0 && slice == 0 && mt == 0;
    const int kslot = kt ? ktrace_begin(p.ktrace, p.ktag) : 0;
    const int sh = (p.len.shift && valid) ? p.len.shift[m] : 0;
    // exchange addressing: row `row` of source s lives at sX + buf*XBUF + slot(s)*XSRC + row*64, its four 16-byte
    // chunks XOR-swizzled with (row >> 1) & 3 so that neither the remote stores nor the local loads conflict
    const uint32_t xrow_off = (uint32_t)row * LstmBwd4Smem::XROW;
    const uint32_t swz = (uint32_t)(row >> 1) & 3u;
    uint32_t peer_x[4], peer_bar[2][4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int my_slot_at_peer = rank < rr ? rank : rank - 1;      // slot index of source `rank` inside peer rr
      peer_x[rr] = mapa_u32(smem_u32(sX) + my_slot_at_peer * LstmBwd4Smem::XSRC + xrow_off, (uint32_t)rr);
      peer_bar[0][rr] = mapa_u32(smem_u32(&xfull[0][warp]), (uint32_t)rr);
      peer_bar[1][rr] = mapa_u32(smem_u32(&xfull[1][warp]), (uint32_t)rr);
    }
    float dc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) dc[u] = (p.dc_last && valid) ? p.dc_last[(size_t)m * H + j0 + u] : 0.f;
    long long* tre = (p.trace && row == 0 && slice == 0 && mt == 0) ? p.trace : nullptr;
    int it = 0;
    for (int t = T - 1; t >= 0; --t) {
      float4 a4[4][4], cn4[4], cp4[4], ex4[4];
      const float* abase = p.acts + ((size_t)t * B + m0w) * G + j0;
#pragma unroll
      for (int g = 0; g < 4; ++g) coop_ldg(abase + g * H, G, rows_valid, lane, a4[g]);
      coop_ldg(p.cs + ((size_t)(t + 1) * B + m0w) * H + j0, H, rows_valid, lane, cn4);
      coop_ldg(p.cs + ((size_t)t * B + m0w) * H + j0, H, rows_valid, lane, cp4);
      const int tg = p.len.t_base + t;
      if (p.dh_ext && p.len.shift && p.len.mem_shift)
        coop_ldg_shift(p.dh_ext + (size_t)t * p.ext_ts + (size_t)m0w * p.ext_ld + j0, (size_t)p.ext_ld, rows_valid, lane, ex4, sh, tg, p.ext_ts);
      else if (p.dh_ext) coop_ldg(p.dh_ext + (size_t)t * p.ext_ts + (size_t)m0w * p.ext_ld + j0, (size_t)p.ext_ld, rows_valid, lane, ex4);
      else {
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) ex4[i2] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float a[64], cn[16], cp[16], ex[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) coop_to_row(stg, lane, a4[g], a + g * 16);
      coop_to_row(stg, lane, cn4, cn);
      coop_to_row(stg, lane, cp4, cp);
      coop_to_row(stg, lane, ex4, ex);
      if (p.dr.p > 0.f) {     // ext is d/d(dropped h_t): back through this layer's output mask
        const float ik = 1.0f / (1.0f - p.dr.p);
        const unsigned long long sd = p.dr.seed + (p.dr.ctr ? *p.dr.ctr : 0ull);
        const unsigned long long e0 = p.dr.base + ((unsigned long long)t * B + m) * H + j0;
#pragma unroll
        for (int u = 0; u < 16; ++u) ex[u] *= drop_scale(sd, p.dr.sid, e0 + u, p.dr.p, ik);
      }
      float dh[16];
      if (t == T - 1 && !p.has_next) {
#pragma unroll
        for (int u = 0; u < 16; ++u) dh[u] = (p.dh_last && valid) ? p.dh_last[(size_t)m * H + j0 + u] : 0.f;
      } else {
        const int xb = it & 1;
        if (tre) tre[t * 16 + 3] = gtime();
        mbar_wait(&mma_done, it & 1);
        if (tre) tre[t * 16 + 4] = gtime();
        tc_fence_after_sync();
        float part[64];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + q * 16, part + q * 16);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(&tmem_free);
        // this warp's receive barrier for the step: 3 peers x 32 rows x 64 bytes will be reported to it
        if (lane == 0) mbar_expect_tx(&xfull[xb][warp], 3 * 32 * LstmBwd4Smem::XROW);
        // send: columns [16 rr, 16 rr + 16) of this CTA's partial belong to peer rr (same warp index there);
        // every 16-byte store reports itself to that warp's barrier -- no fence, no separate arrive
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          if (rr == rank) continue;
          const uint32_t dst = peer_x[rr] + (uint32_t)xb * LstmBwd4Smem::XBUF;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_async_v4(dst + (((uint32_t)c ^ swz) << 4), peer_bar[xb][rr], part[16 * rr + 4 * c], part[16 * rr + 4 * c + 1],
                        part[16 * rr + 4 * c + 2], part[16 * rr + 4 * c + 3]);
        }
        // receive the three peers' partials of this warp's 32 rows
        if (tre) tre[t * 16 + 5] = gtime();
        mbar_wait_cluster(&xfull[xb][warp], (uint32_t)(it >> 1) & 1u);
        if (tre) tre[t * 16 + 6] = gtime();
#pragma unroll
        for (int u = 0; u < 16; ++u) dh[u] = 0.f;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)        // own partial without a trip through shared memory
          if (rr == rank) {
#pragma unroll
            for (int u = 0; u < 16; ++u) dh[u] = part[16 * rr + u];
          }
        const uint8_t* xin = sX + xb * LstmBwd4Smem::XBUF + xrow_off;
#pragma unroll
        for (int sl = 0; sl < 3; ++sl) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(xin + sl * LstmBwd4Smem::XSRC + (((uint32_t)c ^ swz) << 4));
            dh[4 * c] += v.x; dh[4 * c + 1] += v.y; dh[4 * c + 2] += v.z; dh[4 * c + 3] += v.w;
          }
        }
        ++it;
      }
      {
        float dgv[64];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float ig = a[u], fg = a[16 + u], gg = a[32 + u], og = a[48 + u];
          const float d = dh[u] + ex[u];
          const float tc_ = tanh_fast(cn[u]);
          const float dct = dc[u] + d * og * (1.f - tc_ * tc_);
          dgv[u] = dct * gg * ig * (1.f - ig);
          dgv[16 + u] = dct * cp[u] * fg * (1.f - fg);
          dgv[32 + u] = dct * ig * (1.f - gg * gg);
          dgv[48 + u] = d * tc_ * og * (1.f - og);
          dc[u] = dct * fg;
        }
        if (tg < sh) {        // masked step: no gradient reaches the weights or anything earlier
#pragma unroll
          for (int u = 0; u < 64; ++u) dgv[u] = 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) dc[u] = 0.f;
        }
        bf16* dbase = p.dg + ((size_t)t * B + m0w) * G + j0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w8[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            __nv_bfloat162 t2 = __floats2bfloat162_rn(dgv[g * 16 + 2 * v], dgv[g * 16 + 2 * v + 1]);
            w8[v] = *reinterpret_cast<uint32_t*>(&t2);
          }
          row_bf16_to_global(reinterpret_cast<uint32_t*>(stg), lane, w8, dbase + g * H, G, rows_valid);
        }
      }
      if (tre) tre[t * 16 + 7] = gtime();
      epi_bar_sync();
      if (row == 0) {
        if (tre) tre[t * 16 + 8] = gtime();
        __threadfence();
        red_relaxed_gpu_add(p.flags + (size_t)t * p.n_mt + mt, 1u);
        if (tre) tre[t * 16 + 9] = gtime();
        if (p.trace && mt == 0) {
          const unsigned long long g = (unsigned long long)gtime();
          atomicMax((unsigned long long*)&p.trace[t * 16 + 10], g);
          atomicMin((unsigned long long*)&p.trace[t * 16 + 11], g);
        }
      }
    }
    if (p.dc_out) {
      float4 tmp[4];
      row_to_coop(stg, lane, dc, tmp);
      coop_stg(p.dc_out + (size_t)m0w * H + j0, H, rows_valid, lane, tmp);
    }
    if (kt) ktrace_end(p.ktrace, kslot);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                         // no CTA leaves while a peer may still store into its exchange buffers
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 64);
  }
}

// gate-slice packing of W_hh for the forward kernel: row (s*64 + g*16 + u) = W_hh[g*H + s*16 + u, :]
__global__ void pack_whh_fwd_kernel(const float* __restrict__ w, bf16* __restrict__ out, int H) {
  const int r = blockIdx.x;                 // packed row
  const int s = r / 64, g = (r % 64) / 16, u = r % 16;
  const float* src = w + (size_t)(g * H + s * 16 + u) * H;
  bf16* dst = out + (size_t)r * H;
  for (int k = threadIdx.x; k < H; k += blockDim.x) dst[k] = __float2bfloat16_rn(src[k]);
}
// transpose for the backward kernel: out[j, r] = W_hh[r, j]   (H x 4H)
__global__ void pack_whh_bwd_kernel(const float* __restrict__ w, bf16* __restrict__ out, int H) {
  __shared__ float tile[32][33];
  const int G = 4 * H;
  const int r0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = w[(size_t)(r0 + i) * H + j0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    out[(size_t)(j0 + i) * G + r0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x][i]);
}

// y = sum of up to two split-K partial stacks (used to form dh_last for the text encoder)
__global__ void sum_partials_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, long long stride,
                                    float* __restrict__ y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < na; ++k) s += a[(size_t)k * stride + i];
  for (int k = 0; k < nb; ++k) s += b[(size_t)k * stride + i];
  y[i] = s;
}

// ---- host -----------------------------------------------------------------------------------
static long long* g_lstm_trace = nullptr;   // debug hook, see mmqg_debug_lstm_trace()
static long long* g_ktrace = nullptr;       // debug hook, see mmqg_debug_ktrace()
thread_local int tl_ktag = 0;               // tag of the next persistent launch (set by the engine)

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

bool lstm_persist_ok(int B, int H) {
  if (H % 64 != 0 || H > 512 || H < 64) return false;
  const int n_mt = ceil_div(B, 128);
  return (H / 16) * n_mt <= num_sms();
}

int device_sms() { return num_sms(); }
// Which forward kernel: one m-tile per CTA runs a step in 5.3 us, two per CTA in 7.4 us (the operand ingest of the two
// chains, 2 x 128 KB per step, serialises on the SM's ~107 GB/s L2 port) -- so the two-chain kernel is taken when
// the one-chain grid does not leave room for a second co-resident launch (B > 256 at H = 512: the layer wavefront
// would collapse, or the grid would not fit at all).  MMQG_FWD2=2 forces it wherever it applies.
static bool fwd2_launch(int n_slices, int n_mt) {
  static const int mode = []() { const char* e = getenv("MMQG_FWD2"); return e ? atoi(e) : 1; }();
  if (mode == 0 || n_mt < 2) return false;
  return mode == 2 || 2 * n_slices * n_mt > num_sms();
}
// the forward kernel alone: with two m-tiles per CTA it covers twice the batch rows of the BPTT kernel
bool lstm_persist_fwd_ok(int B, int H) {
  if (H % 64 != 0 || H > 512 || H < 64) return false;
  const int n_mt = ceil_div(B, 128);
  return (H / 16) * (fwd2_launch(H / 16, n_mt) ? ceil_div(n_mt, 2) : n_mt) <= num_sms();
}

int lstm_persist_fwd_ctas(int B, int H) {
  const int n_mt = ceil_div(B, 128);
  return (H / 16) * (fwd2_launch(H / 16, n_mt) ? ceil_div(n_mt, 2) : n_mt);
}

int pack_whh(const float* w_hh, void* fwd_packed, void* bwd_packed, int H, cudaStream_t st) {
  MMQG_REQUIRE(w_hh && H % 32 == 0, "pack_whh: bad args");
  if (fwd_packed) {
    pack_whh_fwd_kernel<<<4 * H, 128, 0, st>>>(w_hh, reinterpret_cast<bf16*>(fwd_packed), H);
    MMQG_LAUNCH_CHECK();
  }
  if (bwd_packed) {
    pack_whh_bwd_kernel<<<dim3(4 * H / 32, H / 32), dim3(32, 8), 0, st>>>(w_hh, reinterpret_cast<bf16*>(bwd_packed), H);
    MMQG_LAUNCH_CHECK();
  }
  return 0;
}

int sum_partials(const float* a, int na, const float* b, int nb, long long stride, float* y, int n, cudaStream_t st) {
  sum_partials_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a, na, b, nb, stride, y, n);
  MMQG_LAUNCH_CHECK();
  return 0;
}

template <typename Kern, typename P>
static int launch_coop(Kern kern, dim3 grid, int threads, size_t smem, const CUtensorMap& m0, const CUtensorMap& m1, const P& p,
                       cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MMQG_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, p));
  return 0;
}

// Launch with a (CLX,1,1) cluster AND the cooperative attribute: the clusters of one launch still wait on
// each other through the global arrival counters, so the whole grid must be co-resident.
template <typename Kern, typename P>
static cudaError_t launch_coop_cluster(Kern kern, int clx, dim3 grid, int threads, size_t smem, const CUtensorMap& m0, const CUtensorMap& m1,
                                       const P& p, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = clx;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  // Profiling aid: Nsight Compute refuses a launch that is both cooperative and clustered (LaunchFailed).  Under ncu the
  // kernels of a process are serialised anyway, so MMQG_NCU=1 drops the cooperative attribute (attr[0]) for its runs only.
  static const bool ncu = []() { const char* e = getenv("MMQG_NCU"); return e && e[0] == '1'; }();
  cfg.attrs = ncu ? attr + 1 : attr;
  cfg.numAttrs = ncu ? 1 : 2;
  return cudaLaunchKernelEx(&cfg, kern, m0, m1, p);
}

// gates: (T*B,4H) fp32 pre-gates -> activated gates.  wp_fwd: pack_whh forward layout.  hs slab 0 must be zeros.
int lstm_seq_fwd_persist(float* gates, float* cs, void* hs, const void* wp_fwd, float* mem, void* mem16, long long mem_ld,
                         uint32_t* flags, int T, int B, int H, int load_c0, cudaStream_t st, DropSpec dr, bool zero_flags, LenSpec len) {
  MMQG_REQUIRE(lstm_persist_fwd_ok(B, H), "lstm_seq_fwd_persist: shape B=%d H=%d not supported", B, H);
  LstmFwdP p{gates, cs, reinterpret_cast<bf16*>(hs), mem, reinterpret_cast<bf16*>(mem16), mem_ld, flags, T, B, H, ceil_div(B, 128), H / 16, H / 64, load_c0, g_lstm_trace, g_ktrace, tl_ktag, dr, len};
  { const char* e = getenv("MMQG_DEBUG_NOSAVE"); p.dbg_nosave = e && e[0] == '1'; }
  CUtensorMap tmW, tmH;
  MMQG_TRY(make_tmap_bf16_2d(&tmW, wp_fwd, 4 * (uint64_t)H, H, H, 64, 64));
  MMQG_TRY(make_tmap_bf16_2d(&tmH, hs, (uint64_t)(T + 1) * B, H, H, 128, 64));
  const size_t smem = (size_t)p.KB * (8192 + 16384) + 4 * STG_WARP * sizeof(float) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(lstm_seq_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (8192 + 16384) + 4 * STG_WARP * sizeof(float) + 1024));
    MMQG_CUDA(cudaFuncSetAttribute(lstm_seq_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (8192 + 16384) + 4 * STG_WARP * sizeof(float) + 1024));
    attr = true;
  }
  if (zero_flags) MMQG_CUDA(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)(T + 1) * p.n_mt, st));
  const double fl = 2.0 * T * B * 4.0 * H * H;
  if (fwd2_launch(p.n_slices, p.n_mt)) {      // two m-tiles per CTA: half the CTAs per launch
    const size_t smem2 = (size_t)p.KB * 8192 + FWD2_STAGES * 16384 + 8 * STG_WARP * sizeof(float) + 1024;
    static bool attr2 = false;
    if (!attr2) {
      MMQG_CUDA(cudaFuncSetAttribute(lstm_seq_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     8 * 8192 + FWD2_STAGES * 16384 + 8 * STG_WARP * (int)sizeof(float) + 1024));
      attr2 = true;
    }
    MMQG_PROBE(KC_LSTM_PERSIST, fl, 0);
    MMQG_TRY(launch_coop(lstm_seq_fwd2_kernel, dim3(p.n_slices, ceil_div(p.n_mt, 2)), 320, smem2, tmW, tmH, p, st));
    MMQG_LAUNCH_CHECK();
    return 0;
  }
  MMQG_PROBE(KC_LSTM_PERSIST, fl, 0);
  // MMQG_FWD_MC=1: pairs of unit slices as (2,1,1) clusters sharing the h_{t-1} tile by multicast TMA (read per call: A/B runs)
  const char* mc = getenv("MMQG_FWD_MC");
  if (mc && mc[0] == '1' && p.n_slices % 2 == 0 && p.KB % 2 == 0 && !g_lstm_trace) {
    MMQG_CUDA(launch_coop_cluster(lstm_seq_fwd_kernel<2>, 2, dim3(p.n_slices, p.n_mt), 160, smem, tmW, tmH, p, st));
    MMQG_LAUNCH_CHECK();
    return 0;
  }
  MMQG_TRY(launch_coop(lstm_seq_fwd_kernel<1>, dim3(p.n_slices, p.n_mt), 160, smem, tmW, tmH, p, st));
  MMQG_LAUNCH_CHECK();
  return 0;
}

static size_t bwd4_smem_bytes(int H) {
  return (size_t)(H / 64) * 8192 + LstmBwd4Smem::STAGES * 16384 + 2 * LstmBwd4Smem::XBUF + 4 * STG_WARP * sizeof(float) + 1024;
}

// 1 = split-K cluster kernel usable on this device for hidden size H, 0 = not (fall back to lstm_seq_bwd_kernel).
// Decided once per H by an occupancy query with the exact launch attributes; MMQG_BWD4=0 forces the old kernel.
static int g_bwd4_state[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};      // index H/64
static int bwd4_state(int H) {
  int* state = g_bwd4_state;
  const int idx = H / 64;
  if (H % 64 != 0 || idx < 1 || idx > 8) return 0;
  if (state[idx] >= 0) return state[idx];
  const char* e = getenv("MMQG_BWD4");
  if (e && e[0] == '0') return state[idx] = 0;
  const size_t smem = bwd4_smem_bytes(H);
  if (cudaFuncSetAttribute(lstm_seq_bwd4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd4_smem_bytes(512)) != cudaSuccess) {
    cudaGetLastError();
    return state[idx] = 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(H / 16, 1);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&n_clusters, lstm_seq_bwd4_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return state[idx] = 0;
  }
  return state[idx] = n_clusters >= 1 ? 1 : 0;
}

int lstm_seq_bwd_persist(const float* acts, const float* cs, void* dg, const void* wp_bwd, const float* dh_ext,
                         long long ext_ts, long long ext_ld, const float* dh_last, const float* dc_last, uint32_t* flags,
                         int T, int B, int H, int has_next, float* dc_out, cudaStream_t st, DropSpec dr, bool zero_flags, LenSpec len) {
  MMQG_REQUIRE(lstm_persist_ok(B, H), "lstm_seq_bwd_persist: shape B=%d H=%d not supported", B, H);
  LstmBwdP p{acts, cs, reinterpret_cast<bf16*>(dg), dh_ext, ext_ts, ext_ld, dh_last, dc_last, flags,
             T, B, H, ceil_div(B, 128), H / 16, 4 * H / 64, has_next, dc_out, g_ktrace, 1000 + tl_ktag, dr, len, g_lstm_trace};
  if (bwd4_state(H) == 1) {
    CUtensorMap tmW4, tmG4;
    MMQG_TRY(make_tmap_bf16_2d(&tmW4, wp_bwd, H, 4 * (uint64_t)H, 4 * (uint64_t)H, 64, 64));
    MMQG_TRY(make_tmap_bf16_2d(&tmG4, dg, (uint64_t)(T + (has_next ? 1 : 0)) * B, 4 * (uint64_t)H, 4 * (uint64_t)H, 128, 64));
    if (zero_flags) MMQG_CUDA(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)T * p.n_mt, st));
    const double fl4 = 2.0 * (T - 1) * B * 4.0 * H * H;
    MMQG_PROBE(KC_LSTM_PERSIST, fl4, 0);
    const cudaError_t le = launch_coop_cluster(lstm_seq_bwd4_kernel, 4, dim3(p.n_slices, p.n_mt), 192, bwd4_smem_bytes(H), tmW4, tmG4, p, st);
    if (le == cudaSuccess) {
      MMQG_LAUNCH_CHECK();
      return 0;
    }
    // the driver refused the cooperative cluster launch: say so once and use the single-CTA-slice kernel from now on
    fprintf(stderr, "libmmqg: cooperative (4,1,1)-cluster launch of lstm_seq_bwd4_kernel failed (%s); falling back to lstm_seq_bwd_kernel\n",
            cudaGetErrorString(le));
    cudaGetLastError();
    probe_close(st);
    g_bwd4_state[H / 64] = 0;
  }
  CUtensorMap tmW, tmG;
  MMQG_TRY(make_tmap_bf16_2d(&tmW, wp_bwd, H, 4 * (uint64_t)H, 4 * (uint64_t)H, 16, 64));
  MMQG_TRY(make_tmap_bf16_2d(&tmG, dg, (uint64_t)(T + (has_next ? 1 : 0)) * B, 4 * (uint64_t)H, 4 * (uint64_t)H, 128, 64));
  const size_t smem = (size_t)p.NKB * 2048 + BWD_STAGES * 16384 + 4 * STG_WARP * sizeof(float) + 1024;
  static bool attr = false;
  if (!attr) {
    MMQG_CUDA(cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 2048 + BWD_STAGES * 16384 + 4 * STG_WARP * (int)sizeof(float) + 1024));
    attr = true;
  }
  if (zero_flags) MMQG_CUDA(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)T * p.n_mt, st));
  const double fl = 2.0 * (T - 1) * B * 4.0 * H * H;
  MMQG_PROBE(KC_LSTM_PERSIST, fl, 0);
  MMQG_TRY(launch_coop(lstm_seq_bwd_kernel, dim3(p.n_slices, p.n_mt), 192, smem, tmW, tmG, p, st));
  MMQG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmqg

// Debug hook (not part of the product path): device buffer of 8*T int64 that the next forward
// persistent launches fill with clock64() stamps of CTA (0,0); pass NULL to switch off.
extern "C" void mmqg_debug_lstm_trace(void* dev_buf) { mmqg::g_lstm_trace = reinterpret_cast<long long*>(dev_buf); }

// Debug hook: device buffer of int64 (slot counter, then (tag, start ns, end ns) per persistent
// launch; tag = layer*16 + chunk, +1000 for the backward kernels, 900 = video).  NULL switches off.
extern "C" void mmqg_debug_ktrace(void* dev_buf) { mmqg::g_ktrace = reinterpret_cast<long long*>(dev_buf); }


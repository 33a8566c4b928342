// bf16 mode of the whole-path orchestration (see engine.cu for the fp32 twin and the
// reference line map).  Same control flow and time-major layouts; every contraction runs on
// tcgen05 tensor cores through gemm_bf16 with bf16 operands and fp32 accumulation:
//   * weights are re-packed to bf16 once per step (derived caches of the fp32 parameters;
//     the decoder's W_ih_l0 and the three attention Linears are split into their embedding
//     and state/context column blocks so every TMA base is 16-byte aligned),
//   * h sequences, gathered embeddings, contexts, gate gradients and dlogits are bf16,
//   * pre-activations / activated gates, cell state, attention memories, softmax, losses and
//     every gradient tensor handed back to the caller are fp32.
// dX = dG W reuses the packed (4H,in) weight as an MN-major B operand and dW = dG^T X takes
// both operands MN-major, so no transposed copies are ever made.
#include "kernels.h"

namespace mmqg {

static const int kSplitB = 4;
static const int kSplitW = 8;    // split-K of the text encoder's hoisted dW products (K = T_t*B rows)
typedef uint16_t b16;   // storage type for bf16 buffers on the host side of this file

struct Carver16 {
  char* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Ws16 {
  unsigned long long* seed_ctr;       // FIRST 8 bytes of the workspace: number of forward calls (dropout call counter), zeroed by the caller
  int64_t *idx_ctx, *idx_dec, *tgt_tm, *idx_cur;
  float *bsum_text[MMQG_MAX_LAYERS], *bsum_dec[MMQG_MAX_LAYERS], *bsum_vid, *attn_b_cat, *attn_dw_cat, *attn_db_cat;
  b16 *wt_ih[MMQG_MAX_LAYERS], *wt_hh[MMQG_MAX_LAYERS], *wv_ih, *wv_hh;
  // decoder LSTM weights of layer l as ONE bf16 matrix [W_hh | W_in] (4H x (H + I_l)), I_0 = context
  // columns of W_ih_l0 (the embedding columns are hoisted: wd_e), I_l = H: the forward step reads
  // the two column blocks as separate K-major operands, the backward step multiplies dG by the
  // whole matrix in one launch (dh_rec and dx side by side).
  b16 *wd_e, *wd_cat[MMQG_MAX_LAYERS], *wa_e, *wa_h, *wo;
  float* dw_part;                    // split-K partials of a hoisted text weight-gradient product, (kSplitW, 4H, H)
  float* gpart;                      // split-K partials of the decoder's forward step product, (2, B, 4H)
  float* dcat[MMQG_MAX_LAYERS];      // split-K partials of dG_l [W_hh | W_in]: (kSplitB, B, H + I_l)
  b16 *x0, *frames16, *hs_text[MMQG_MAX_LAYERS], *hs_v, *e_dec, *ds16, *ctx16, *hs_dec[MMQG_MAX_LAYERS], *dlogits16;
  b16 *dg_text[MMQG_MAX_LAYERS], *dg_v, *dg_dec[MMQG_MAX_LAYERS];
  float *acts_text[MMQG_MAX_LAYERS], *cs_text[MMQG_MAX_LAYERS], *m_txt, *m_aud, *m_vid, *acts_v, *cs_v;
  float *attn_all, *ds_all, *ctx_tmp, *acts_dec[MMQG_MAX_LAYERS], *cs_dec[MMQG_MAX_LAYERS], *logits, *nll, *dhtop;
  float *dh_rec[MMQG_MAX_LAYERS], *dc[MMQG_MAX_LAYERS], *dx_above, *dq_h, *dctx_all, *dm_txt, *dm_vid, *de_dec;
  float *dh_rec_enc, *dh_rec_vid, *dx_text, *dc_v;
  // persistent recurrent kernels: gate-slice packed W_hh (forward) and W_hh^T (backward),
  // arrival counters, and the summed d loss / d h_final handed to the text encoder
  b16 *wtp_f[MMQG_MAX_LAYERS], *wtp_b[MMQG_MAX_LAYERS], *wvp_f, *wvp_b;
  b16 *xdrop_text[MMQG_MAX_LAYERS], *hdrop_dec[MMQG_MAX_LAYERS];   // dropped layer outputs (inputs of the next layer)
  b16 *m_txt16, *m_vid16;     // bf16 attention memories (what the attention kernels read in this mode)
  uint32_t *flags, *flags_v, *flags_t[MMQG_MAX_LAYERS];
  int *shift_t, *shift_v;             // variable lengths: first real time step of every sample (text / video)
  float* row_w;                       // ... and the 0/1 weight of every (target step, sample) loss row
  float *dh_last, *dh_last_l[MMQG_MAX_LAYERS], *dx_emb;
  // persistent decoder-step kernel (dec_persist.cu): decoder weights in 16-unit gate-slice row order, arrival counters
  b16 *wdp_hh[MMQG_MAX_LAYERS], *wdp_in[MMQG_MAX_LAYERS];
  uint32_t* flags_dec;
  // persistent decoder BPTT kernel (dec_persist_bwd.cu): transposed decoder weights (W_hh_l^T, W_ih_l^T / context columns of
  // W_ih_l0^T: rows = output column, K = gate index), transposed state columns of the attention Linears, arrival counters
  b16 *wdpb[MMQG_MAX_LAYERS], *wahT;
  uint32_t* flags_decb;
  int Spp;
  // fused loss head (vocab_nll.cu): per-tile softmax partials, per-row lse / gradient scale / target logit, split-K scratch of dH
  float *stat_a, *stat_b, *lse, *rscale, *tgt_logit, *dh_part;
  int Sp, Ep, Vp, Rc, Rs;     // Rc: rows of one bf16 d-logits chunk; Rs: rows of the fp32 logits scratch of the sampling decode
  size_t bytes;
};

static int sample_chunk_rows16(int R, int V) {
  // fp32 logits of one chunk of rows for the 'sampling' decode strategy (the only consumer of materialised logits): 8 Mi elements
  const long long budget = 8ll << 20;
  long long rc = budget / (V > 0 ? V : 1);
  rc = rc / 128 * 128;
  if (rc < 128) rc = 128;
  if (rc > R) rc = R;
  return (int)rc;
}

static Ws16 carve16(const mmqg_dims& d, int T_q, void* base) {
  Ws16 w{};
  Carver16 c{reinterpret_cast<char*>(base), 0};
  w.seed_ctr = c.take<unsigned long long>(1);
  const size_t B = d.B, H = d.H, G = 4 * (size_t)d.H, Hv = d.H_v, Gv = 4 * (size_t)d.H_v;
  const int S = d.TM + 2 * d.AM;
  w.Sp = (S + 7) / 8 * 8;
  w.Ep = (d.E + 7) / 8 * 8;
  w.Vp = (d.V + 7) / 8 * 8;
  const size_t Sp = w.Sp, Ep = w.Ep, Q = d.E + d.H, C = (size_t)d.H + d.H_a + d.H_v;
  const size_t R = (size_t)T_q * B, Rt = (size_t)d.T_t * B, Rv = (size_t)d.T_v * B;
  w.Rc = vocab_chunk_rows((int)R, w.Vp);
  w.Rs = sample_chunk_rows16((int)B, d.V);
  w.idx_ctx = c.take<int64_t>(Rt); w.idx_dec = c.take<int64_t>(R); w.tgt_tm = c.take<int64_t>(R);
  w.idx_cur = c.take<int64_t>(B);
  for (int l = 0; l < d.L; ++l) { w.bsum_text[l] = c.take<float>(G); w.bsum_dec[l] = c.take<float>(G); }
  w.bsum_vid = c.take<float>(Gv);
  w.attn_b_cat = c.take<float>(Sp); w.attn_dw_cat = c.take<float>(Sp * Q); w.attn_db_cat = c.take<float>(Sp);
  for (int l = 0; l < d.L; ++l) {
    w.wt_ih[l] = c.take<b16>(G * (l == 0 ? Ep : H)); w.wt_hh[l] = c.take<b16>(G * H);
    w.wd_cat[l] = c.take<b16>(G * (H + (l == 0 ? C : H)));
    w.dcat[l] = c.take<float>(kSplitB * B * (H + (l == 0 ? C : H)));
  }
  w.wv_ih = c.take<b16>(Gv * d.F_v); w.wv_hh = c.take<b16>(Gv * Hv);
  w.wd_e = c.take<b16>(G * Ep);
  w.wa_e = c.take<b16>(Sp * Ep); w.wa_h = c.take<b16>(Sp * H);
  w.wo = c.take<b16>((size_t)d.V * H);
  w.x0 = c.take<b16>(Rt * Ep);
  w.frames16 = c.take<b16>(B * d.T_v * d.F_v);
  for (int l = 0; l < d.L; ++l) {
    w.acts_text[l] = c.take<float>(Rt * G);
    w.hs_text[l] = c.take<b16>((Rt + B) * H);
    w.cs_text[l] = c.take<float>((Rt + B) * H);
    w.dg_text[l] = c.take<b16>(Rt * G);
  }
  w.m_txt = c.take<float>(B * d.TM * H); w.m_aud = c.take<float>(B * d.AM * d.H_a); w.m_vid = c.take<float>(B * d.AM * Hv);
  w.acts_v = c.take<float>(Rv * Gv); w.hs_v = c.take<b16>((Rv + B) * Hv); w.cs_v = c.take<float>((Rv + B) * Hv);
  w.dg_v = c.take<b16>(Rv * Gv);
  w.e_dec = c.take<b16>(R * Ep);
  w.attn_all = c.take<float>(R * Sp); w.ds_all = c.take<float>(R * Sp); w.ds16 = c.take<b16>(R * Sp);
  w.ctx_tmp = c.take<float>(B * C); w.ctx16 = c.take<b16>(R * C);
  for (int l = 0; l < d.L; ++l) {
    w.acts_dec[l] = c.take<float>(R * G);
    w.hs_dec[l] = c.take<b16>((R + B) * H);
    w.cs_dec[l] = c.take<float>((R + B) * H);
    w.dg_dec[l] = c.take<b16>(R * G);
  }
  w.logits = c.take<float>((size_t)w.Rs * d.V);
  w.dlogits16 = c.take<b16>((size_t)w.Rc * w.Vp);
  {
    const size_t rows = R > B ? R : B, ns = vocab_stat_floats((int)rows, d.V);
    w.stat_a = c.take<float>(ns); w.stat_b = c.take<float>(ns);
    w.lse = c.take<float>(rows); w.rscale = c.take<float>(rows); w.tgt_logit = c.take<float>(rows);
    w.dh_part = c.take<float>((size_t)16 * w.Rc * H);
  }
  w.nll = c.take<float>(R); w.dhtop = c.take<float>(R * H);
  for (int l = 0; l < d.L; ++l) { w.dh_rec[l] = c.take<float>(kSplitB * B * H); w.dc[l] = c.take<float>(B * H); }
  w.dx_above = c.take<float>(kSplitB * B * H); w.dq_h = c.take<float>(kSplitB * B * H);
  w.dctx_all = c.take<float>(R * C);
  w.gpart = c.take<float>(2 * B * G);
  w.dw_part = c.take<float>((size_t)kSplitW * G * (H > Ep ? H : Ep));
  w.dm_txt = c.take<float>(B * d.TM * H); w.dm_vid = c.take<float>(B * d.AM * Hv);
  w.de_dec = c.take<float>(R * d.E);
  w.dh_rec_enc = c.take<float>(kSplitB * B * H); w.dh_rec_vid = c.take<float>(kSplitB * B * Hv);
  w.dx_text = c.take<float>(Rt * (d.E > d.H ? d.E : d.H));
  w.dc_v = c.take<float>(B * Hv);
  for (int l = 0; l < d.L; ++l) { w.wtp_f[l] = c.take<b16>(G * H); w.wtp_b[l] = c.take<b16>(G * H); }
  w.wvp_f = c.take<b16>(Gv * Hv); w.wvp_b = c.take<b16>(Gv * Hv);
  w.flags = c.take<uint32_t>((size_t)((d.T_t > d.T_v ? d.T_t : d.T_v) + 1) * ((B + 127) / 128));
  w.flags_v = c.take<uint32_t>((size_t)(d.T_v + 1) * ((B + 127) / 128));
  w.dh_last = c.take<float>(B * H);
  for (int l = 0; l < d.L; ++l) {
    w.flags_t[l] = c.take<uint32_t>((size_t)(d.T_t + 1) * ((B + 127) / 128));
    w.dh_last_l[l] = c.take<float>(B * H);
  }
  w.dx_emb = c.take<float>(Rt * d.E);
  for (int l = 0; l + 1 < d.L; ++l) { w.xdrop_text[l] = c.take<b16>(Rt * H); w.hdrop_dec[l] = c.take<b16>(R * H); }
  w.shift_t = c.take<int>(B); w.shift_v = c.take<int>(B); w.row_w = c.take<float>(R);
  for (int l = 0; l < d.L; ++l) { w.wdp_hh[l] = c.take<b16>(G * H); w.wdp_in[l] = c.take<b16>(G * (l == 0 ? C : H)); }
  w.flags_dec = c.take<uint32_t>((size_t)((B + 63) / 64) * (3 * (T_q + 1) + 2 * T_q) + 64);
  w.Spp = (w.Sp + 63) / 64 * 64;
  for (int l = 0; l < d.L; ++l) w.wdpb[l] = c.take<b16>(dec_bwd_weight_rows(d.H, l) * G);
  w.wahT = c.take<b16>(H * (size_t)w.Spp);
  w.flags_decb = c.take<uint32_t>((size_t)((B + 63) / 64) * 5 * T_q + 64);
  w.m_txt16 = c.take<b16>(B * d.TM * H);
  w.m_vid16 = c.take<b16>(B * d.AM * Hv);
  w.bytes = align_up(c.off, 256);
  return w;
}

size_t train_workspace_bytes_bf16(const mmqg_dims& d, int T_q) { return carve16(d, T_q, nullptr).bytes; }

int check_dims_bf16(const mmqg_dims& d) {
  MMQG_REQUIRE(d.H % 8 == 0 && d.H_v % 8 == 0 && d.H_a % 8 == 0 && d.F_v % 8 == 0,
               "bf16 mode needs H, H_v, H_a, F_v to be multiples of 8 (TMA 16-byte rows); got %d %d %d %d", d.H, d.H_v,
               d.H_a, d.F_v);
  return 0;
}

// gemm_bf16 call helper ---------------------------------------------------------------------
struct Tc {
  mmqg_gemm_bf16_args a;
  // A: (M,K) K-major unless amn (then stored (K,M)); B: (N,K) K-major unless bmn (then stored (K,N)).
  Tc(const void* A, int lda, bool amn, const void* Bm, int ldb, bool bmn, int M, int N, int K, float* C, int ldc) {
    a = mmqg_gemm_bf16_args{};
    a.A = A; a.lda = lda; a.a_mn_major = amn; a.B = Bm; a.ldb = ldb; a.b_mn_major = bmn;
    a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.alpha = 1.f; a.beta = 0.f; a.split_k = 1;
  }
  Tc& second(const void* A2, int lda2, const void* B2, int ldb2, int K2) {
    a.A2 = A2; a.lda2 = lda2; a.B2 = B2; a.ldb2 = ldb2; a.K2 = K2; return *this;
  }
  Tc& accumulate(bool on) { if (on) { a.Cin = reinterpret_cast<const float*>(a.C); a.ldcin = a.ldc; a.beta = 1.f; } return *this; }
  Tc& bias(const float* b) { a.bias = b; return *this; }
  Tc& split(int s, long long stride) { a.split_k = s; a.c_split_stride = stride; return *this; }
  int run(cudaStream_t st) { return gemm_bf16(a, st); }
};

static AttnShape attn_shape16(const mmqg_dims& d, const Ws16& w) {
  AttnShape a{d.B, d.TM, d.AM, d.H, d.H_a, d.H_v, d.T_t, d.T_v};
  a.m_txt16 = w.m_txt16;
  a.m_vid16 = w.m_vid16;
  return a;
}

// Persisting-L2 window over the bf16 attention memories (m_txt16 and m_vid16 are adjacent in the workspace):
// the set-aside is sized to the bytes the decoder really touches (T_t and T_v rows per sample).
static size_t attn_l2_window(const mmqg_dims& d, const Ws16& w) {
  static const bool on = []() { const char* e = getenv("MMQG_L2WIN"); return e && e[0] == '1'; }();
  if (!on || !w.m_txt16 || !w.m_vid16) return 0;
  const size_t span = (size_t)(reinterpret_cast<const char*>(w.m_vid16) - reinterpret_cast<const char*>(w.m_txt16)) +
                      sizeof(b16) * (size_t)d.B * d.AM * d.H_v;
  const size_t touched = sizeof(b16) * (size_t)d.B * ((size_t)d.T_t * d.H + (size_t)d.T_v * d.H_v);
  return l2_window_reserve(touched + (touched >> 2), span);
}

// Debug hook (tools/sections.py; not used by the product path): CUDA events at the section
// boundaries of one eager step on the caller's stream -- 0 start, 1 weights packed, 2 encoders done,
// 3 decoder hoisted products, 4 decoder step loop, 5 loss head, 6 backward start, 7 decoder BPTT
// loop, 8 text BPTT (main-stream part), 9 end.
struct SectionMarks {
  bool on = false;
  cudaEvent_t ev[10] = {};
  bool created = false;
};
static SectionMarks g_sec;
static void mark(int i, cudaStream_t st) {
  if (!g_sec.on) return;
  if (!g_sec.created) {
    for (auto& e : g_sec.ev) cudaEventCreate(&e);
    g_sec.created = true;
  }
  cudaEventRecord(g_sec.ev[i], st);
}

// inter-layer dropout of the current call (set at the entry points; 0 = off)
static thread_local float g_drop_p = 0.f;
static thread_local unsigned long long g_drop_seed = 0;
static thread_local const unsigned long long* g_drop_ctr = nullptr;
static const int kSidText = 10, kSidDec = 20;
// variable-length batch of the current call (mmqg_batch.ctx_len / tgt_len / n_frames): off = all null
struct LenState { bool on = false; const int* shift_t = nullptr; const int* shift_v = nullptr; const float* row_w = nullptr; };
static thread_local LenState g_len;
static LenSpec len_text(int t_base, bool mem) { LenSpec l; if (g_len.on) { l.shift = g_len.shift_t; l.t_base = t_base; l.mem_shift = mem ? 1 : 0; } return l; }
static LenSpec len_video() { LenSpec l; if (g_len.on) { l.shift = g_len.shift_v; l.mem_shift = 1; } return l; }
static int set_len_state(const mmqg_dims& d, const mmqg_batch& bt, const Ws16& w) {
  g_len = LenState();
  if (!(bt.ctx_len || bt.tgt_len || bt.n_frames)) return 0;
  g_len.on = true; g_len.shift_t = w.shift_t; g_len.shift_v = w.shift_v; g_len.row_w = w.row_w;
  (void)d;
  return 0;
}

// The encoders run on the persistent recurrent kernels when the shape allows it (lstm_persist.cu);
// MMQG_PERSIST=0 forces the one-GEMM-plus-pointwise-launch-per-step path (for A/B comparison).
static bool persist_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MMQG_PERSIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
// 1 = persistent recurrent kernels (global arrival counters), 0 = per-step launches
static thread_local bool g_fwd_only = false;      // greedy decode: only the forward kernels (which cover twice the batch rows) are needed
static int persist_kind(int B, int H) {
  if (!persist_enabled()) return 0;
  return (g_fwd_only ? lstm_persist_fwd_ok(B, H) : lstm_persist_ok(B, H)) ? 1 : 0;
}
static bool persist_text(const mmqg_dims& d) { return persist_kind(d.B, d.H) != 0; }
static bool persist_video(const mmqg_dims& d) { return persist_kind(d.B, d.H_v) != 0; }
// packed W_hh of a persistent layer: the forward layout is needed at once, the backward (transposed)
// layout only by the BPTT -- either may be null
static int pack_rec(const float* w_hh, void* fwd, void* bwd, int B, int H, cudaStream_t st) {
  (void)B;
  return pack_whh(w_hh, fwd, bwd, H, st);
}
static int rec_fwd(float* gates, float* cs, void* hs, const void* wp, float* mem, void* mem16, long long mem_ld, uint32_t* flags, int T,
                   int B, int H, cudaStream_t st, LenSpec len = LenSpec()) {
  return lstm_seq_fwd_persist(gates, cs, hs, wp, mem, mem16, mem_ld, flags, T, B, H, 0, st, DropSpec(), true, len);
}
static int rec_bwd(const float* acts, const float* cs, void* dg, const void* wp, const float* ext, long long ts, long long ld,
                   const float* dh_last, const float* dc_last, uint32_t* flags, int T, int B, int H, cudaStream_t st,
                   LenSpec len = LenSpec()) {
  return lstm_seq_bwd_persist(acts, cs, dg, wp, ext, ts, ld, dh_last, dc_last, flags, T, B, H, 0, nullptr, st, DropSpec(), true, len);
}

// The decoder's teacher-forced steps run inside the persistent decoder-step kernel (dec_persist.cu) when the
// shape allows it; MMQG_DEC_PERSIST=0 keeps the launch-per-phase loop (for A/B comparison).
static DecPersistShape dec_shape(const mmqg_dims& d, const Ws16& w) {
  return DecPersistShape{d.B, d.H, d.H + d.H_a + d.H_v, w.Sp, d.TM, d.AM, d.T_t, d.T_v, d.H_a, d.H_v, d.L};
}
static bool dec_persist_enabled(const mmqg_dims& d, const Ws16& w) {
  const char* e = getenv("MMQG_DEC_PERSIST");
  if (e && e[0] == '0') return false;
  return dec_persist_ok(dec_shape(d, w));
}

// Their BPTT has a persistent kernel too (dec_persist_bwd.cu: one launch for all T_q steps, 187 instead of 348 launches
// per train step at cfg-2).  It is selected with MMQG_DEC_BWD_PERSIST=1 and is NOT the default: measured on B200 at
// cfg-2 it runs a step in 50 us against 68 us for the launch-per-phase loop, but it holds 128 SMs, so the loss-head
// backward (0.45 ms when it has the GPU to itself) can no longer run beside it -- 1.47 ms against 1.37 ms for
// loss-head backward + decoder BPTT together, 5.60 against 5.44 ms per train step.
static bool dec_bwd_persist_enabled(const mmqg_dims& d, const Ws16& w) {
  const char* e = getenv("MMQG_DEC_BWD_PERSIST");      // read per call: tests switch it inside one process
  if (!(e && e[0] == '1')) return false;
  return dec_persist_enabled(d, w) && dec_bwd_persist_ok(dec_shape(d, w));
}

// With the persistent decoder kernel holding 128 SMs, the loss head cannot hide under the decoder's forward steps
// any more.  Its forward half (the loss) then runs right behind the decoder, and its backward half (d logits -> dH,
// dW_out, db_out) is DEFERRED to the backward call, where it runs group by group, last steps first, on its own stream
// under the launch-per-phase decoder BPTT loop, each BPTT step waiting only for the group that holds its rows.
// Forward and backward of one step must agree on this, so it depends on the shape and the environment only.
static bool lh_deferred(const mmqg_dims& d, const Ws16& w) {
  const char* e = getenv("MMQG_LH_DEFER");
  if (e && e[0] == '0') return false;
  return dec_persist_enabled(d, w);
}
static constexpr int kMaxLhGroups = 16;
static cudaEvent_t ev_lh(int g);      // defined behind AuxStream

// fp32 parameters -> packed bf16 caches + summed biases + concatenated attention bias.
// Two halves so that only what the text encoder needs sits in front of it on the caller's
// stream; the rest is packed on an auxiliary stream beside the text encoder.
static int pack_weights_text(const mmqg_dims& d, const mmqg_tensors& P, Ws16& w, cudaStream_t st) {
  const int H = d.H, G = 4 * d.H, E = d.E, Ep = w.Ep;
  for (int l = 0; l < d.L; ++l) {
    const int I = l == 0 ? E : H, Ip = l == 0 ? Ep : H;
    MMQG_TRY(cvt_f32_bf16_2d(P.text_w_ih[l], I, w.wt_ih[l], Ip, G, I, Ip, st));
    if (!persist_text(d)) MMQG_TRY(cvt_f32_bf16_2d(P.text_w_hh[l], H, w.wt_hh[l], H, G, H, H, st));   // per-step path only
    MMQG_TRY(add2(P.text_b_ih[l], P.text_b_hh[l], w.bsum_text[l], G, st));
  }
  if (persist_text(d))      // forward layout only: this sits in front of the text encoder on the critical path
    for (int l = 0; l < d.L; ++l) MMQG_TRY(pack_rec(P.text_w_hh[l], w.wtp_f[l], nullptr, d.B, H, st));
  return 0;
}
static int pack_weights_rest(const mmqg_dims& d, const mmqg_tensors& P, Ws16& w, cudaStream_t st) {
  const int H = d.H, G = 4 * d.H, Hv = d.H_v, Gv = 4 * d.H_v, E = d.E, Ep = w.Ep, Q = d.E + d.H;
  const int C = d.H + d.H_a + d.H_v, X0 = E + C, Sp = w.Sp;
  for (int l = 0; l < d.L; ++l) {
    MMQG_TRY(cvt_f32_bf16_2d(P.dec_w_hh[l], H, w.wd_cat[l], H + (l == 0 ? C : H), G, H, H, st));
    if (l > 0) MMQG_TRY(cvt_f32_bf16_2d(P.dec_w_ih[l], H, w.wd_cat[l] + H, 2 * H, G, H, H, st));
    MMQG_TRY(add2(P.dec_b_ih[l], P.dec_b_hh[l], w.bsum_dec[l], G, st));
  }
  MMQG_TRY(cvt_f32_bf16_2d(P.vid_w_ih, d.F_v, w.wv_ih, d.F_v, Gv, d.F_v, d.F_v, st));
  MMQG_TRY(cvt_f32_bf16_2d(P.vid_w_hh, Hv, w.wv_hh, Hv, Gv, Hv, Hv, st));
  MMQG_TRY(add2(P.vid_b_ih, P.vid_b_hh, w.bsum_vid, Gv, st));
  MMQG_TRY(cvt_f32_bf16_2d(P.dec_w_ih[0], X0, w.wd_e, Ep, G, E, Ep, st));
  MMQG_TRY(cvt_f32_bf16_2d(P.dec_w_ih[0] + E, X0, w.wd_cat[0] + H, H + C, G, C, C, st));
  MMQG_CUDA(cudaMemsetAsync(w.wa_e, 0, sizeof(b16) * (size_t)Sp * Ep, st));
  MMQG_CUDA(cudaMemsetAsync(w.wa_h, 0, sizeof(b16) * (size_t)Sp * H, st));
  MMQG_CUDA(cudaMemsetAsync(w.attn_b_cat, 0, sizeof(float) * Sp, st));
  const int off[3] = {0, d.TM, d.TM + d.AM}, len[3] = {d.TM, d.AM, d.AM};
  for (int i = 0; i < 3; ++i) {
    MMQG_TRY(cvt_f32_bf16_2d(P.attn_w[i], Q, w.wa_e + (size_t)off[i] * Ep, Ep, len[i], E, Ep, st));
    MMQG_TRY(cvt_f32_bf16_2d(P.attn_w[i] + E, Q, w.wa_h + (size_t)off[i] * H, H, len[i], H, H, st));
    MMQG_CUDA(cudaMemcpyAsync(w.attn_b_cat + off[i], P.attn_b[i], sizeof(float) * len[i], cudaMemcpyDeviceToDevice, st));
  }
  MMQG_TRY(cvt_f32_bf16_2d(P.out_w, H, w.wo, H, d.V, H, H, st));
  if (dec_persist_enabled(d, w))
    for (int l = 0; l < d.L; ++l) {
      MMQG_TRY(pack_whh(P.dec_w_hh[l], w.wdp_hh[l], nullptr, H, st));
      MMQG_TRY(pack_rows_gate16(l == 0 ? P.dec_w_ih[0] + E : P.dec_w_ih[l], l == 0 ? X0 : H, l == 0 ? C : H, l == 0 ? C : H, w.wdp_in[l], H, st));
    }
  if (dec_bwd_persist_enabled(d, w) && !g_fwd_only) {
    for (int l = 0; l < d.L; ++l)
      MMQG_TRY(pack_decb_weights(P.dec_w_hh[l], H, l == 0 ? P.dec_w_ih[0] + E : P.dec_w_ih[l], l == 0 ? X0 : H, l == 0 ? C : H, l, w.wdpb[l], st));
    MMQG_CUDA(cudaMemsetAsync(w.wahT, 0, sizeof(b16) * (size_t)H * w.Spp, st));
    for (int i = 0; i < 3; ++i) MMQG_TRY(transpose_f32_bf16(P.attn_w[i] + E, Q, len[i], H, w.wahT + off[i], w.Spp, st));
  }
  if (persist_video(d)) MMQG_TRY(pack_rec(P.vid_w_hh, w.wvp_f, w.wvp_b, d.B, Hv, st));
  if (persist_text(d))      // BPTT layouts of the text layers: not needed before the backward pass
    for (int l = 0; l < d.L; ++l) MMQG_TRY(pack_rec(P.text_w_hh[l], nullptr, w.wtp_b[l], d.B, H, st));
  return 0;
}

// Internal auxiliary streams (created lazily; every use forks from and joins back onto the
// caller's stream, so the calls stay CUDA-graph capturable).  They let (a) the hoisted
// weight-gradient products run beside the latency-bound persistent BPTT kernels, which occupy
// only 64 of the 148 SMs, and (b) consecutive LSTM layers overlap chunk by chunk in time.
// Priorities: the serial chains (persistent recurrent kernels, decoder step loops) are latency-bound
// and own the critical path, the hoisted products are throughput work that only has to finish by
// the end of the step -- so the chain streams (s[0..NS-3] and `chain`, onto which the work of the
// caller's stream is moved for the duration of a call) get a high stream priority and the
// auxiliary streams s[NS-2] (loss head) and s[NS-1] (hoisted products) the lowest: when SMs free up,
// waiting chain CTAs are placed first.  The short products BETWEEN the chunks of the pipelined text
// encoder (input projection of the next layer / input gradient for the layer below) go to the
// streams g[l] with the very highest priority: two persistent kernels occupy 128 of the 148 SMs
// with ~200 KB of shared memory each, so such a product can only run where a persistent kernel has
// just exited, and it must win those SMs against the next waiting persistent kernel -- otherwise it
// crawls on the 20 spare SMs for 50-100 us while the layer that needs it idles (measured with
// tools/ktrace.py).  MMQG_PRIO=0 creates all streams alike.
struct AuxStream {
  static constexpr int NS = 4, NE = 288;
  cudaStream_t s[NS] = {};
  cudaStream_t g[NS] = {};
  cudaStream_t chain = nullptr;
  cudaEvent_t ev[NE] = {};
  bool ready = false, prio = false;
  int init() {
    if (ready) return 0;
    const char* e = getenv("MMQG_PRIO");
    prio = !(e && e[0] == '0');
    int lo = 0, hi = 0;
    MMQG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least (numerically largest), hi = greatest
    // pipeline stages of the text encoder: the LATER a stage, the higher its priority (a ready chunk
    // of a downstream layer must not wait behind the head layer running ahead): chain = stage 0,
    // s[0] = stage 1, s[1] = stage 2.  Only two persistent kernels fit on the SMs at a time.
    auto clampp = [&](int p) { return p < hi ? hi : (p > lo ? lo : p); };
    for (int i = 0; i < NS; ++i) {
      MMQG_CUDA(cudaStreamCreateWithPriority(&s[i], cudaStreamNonBlocking, prio ? (i >= NS - 2 ? lo : clampp(hi + 2 - i)) : lo));
      MMQG_CUDA(cudaStreamCreateWithPriority(&g[i], cudaStreamNonBlocking, prio ? hi : lo));
    }
    MMQG_CUDA(cudaStreamCreateWithPriority(&chain, cudaStreamNonBlocking, prio ? clampp(hi + 3) : lo));
    for (auto& e2 : ev) MMQG_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    ready = true;
    return 0;
  }
  // move the caller's stream onto the high-priority chain stream for one call ...
  int enter(cudaStream_t user, cudaStream_t* st) {
    MMQG_TRY(init());
    if (!prio) { *st = user; return 0; }
    MMQG_CUDA(cudaEventRecord(ev[12], user));
    MMQG_CUDA(cudaStreamWaitEvent(chain, ev[12], 0));
    *st = chain;
    return 0;
  }
  // ... and back: everything issued during the call is ordered on the caller's stream again
  int leave(cudaStream_t user, cudaStream_t st) {
    if (st == user) return 0;
    MMQG_CUDA(cudaEventRecord(ev[13], st));
    MMQG_CUDA(cudaStreamWaitEvent(user, ev[13], 0));
    return 0;
  }
  // Error path of an entry point: whatever was already enqueued on the internal streams is joined back
  // onto the caller's stream (the header promises single-stream semantics, and an unjoined fork would
  // invalidate a CUDA-graph capture with a confusing error).  Best effort: CUDA errors here are dropped,
  // the caller returns the ORIGINAL status.  While capturing, only streams that joined the capture are touched.
  void join_after_error(cudaStream_t user, cudaStream_t st) {
    if (!ready) return;
    cudaStreamCaptureStatus ucap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(user, &ucap);
    auto join = [&](cudaStream_t s) {
      if (!s || s == user) return;
      cudaStreamCaptureStatus c = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &c);
      if (ucap == cudaStreamCaptureStatusActive && c != cudaStreamCaptureStatusActive) return;
      if (cudaEventRecord(ev[14], s) == cudaSuccess) cudaStreamWaitEvent(user, ev[14], 0);
    };
    for (int i = 0; i < NS; ++i) { join(s[i]); join(g[i]); }
    join(chain);
    (void)st;
    cudaGetLastError();
  }
};
static AuxStream g_aux;
static constexpr int kMaxChunks = 16;
// event slots: [0,16) misc, [16,32) loss-head groups of the deferred backward, then 64 each (NS layers x kMaxChunks):
// forward done(l,c), backward done(l,c), and the two hand-overs between a layer stream and its product stream g[l]
static cudaEvent_t ev_lh(int g) { return g_aux.ev[16 + g]; }
static cudaEvent_t ev_fwd(int l, int c) { return g_aux.ev[32 + l * kMaxChunks + c]; }
static cudaEvent_t ev_bwd(int l, int c) { return g_aux.ev[96 + l * kMaxChunks + c]; }
static cudaEvent_t ev_gf(int l, int c) { return g_aux.ev[160 + l * kMaxChunks + c]; }
static cudaEvent_t ev_gb(int l, int c) { return g_aux.ev[224 + l * kMaxChunks + c]; }
static bool split_products() {
  static const bool on = []() { const char* e = getenv("MMQG_GSTREAMS"); return !(e && e[0] == '0'); }();
  return on && g_aux.prio;
}

// Number of time chunks the text-encoder layers are pipelined over (1 = layer after layer).
// Layer l can run chunk c as soon as layer l-1 has finished chunk c, so with the layers on
// separate streams the serial chain shrinks from L*T_t to about (NC+L-1)/NC * T_t steps.
static int text_chunks(const mmqg_dims& d) {
  // read on every call (tests switch schedules inside one process); getenv is noise next to a step's launches
  const char* e = getenv("MMQG_CHUNKS");
  int want = e ? atoi(e) : 6;
  if (want < 1) want = 1;
  if (want > kMaxChunks) want = kMaxChunks;
  if (persist_kind(d.B, d.H) != 1 || d.L < 2 || d.L > AuxStream::NS) return 1;
  if (2 * lstm_persist_fwd_ctas(d.B, d.H) > device_sms()) return 1;      // two layer kernels cannot be co-resident: nothing to pipeline
  int nc = want;
  while (nc > 1 && d.T_t / nc < 8) --nc;
  return nc;
}

// audio pass-through + video LSTM -> zero-padded attention memories (train.py:155-157)
static int video_forward16(const mmqg_dims& d, const mmqg_batch& bt, Ws16& w, cudaStream_t st) {
  const int B = d.B, Hv = d.H_v, Gv = 4 * d.H_v;
  if (g_len.on) {
    MMQG_REQUIRE(persist_video(d), "variable-length batches need the persistent recurrent kernels (B=%d H_v=%d)", B, Hv);
    MMQG_TRY(audio_pad(bt.audio, w.m_aud, bt.n_frames, B, d.T_v, d.AM, d.H_a, st));
    MMQG_CUDA(cudaMemsetAsync(w.m_vid, 0, sizeof(float) * (size_t)B * d.AM * Hv, st));      // rows >= n_frames stay zero
    MMQG_CUDA(cudaMemsetAsync(w.m_vid16, 0, sizeof(b16) * (size_t)B * d.AM * Hv, st));
  } else {
    MMQG_CUDA(cudaMemcpy2DAsync(w.m_aud, sizeof(float) * (size_t)d.AM * d.H_a, bt.audio,
                                sizeof(float) * (size_t)d.T_v * d.H_a, sizeof(float) * (size_t)d.T_v * d.H_a, B,
                                cudaMemcpyDeviceToDevice, st));
  }
  // video LSTM (encoder.py:69): frames to time-major bf16 (right-aligned when lengths vary), ONE input
  // projection over all T_v*B rows, then one persistent launch for the T_v recurrent steps
  MMQG_TRY(frames_to_time_major_bf16(bt.frames, w.frames16, g_len.on ? g_len.shift_v : nullptr, B, d.T_v, d.F_v, st));
  if (persist_video(d)) {
    MMQG_TRY(Tc(w.frames16, d.F_v, false, w.wv_ih, d.F_v, false, d.T_v * B, Gv, d.F_v, w.acts_v, Gv).bias(w.bsum_vid).run(st));
    MMQG_CUDA(cudaMemsetAsync(w.hs_v, 0, sizeof(b16) * (size_t)B * Hv, st));
    MMQG_CUDA(cudaMemsetAsync(w.cs_v, 0, sizeof(float) * (size_t)B * Hv, st));
    tl_ktag = 900;
    // the kernel writes the fp32 memory rows and their bf16 copy itself: no conversion pass over the padded (B, AM, H_v) tensor
    return rec_fwd(w.acts_v, w.cs_v, w.hs_v, w.wvp_f, w.m_vid, w.m_vid16, (long long)d.AM * Hv, w.flags_v, d.T_v, B, Hv, st, len_video());
  }
  for (int t = 0; t < d.T_v; ++t) {
    StepGemmScope step_scope;
    float* acts = w.acts_v + (size_t)t * B * Gv;
    Tc g(w.frames16 + (size_t)t * B * d.F_v, d.F_v, false, w.wv_ih, d.F_v, false, B, Gv, d.F_v, acts, Gv);
    g.bias(w.bsum_vid);
    if (t > 0) g.second(w.hs_v + (size_t)t * B * Hv, Hv, w.wv_hh, Hv, Hv);
    MMQG_TRY(g.run(st));
    MMQG_TRY(lstm_pointwise_fwd_bf16(acts, Gv, t > 0 ? w.cs_v + (size_t)t * B * Hv : nullptr, Hv,
                                     w.cs_v + (size_t)(t + 1) * B * Hv, Hv, w.hs_v + (size_t)(t + 1) * B * Hv, Hv,
                                     w.m_vid + (size_t)t * Hv, d.AM * Hv, B, Hv, st));
  }
  return cvt_f32_bf16_2d(w.m_vid, Hv, w.m_vid16, Hv, (long long)B * d.AM, Hv, Hv, st);
}

// text LSTM stack (encoder.py:95-100): hoisted input projection per layer, recurrent part per step
static int text_forward16(const mmqg_dims& d, const mmqg_tensors& P, Ws16& w, cudaStream_t st) {
  const int B = d.B, H = d.H, G = 4 * d.H;
  MMQG_TRY(embedding_gather_bf16(P.emb, w.idx_ctx, w.x0, w.Ep, d.T_t * B, d.E, w.Ep, d.V, st));
  const int NC = text_chunks(d);
  if (g_len.on) {     // memory rows beyond a sample's length are the zero padding of train.py:160
    MMQG_REQUIRE(persist_text(d), "variable-length batches need the persistent recurrent kernels (B=%d H=%d)", B, H);
    MMQG_CUDA(cudaMemsetAsync(w.m_txt16, 0, sizeof(b16) * (size_t)B * d.TM * H, st));
    if (NC <= 1) MMQG_CUDA(cudaMemsetAsync(w.m_txt, 0, sizeof(float) * (size_t)B * d.TM * H, st));
  }
  if (NC > 1) {
    // layers pipelined over NC time chunks, one stream per layer
    MMQG_TRY(g_aux.init());
    const int CL = (d.T_t + NC - 1) / NC;
    MMQG_TRY(Tc(w.x0, w.Ep, false, w.wt_ih[0], w.Ep, false, d.T_t * B, G, w.Ep, w.acts_text[0], G).bias(w.bsum_text[0]).run(st));
    const int n_mt = (B + 127) / 128;
    for (int l = 0; l < d.L; ++l) {
      MMQG_CUDA(cudaMemsetAsync(w.hs_text[l], 0, sizeof(b16) * (size_t)B * H, st));
      MMQG_CUDA(cudaMemsetAsync(w.cs_text[l], 0, sizeof(float) * (size_t)B * H, st));
      MMQG_CUDA(cudaMemsetAsync(w.flags_t[l], 0, sizeof(uint32_t) * (size_t)(d.T_t + 1) * n_mt, st));
    }
    MMQG_CUDA(cudaEventRecord(g_aux.ev[0], st));
    for (int l = 1; l < d.L; ++l) MMQG_CUDA(cudaStreamWaitEvent(g_aux.s[l - 1], g_aux.ev[0], 0));
    for (int c = 0; c < NC; ++c) {
      const int t0 = c * CL, nT = (d.T_t - t0 < CL) ? d.T_t - t0 : CL;
      for (int l = 0; l < d.L; ++l) {
        cudaStream_t s = l == 0 ? st : g_aux.s[l - 1];
        if (l > 0) {
          // input projection of this chunk on the layer's product stream (highest priority), then hand over
          cudaStream_t gs = split_products() ? g_aux.g[l] : s;
          MMQG_CUDA(cudaStreamWaitEvent(gs, ev_fwd(l - 1, c), 0));
          const b16* X = g_drop_p > 0.f ? w.xdrop_text[l - 1] + (size_t)t0 * B * H : w.hs_text[l - 1] + (size_t)(t0 + 1) * B * H;
          MMQG_TRY(Tc(X, H, false, w.wt_ih[l], H, false, nT * B, G, H, w.acts_text[l] + (size_t)t0 * B * G, G)
                       .bias(w.bsum_text[l]).run(gs));
          if (gs != s) {
            MMQG_CUDA(cudaEventRecord(ev_gf(l, c), gs));
            MMQG_CUDA(cudaStreamWaitEvent(s, ev_gf(l, c), 0));
          }
        }
        tl_ktag = l * 16 + c;
        DropSpec dr;           // the kernel writes the dropped copy of its h_t itself (no extra launch between chunks)
        if (g_drop_p > 0.f && l + 1 < d.L) {
          dr.out = w.xdrop_text[l] + (size_t)t0 * B * H; dr.ld = H; dr.seed = g_drop_seed; dr.ctr = g_drop_ctr;
          dr.sid = kSidText + l; dr.base = (unsigned long long)t0 * B * H; dr.p = g_drop_p;
        }
        // arrival counters: one region per chunk of the layer's array, zeroed once above
        MMQG_TRY(lstm_seq_fwd_persist(w.acts_text[l] + (size_t)t0 * B * G, w.cs_text[l] + (size_t)t0 * B * H,
                                      w.hs_text[l] + (size_t)t0 * B * H, w.wtp_f[l],
                                      nullptr, l == d.L - 1 ? w.m_txt16 + (size_t)t0 * H : nullptr, (long long)d.TM * H,
                                      w.flags_t[l] + (size_t)t0 * n_mt, nT, B, H, c > 0 ? 1 : 0, s, dr, false, len_text(t0, l == d.L - 1)));
        MMQG_CUDA(cudaEventRecord(ev_fwd(l, c), s));
      }
    }
    for (int l = 1; l < d.L; ++l) MMQG_CUDA(cudaStreamWaitEvent(st, ev_fwd(l, NC - 1), 0));   // join
  } else
  for (int l = 0; l < d.L; ++l) {
    if (l > 0 && g_drop_p > 0.f)
      MMQG_TRY(dropout_bf16(w.hs_text[l - 1] + (size_t)B * H, w.xdrop_text[l - 1], (long long)d.T_t * B * H, g_drop_seed,
                            g_drop_ctr, kSidText + l - 1, 0, g_drop_p, st));
    const b16* X = l == 0 ? w.x0 : (g_drop_p > 0.f ? w.xdrop_text[l - 1] : w.hs_text[l - 1] + (size_t)B * H);
    const int Ip = l == 0 ? w.Ep : H;
    MMQG_TRY(Tc(X, Ip, false, w.wt_ih[l], Ip, false, d.T_t * B, G, Ip, w.acts_text[l], G).bias(w.bsum_text[l]).run(st));
    if (persist_text(d)) {
      MMQG_CUDA(cudaMemsetAsync(w.hs_text[l], 0, sizeof(b16) * (size_t)B * H, st));
      MMQG_CUDA(cudaMemsetAsync(w.cs_text[l], 0, sizeof(float) * (size_t)B * H, st));
      // the top layer writes the fp32 memory rows and their bf16 copy itself (no conversion pass over the padded (B, TM, H) tensor)
      MMQG_TRY(rec_fwd(w.acts_text[l], w.cs_text[l], w.hs_text[l], w.wtp_f[l], l == d.L - 1 ? w.m_txt : nullptr, l == d.L - 1 ? w.m_txt16 : nullptr,
                       (long long)d.TM * H, w.flags, d.T_t, B, H, st, len_text(0, l == d.L - 1)));
      continue;
    }
    for (int t = 0; t < d.T_t; ++t) {
      StepGemmScope step_scope;
      float* acts = w.acts_text[l] + (size_t)t * B * G;
      if (t > 0)
        MMQG_TRY(Tc(w.hs_text[l] + (size_t)t * B * H, H, false, w.wt_hh[l], H, false, B, G, H, acts, G).accumulate(true).run(st));
      MMQG_TRY(lstm_pointwise_fwd_bf16(acts, G, t > 0 ? w.cs_text[l] + (size_t)t * B * H : nullptr, H,
                                       w.cs_text[l] + (size_t)(t + 1) * B * H, H, w.hs_text[l] + (size_t)(t + 1) * B * H, H,
                                       l == d.L - 1 ? w.m_txt + (size_t)t * H : nullptr, d.TM * H, B, H, st));
    }
  }
  // bf16 attention memories: the pipelined persistent path writes the text memory in bf16 directly;
  // every other path produced fp32 rows, converted here (rows beyond T_t / T_v are never read)
  if (NC <= 1 && !persist_text(d)) MMQG_TRY(cvt_f32_bf16_2d(w.m_txt, H, w.m_txt16, H, (long long)B * d.TM, H, H, st));
  // decoder state slab 0 := encoder final state (train.py:169)
  const size_t n = (size_t)B * H;
  for (int l = 0; l < d.L; ++l) {
    MMQG_CUDA(cudaMemcpyAsync(w.hs_dec[l], w.hs_text[l] + (size_t)d.T_t * n, sizeof(b16) * n, cudaMemcpyDeviceToDevice, st));
    MMQG_CUDA(cudaMemcpyAsync(w.cs_dec[l], w.cs_text[l] + (size_t)d.T_t * n, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

static int train_forward_bf16_on(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                                 size_t workspace_bytes, float* loss_out, int want_grads, mmqg_tensors* grads, float grad_scale,
                                 float dropout_p, unsigned long long seed, cudaStream_t st);

int train_forward_bf16(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                       size_t workspace_bytes, float* loss_out, int want_grads, mmqg_tensors* grads, float grad_scale,
                       float dropout_p, unsigned long long seed, cudaStream_t user) {
  // validate before any work is moved onto the internal streams
  MMQG_TRY(check_dims_bf16(d));
  const size_t need = train_workspace_bytes_bf16(d, d.T_q);
  if (need > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, need);
  cudaStream_t st;
  MMQG_TRY(g_aux.enter(user, &st));
  const int rc = train_forward_bf16_on(d, P, bt, workspace, workspace_bytes, loss_out, want_grads, grads, grad_scale, dropout_p, seed, st);
  if (rc != 0) { g_aux.join_after_error(user, st); return rc; }
  return g_aux.leave(user, st);
}

static int train_forward_bf16_on(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                                 size_t workspace_bytes, float* loss_out, int want_grads, mmqg_tensors* grads, float grad_scale,
                                 float dropout_p, unsigned long long seed, cudaStream_t st) {
  MMQG_TRY(check_dims_bf16(d));
  g_drop_p = dropout_p;
  g_drop_seed = seed;
  g_fwd_only = false;
  Ws16 w = carve16(d, d.T_q, workspace);
  if (w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  const int B = d.B, H = d.H, G = 4 * d.H, C = d.H + d.H_a + d.H_v, R = d.T_q * B, Sp = w.Sp, Ep = w.Ep;

  g_drop_ctr = w.seed_ctr;
  MMQG_TRY(set_len_state(d, bt, w));
  mark(0, st);
  if (dropout_p > 0.f) MMQG_TRY(bump_counter(w.seed_ctr, st));     // this call's masks: seed + (calls so far)
  MMQG_TRY(g_aux.init());
  cudaStream_t ax = g_aux.s[AuxStream::NS - 1], lh = g_aux.s[AuxStream::NS - 2];
  if (g_len.on) MMQG_TRY(prep_lengths(bt.ctx_len, bt.tgt_len, bt.n_frames, w.shift_t, w.shift_v, w.row_w, B, d.T_t, d.T_v, d.T_q, st));
  MMQG_TRY(build_indices(bt.context, bt.target, w.idx_ctx, w.idx_dec, w.tgt_tm, B, d.T_t, d.T_q, st, g_len.on ? w.shift_t : nullptr));
  // fork: everything the text encoder does not need -- the other weight packs, the video LSTM
  // and the decoder's hoisted embedding-column products over all teacher-forced steps -- runs on
  // an auxiliary stream beside the text encoder and is joined in front of the decoder loop
  MMQG_CUDA(cudaEventRecord(g_aux.ev[8], st));
  MMQG_CUDA(cudaStreamWaitEvent(ax, g_aux.ev[8], 0));
  MMQG_TRY(pack_weights_rest(d, P, w, ax));
  MMQG_TRY(video_forward16(d, bt, w, ax));
  MMQG_TRY(embedding_gather_bf16(P.emb, w.idx_dec, w.e_dec, Ep, R, d.E, Ep, d.V, ax));
  MMQG_TRY(Tc(w.e_dec, Ep, false, w.wd_e, Ep, false, R, G, Ep, w.acts_dec[0], G).bias(w.bsum_dec[0]).run(ax));
  MMQG_TRY(Tc(w.e_dec, Ep, false, w.wa_e, Ep, false, R, Sp, Ep, w.attn_all, Sp).bias(w.attn_b_cat).run(ax));
  MMQG_CUDA(cudaEventRecord(g_aux.ev[9], ax));
  MMQG_TRY(pack_weights_text(d, P, w, st));
  mark(1, st);
  MMQG_TRY(text_forward16(d, P, w, st));
  mark(2, st);
  MMQG_CUDA(cudaStreamWaitEvent(st, g_aux.ev[9], 0));
  mark(3, st);
  AttnShape as = attn_shape16(d, w);
  as.ldctx16 = C;
  // loss head over rows [r0, r0+n) of h_top (row = t*B + b): fused vocabulary projection + log-softmax + NLL
  // and its backward (vocab_nll.cu); logits are never stored (decoder.py:106, train.py:174).
  const b16* htop = w.hs_dec[d.L - 1] + (size_t)B * H;
  const float dscale = want_grads ? grad_scale / (float)B : 0.f;
  const bool defer_bwd = lh_deferred(d, w);
  auto loss_head = [&](int r0, int n, bool first, cudaStream_t s) -> int {
    PdlScope no_pdl(false);
    // forward: logits tiles live in tensor memory only; per-row (max, sum exp, target logit) -> lse, NLL
    MMQG_TRY(vocab_nll_fwd(htop + (size_t)r0 * H, H, w.wo, H, P.out_b, w.tgt_tm + r0, g_len.on ? w.row_w + r0 : nullptr, n, d.V, H, dscale,
                           w.nll + r0, w.lse + r0, w.rscale + r0, w.stat_a, w.stat_b, w.tgt_logit + r0, s));
    // backward: tiles recomputed, bf16 d-logits in L2-sized row chunks -> dH, dW_out (+=), db_out (+=)
    if (want_grads && !defer_bwd)
      MMQG_TRY(vocab_nll_bwd(htop + (size_t)r0 * H, H, w.wo, H, P.out_b, w.tgt_tm + r0, w.lse + r0, w.rscale + r0, n, d.V, H, w.dlogits16,
                             w.Vp, w.Rc, w.dh_part, w.dhtop + (size_t)r0 * H, H, grads->out_w, grads->out_b, !first, s));
    return 0;
  };
  // The loss head of the steps already decoded runs on its own stream under the remaining
  // (latency-bound, 64-CTA) decoder steps: groups of `lh_steps` whole steps, at most Rc rows.
  static const bool lh_env = []() { const char* e = getenv("MMQG_LOSS_OVERLAP"); return !(e && e[0] == '0'); }();
  static const int lh_group_env = []() { const char* e = getenv("MMQG_LH_STEPS"); return e ? atoi(e) : 0; }();
  const int lh_steps = lh_group_env > 0 ? lh_group_env : (w.Rc / B > 1 ? w.Rc / B : 1);      // whole steps per loss-head group
  static const int lh_tail = []() { const char* e = getenv("MMQG_LH_TAIL"); return e ? atoi(e) : 1; }();
  const bool lh_overlap = lh_env && lh_steps >= 1 && d.T_q > 1 && !defer_bwd;     // deferred: one decoder launch, then the loss
  int lh_done = 0;      // steps whose loss head has been issued
  PdlScope pdl_scope(pdl_enabled());      // the dependent launches below overlap prologue and tail
  L2WindowScope l2_scope(w.m_txt16, attn_l2_window(d, w));     // attention memories stay in L2 across the T_q steps
  // MMQG_STEP_FUSE (launch-per-phase loop only): 2 = split-K product + summing cell kernel (default), 0 = product + cell kernel
  static const int step_mode_env = []() { const char* e = getenv("MMQG_STEP_FUSE"); return e ? atoi(e) : 2; }();
  const int step_mode = step_mode_env;
  const bool dec_persist = dec_persist_enabled(d, w);
  DecPersistArgs dpa{};
  int seg_start = 0;
  if (dec_persist) {
    dpa.shape = dec_shape(d, w);
    dpa.Tq = d.T_q; dpa.attn_all = w.attn_all; dpa.ctx16 = w.ctx16;
    for (int l = 0; l < d.L; ++l) {
      dpa.acts[l] = w.acts_dec[l]; dpa.cs[l] = w.cs_dec[l]; dpa.hs[l] = w.hs_dec[l];
      dpa.hdrop[l] = (g_drop_p > 0.f && l + 1 < d.L) ? w.hdrop_dec[l] : nullptr;
      dpa.bias[l] = l == 0 ? nullptr : w.bsum_dec[l];
      dpa.w_hh[l] = w.wdp_hh[l]; dpa.w_in[l] = w.wdp_in[l];
    }
    dpa.wa_h = w.wa_h; dpa.m_txt16 = w.m_txt16; dpa.m_vid16 = w.m_vid16; dpa.m_aud = w.m_aud; dpa.flags = w.flags_dec;
    dpa.drop_p = g_drop_p; dpa.seed = g_drop_seed; dpa.ctr = g_drop_ctr; dpa.sid0 = kSidDec;
    MMQG_CUDA(cudaMemsetAsync(w.flags_dec, 0, sizeof(uint32_t) * dec_persist_flag_words(dpa.shape, d.T_q), st));
  }
  for (int t = 0; t < d.T_q; ++t) {
    if (dec_persist) {
      // the steps of one loss-head group (or all of them) run inside ONE persistent launch
      const bool group_end = lh_overlap && (t + 1 - lh_done == lh_steps || t == d.T_q - 1 - lh_tail);
      if (group_end || t == d.T_q - 1) {
        PdlScope no_pdl(false);
        MMQG_TRY(dec_seq_fwd_persist(dpa, seg_start, t + 1 - seg_start, st));
        seg_start = t + 1;
        if (lh_overlap) {
          MMQG_CUDA(cudaEventRecord(g_aux.ev[10], st));
          MMQG_CUDA(cudaStreamWaitEvent(lh, g_aux.ev[10], 0));
          MMQG_TRY(loss_head(lh_done * B, (t + 1 - lh_done) * B, lh_done == 0, lh));
          lh_done = t + 1;
          if (t == d.T_q - 1) MMQG_CUDA(cudaEventRecord(g_aux.ev[11], lh));
        }
      }
      continue;
    }
    StepGemmScope step_scope;
    const b16* htop_prev = w.hs_dec[d.L - 1] + (size_t)t * B * H;
    float* sc = w.attn_all + (size_t)t * B * Sp;
    b16* ctx = w.ctx16 + (size_t)t * B * C;
    MMQG_TRY(Tc(htop_prev, H, false, w.wa_h, H, false, B, Sp, H, sc, Sp).accumulate(true).run(st));
    as.ctx16 = ctx;
    MMQG_TRY(attn_fwd(sc, Sp, w.m_txt, w.m_aud, w.m_vid, w.ctx_tmp, C, as, st));
    for (int l = 0; l < d.L; ++l) {
      float* acts = w.acts_dec[l] + (size_t)t * B * G;
      const b16* hprev = w.hs_dec[l] + (size_t)t * B * H;
      const b16* xin = l == 0 ? ctx : (g_drop_p > 0.f ? w.hdrop_dec[l - 1] + (size_t)t * B * H : w.hs_dec[l - 1] + (size_t)(t + 1) * B * H);
      const int Kin = l == 0 ? C : H, Nl = H + Kin;      // wd_cat[l] = [W_hh | W_in], row pitch Nl
      DropSpec dr;
      if (g_drop_p > 0.f && l + 1 < d.L) {
        dr.out = w.hdrop_dec[l] + (size_t)t * B * H; dr.ld = H; dr.seed = g_drop_seed; dr.ctr = g_drop_ctr; dr.sid = kSidDec + l;
        dr.base = (unsigned long long)t * B * H; dr.p = g_drop_p;
      }
      float* c_prev = w.cs_dec[l] + (size_t)t * B * H;
      float* c_new = w.cs_dec[l] + (size_t)(t + 1) * B * H;
      b16* h_new = w.hs_dec[l] + (size_t)(t + 1) * B * H;
      if (step_mode == 2) {  // split-K over twice as many CTAs (the product is bound by per-SM operand ingest); the
                             // cell kernel sums the partials, the bias and (layer 0) the hoisted embedding product
        MMQG_TRY(Tc(xin, Kin, false, w.wd_cat[l] + H, Nl, false, B, G, Kin, w.gpart, G).second(hprev, H, w.wd_cat[l], Nl, H)
                     .split(2, (long long)B * G).run(st));
        PreSpec pre;
        pre.part = w.gpart; pre.n_part = 2; pre.ld = G; pre.stride = (long long)B * G;
        pre.bias = l == 0 ? nullptr : w.bsum_dec[l]; pre.add_gates = l == 0 ? 1 : 0;
        MMQG_TRY(lstm_pointwise_fwd_bf16(acts, G, c_prev, H, c_new, H, h_new, H, nullptr, 0, B, H, st, dr, pre));
        continue;
      }
      if (l == 0)
        MMQG_TRY(Tc(xin, Kin, false, w.wd_cat[0] + H, Nl, false, B, G, Kin, acts, G).second(hprev, H, w.wd_cat[0], Nl, H)
                     .accumulate(true).run(st));
      else
        MMQG_TRY(Tc(xin, Kin, false, w.wd_cat[l] + H, Nl, false, B, G, Kin, acts, G).second(hprev, H, w.wd_cat[l], Nl, H)
                     .bias(w.bsum_dec[l]).run(st));
      MMQG_TRY(lstm_pointwise_fwd_bf16(acts, G, c_prev, H, c_new, H, h_new, H, nullptr, 0, B, H, st, dr));
    }
    // groups of lh_steps steps; the last lh_tail steps form their own group so that only a short loss-head tail
    // is exposed after the decoder loop
    if (lh_overlap && (t + 1 - lh_done == lh_steps || t == d.T_q - 1 || t == d.T_q - 1 - lh_tail)) {
      MMQG_CUDA(cudaEventRecord(g_aux.ev[10], st));
      MMQG_CUDA(cudaStreamWaitEvent(lh, g_aux.ev[10], 0));
      MMQG_TRY(loss_head(lh_done * B, (t + 1 - lh_done) * B, lh_done == 0, lh));
      lh_done = t + 1;
      if (t == d.T_q - 1) MMQG_CUDA(cudaEventRecord(g_aux.ev[11], lh));
    }
  }
  mark(4, st);
  if (!lh_overlap) MMQG_TRY(loss_head(0, R, true, st));
  else MMQG_CUDA(cudaStreamWaitEvent(st, g_aux.ev[11], 0));      // join the loss-head stream
  MMQG_TRY(sum_scale(w.nll, R, 1.0f / (float)B, loss_out, st));
  mark(5, st);
  return 0;
}

// ---- greedy decode on the tensor-core path (reference train.py:81-110, evaluate.py:45-80) ------------
// Encoder as in training (no dropout), then max_len decoder steps feeding back argmax(logits)
// (lowest index on ties).  bf16 operands: tokens can leave the fp32 oracle's path at near-ties --
// the fp32 mode carries the token-exact claim, this mode the throughput.
static int greedy_decode_bf16_on(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, Ws16& w, int64_t* tokens_out,
                                 int max_len, cudaStream_t st, int sample, unsigned long long seed);

int greedy_decode_bf16(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace, size_t workspace_bytes,
                       int64_t* tokens_out, int max_len, cudaStream_t user, int sample, unsigned long long seed) {
  MMQG_TRY(check_dims_bf16(d));
  Ws16 w = carve16(d, max_len, workspace);
  if (w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  g_drop_p = 0.f;
  g_drop_seed = 0;
  g_drop_ctr = w.seed_ctr;
  g_fwd_only = true;
  MMQG_TRY(set_len_state(d, bt, w));
  cudaStream_t st;
  MMQG_TRY(g_aux.enter(user, &st));
  const int rc = greedy_decode_bf16_on(d, P, bt, w, tokens_out, max_len, st, sample, seed);
  if (rc != 0) { g_aux.join_after_error(user, st); return rc; }
  return g_aux.leave(user, st);
}

static int greedy_decode_bf16_on(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, Ws16& w, int64_t* tokens_out,
                                 int max_len, cudaStream_t st, int sample, unsigned long long seed) {
  const int B = d.B, H = d.H, G = 4 * d.H, C = d.H + d.H_a + d.H_v, Sp = w.Sp, Ep = w.Ep;
  cudaStream_t ax = g_aux.s[AuxStream::NS - 1];
  if (g_len.on) MMQG_TRY(prep_lengths(bt.ctx_len, nullptr, bt.n_frames, w.shift_t, w.shift_v, w.row_w, B, d.T_t, d.T_v, 0, st));
  MMQG_TRY(build_indices(bt.context, nullptr, w.idx_ctx, nullptr, nullptr, B, d.T_t, 0, st, g_len.on ? w.shift_t : nullptr));
  MMQG_CUDA(cudaEventRecord(g_aux.ev[8], st));
  MMQG_CUDA(cudaStreamWaitEvent(ax, g_aux.ev[8], 0));
  MMQG_TRY(pack_weights_rest(d, P, w, ax));
  MMQG_TRY(video_forward16(d, bt, w, ax));
  MMQG_CUDA(cudaEventRecord(g_aux.ev[9], ax));
  MMQG_TRY(pack_weights_text(d, P, w, st));
  MMQG_TRY(text_forward16(d, P, w, st));
  MMQG_CUDA(cudaStreamWaitEvent(st, g_aux.ev[9], 0));
  AttnShape as = attn_shape16(d, w);
  as.ldctx16 = C;
  MMQG_TRY(fill_i64(w.idx_cur, B, 1, st));                      // <start>, train.py:84
  PdlScope pdl_scope(pdl_enabled());
  // K6: one decoder step = embedding gather, the two embedding-column products, ONE launch of the persistent
  // decoder-step kernel (scores, attention heads, the three LSTM cells; dec_persist.cu with T = 1) and the
  // vocabulary projection with the arg-max in its epilogue
  // (measured at B = 1024, two waves of 128-row groups: 8.2 ms per batch against 7.9 ms for the launch-per-phase step,
  // so batches that need more than one wave keep the latter)
  const bool dec_persist = dec_persist_enabled(d, w) && dec_persist_waves(dec_shape(d, w)) == 1;
  DecPersistArgs dpa{};
  if (dec_persist) {
    dpa.shape = dec_shape(d, w);
    dpa.Tq = max_len; dpa.attn_all = w.attn_all; dpa.ctx16 = w.ctx16;
    for (int l = 0; l < d.L; ++l) {
      dpa.acts[l] = w.acts_dec[l]; dpa.cs[l] = w.cs_dec[l]; dpa.hs[l] = w.hs_dec[l];
      dpa.hdrop[l] = nullptr; dpa.bias[l] = l == 0 ? nullptr : w.bsum_dec[l];
      dpa.w_hh[l] = w.wdp_hh[l]; dpa.w_in[l] = w.wdp_in[l];
    }
    dpa.wa_h = w.wa_h; dpa.m_txt16 = w.m_txt16; dpa.m_vid16 = w.m_vid16; dpa.m_aud = w.m_aud; dpa.flags = w.flags_dec;
    dpa.drop_p = 0.f; dpa.sid0 = kSidDec;
    MMQG_CUDA(cudaMemsetAsync(w.flags_dec, 0, sizeof(uint32_t) * dec_persist_flag_words(dpa.shape, max_len), st));
  }
  for (int t = 0; t < max_len; ++t) {
    StepGemmScope step_scope;
    b16* e_t = w.e_dec + (size_t)t * B * Ep;
    {
      PdlScope no_pdl(false);
      MMQG_TRY(embedding_gather_bf16(P.emb, w.idx_cur, e_t, Ep, B, d.E, Ep, d.V, st));
    }
    const b16* htop_prev = w.hs_dec[d.L - 1] + (size_t)t * B * H;
    float* sc = w.attn_all + (size_t)t * B * Sp;
    b16* ctx = w.ctx16 + (size_t)t * B * C;
    if (dec_persist) {
      PdlScope no_pdl(false);
      MMQG_TRY(Tc(e_t, Ep, false, w.wd_e, Ep, false, B, G, Ep, w.acts_dec[0] + (size_t)t * B * G, G).bias(w.bsum_dec[0]).run(st));
      MMQG_TRY(Tc(e_t, Ep, false, w.wa_e, Ep, false, B, Sp, Ep, sc, Sp).bias(w.attn_b_cat).run(st));
      MMQG_TRY(dec_seq_fwd_persist(dpa, t, 1, st));
    } else {
    MMQG_TRY(Tc(e_t, Ep, false, w.wa_e, Ep, false, B, Sp, Ep, sc, Sp).second(htop_prev, H, w.wa_h, H, H).bias(w.attn_b_cat).run(st));
    as.ctx16 = ctx;
    MMQG_TRY(attn_fwd(sc, Sp, w.m_txt, w.m_aud, w.m_vid, w.ctx_tmp, C, as, st));
    for (int l = 0; l < d.L; ++l) {
      float* acts = w.acts_dec[l] + (size_t)t * B * G;
      const b16* hprev = w.hs_dec[l] + (size_t)t * B * H;
      if (l == 0) {
        MMQG_TRY(Tc(e_t, Ep, false, w.wd_e, Ep, false, B, G, Ep, acts, G).bias(w.bsum_dec[0]).run(st));
        MMQG_TRY(Tc(ctx, C, false, w.wd_cat[0] + H, H + C, false, B, G, C, acts, G).second(hprev, H, w.wd_cat[0], H + C, H)
                     .accumulate(true).run(st));
      } else {
        MMQG_TRY(Tc(w.hs_dec[l - 1] + (size_t)(t + 1) * B * H, H, false, w.wd_cat[l] + H, 2 * H, false, B, G, H, acts, G)
                     .second(hprev, H, w.wd_cat[l], 2 * H, H).bias(w.bsum_dec[l]).run(st));
      }
      PreSpec fwd_only;
      fwd_only.no_save = 1;          // decoding has no backward pass: the activated gates are not kept
      MMQG_TRY(lstm_pointwise_fwd_bf16(acts, G, w.cs_dec[l] + (size_t)t * B * H, H, w.cs_dec[l] + (size_t)(t + 1) * B * H, H,
                                       w.hs_dec[l] + (size_t)(t + 1) * B * H, H, nullptr, 0, B, H, st, DropSpec(), fwd_only));
    }
    }
    const b16* htop = w.hs_dec[d.L - 1] + (size_t)(t + 1) * B * H;
    PdlScope no_pdl(false);
    if (!sample) {     // vocabulary projection with the arg-max in its epilogue: no logits in memory (train.py:106-108)
      MMQG_TRY(vocab_argmax(htop, H, w.wo, H, P.out_b, B, d.V, H, w.stat_a, reinterpret_cast<int*>(w.stat_b), tokens_out + t, max_len,
                            w.idx_cur, st));
      continue;
    }
    for (int r0 = 0; r0 < B; r0 += w.Rs) {
      const int rc = B - r0 < w.Rs ? B - r0 : w.Rs;
      MMQG_TRY(Tc(htop + (size_t)r0 * H, H, false, w.wo, H, false, rc, d.V, H, w.logits, d.V).bias(P.out_b).run(st));
      MMQG_TRY(sample_rows(w.logits, d.V, tokens_out + (size_t)r0 * max_len + t, max_len, w.idx_cur + r0, rc, d.V, seed,
                           (unsigned long long)t, r0, B, st));
    }
  }
  return 0;
}

size_t greedy_workspace_bytes_bf16(const mmqg_dims& d, int max_len) { return carve16(d, max_len, nullptr).bytes; }

// ---- backward -----------------------------------------------------------------------------
struct Bwd16 {
  const mmqg_dims& d; const mmqg_tensors& P; const mmqg_batch& bt; Ws16& w; mmqg_tensors& Gd;
  int B, H, G, E, Ep, Q, C, X0, R, Sp, L;
  long long ps;
  Bwd16(const mmqg_dims& d_, const mmqg_tensors& P_, const mmqg_batch& bt_, Ws16& w_, mmqg_tensors& Gd_)
      : d(d_), P(P_), bt(bt_), w(w_), Gd(Gd_) {
    B = d.B; H = d.H; G = 4 * d.H; E = d.E; Ep = w.Ep; Q = d.E + d.H; C = d.H + d.H_a + d.H_v; X0 = E + C;
    R = d.T_q * B; Sp = w.Sp; L = d.L; ps = (long long)B * H;
  }

  // Deferred backward of the loss head (see lh_deferred()): rows of the steps [t_lo, t_hi) in chunks of Rc rows.
  int loss_head_bwd(int t_lo, int t_hi, bool first, cudaStream_t s) {
    PdlScope no_pdl(false);
    const b16* htop = w.hs_dec[L - 1] + (size_t)B * H;
    const size_t r0 = (size_t)t_lo * B;
    return vocab_nll_bwd(htop + r0 * H, H, w.wo, H, P.out_b, w.tgt_tm + r0, w.lse + r0, w.rscale + r0, (t_hi - t_lo) * B, d.V, H,
                         w.dlogits16, w.Vp, w.Rc, w.dh_part, w.dhtop + r0 * H, H, Gd.out_w, Gd.out_b, !first, s);
  }
  // Groups of decoder steps in the order the BPTT needs them (last steps first): a short first group so that the
  // BPTT can start early, then groups of about Rc rows.  lo[k] = first step of group k (its last is lo[k-1] - 1).
  int lh_groups(int* lo) const {
    int n = 0, hi = d.T_q;
    int size = 2;
    const int cap = w.Rc / B > 1 ? w.Rc / B : 1;
    while (hi > 0 && n < kMaxLhGroups - 1) {
      int take = size < cap ? size : cap;
      if (take > hi) take = hi;
      hi -= take;
      lo[n++] = hi;
      size *= 2;
    }
    if (hi > 0) lo[n - 1] = 0;        // (cannot happen with 15 doubling groups; keeps the cover complete)
    return n;
  }

  // decoder BPTT inside the persistent kernel: one launch for all T_q steps (dec_persist_bwd.cu)
  bool dec_pers() const { return dec_bwd_persist_enabled(d, w); }
  int dec_loop_persist(cudaStream_t st) {
    AttnShape as = attn_shape16(d, w);
    MMQG_CUDA(cudaMemsetAsync(w.ds_all, 0, sizeof(float) * (size_t)R * Sp, st));       // padding columns stay zero
    MMQG_CUDA(cudaMemsetAsync(w.ds16, 0, sizeof(b16) * (size_t)R * Sp, st));
    DecPersistBwdArgs a{};
    a.shape = dec_shape(d, w);
    a.Tq = d.T_q; a.attn_all = w.attn_all; a.ds_all = w.ds_all; a.ds16 = w.ds16; a.dctx_all = w.dctx_all; a.dhtop = w.dhtop;
    for (int l = 0; l < L; ++l) {
      a.acts[l] = w.acts_dec[l]; a.cs[l] = w.cs_dec[l]; a.dg[l] = w.dg_dec[l]; a.dc[l] = w.dc[l]; a.dh_rec[l] = w.dh_rec[l];
      a.wT[l] = w.wdpb[l];
    }
    a.wahT = w.wahT; a.Spp = w.Spp;
    a.m_txt16 = w.m_txt16; a.m_vid16 = w.m_vid16; a.m_aud = w.m_aud; a.flags = w.flags_decb;
    a.drop_p = g_drop_p; a.seed = g_drop_seed; a.ctr = g_drop_ctr; a.sid0 = kSidDec;
    MMQG_CUDA(cudaMemsetAsync(w.flags_decb, 0, sizeof(uint32_t) * dec_bwd_persist_flag_words(a.shape, d.T_q), st));
    {
      PdlScope no_pdl(false);
      MMQG_TRY(dec_seq_bwd_persist(a, 0, d.T_q, st));
    }
    return attn_dmem(w.attn_all, Sp, w.dctx_all, C, w.dm_txt, w.dm_vid, d.T_q, as, st);
  }
  // d loss / d (final encoder state of layer l) = the decoder's gradient w.r.t. its initial state (train.py:169) plus, for
  // the top layer, the step-0 attention query: split-K partials of the launch-per-phase loop, one full sum from the kernel
  int sum_dh_last(int l, float* out, cudaStream_t s) {
    if (dec_pers()) return sum_partials(w.dh_rec[l], 1, nullptr, 0, ps, out, B * H, s);
    return sum_partials(w.dh_rec[l], kSplitB, l == L - 1 ? w.dq_h : nullptr, l == L - 1 ? kSplitB : 0, ps, out, B * H, s);
  }

  // decoder BPTT, reverse of decoder.py:74-107 for t = T_q-1 .. 0
  int dec_loop(cudaStream_t st, const int* lh_lo = nullptr, int lh_n = 0) {
    AttnShape as = attn_shape16(d, w);
    as.ldds16 = Sp;
    MMQG_CUDA(cudaMemsetAsync(w.ds_all, 0, sizeof(float) * (size_t)R * Sp, st));
    MMQG_CUDA(cudaMemsetAsync(w.ds16, 0, sizeof(b16) * (size_t)R * Sp, st));
    PdlScope pdl_scope(pdl_enabled());
    L2WindowScope l2_scope(w.m_txt16, attn_l2_window(d, w));
    int lh_next = 0;          // next loss-head group whose d h_top this loop has not waited for yet
    for (int t = d.T_q - 1; t >= 0; --t) {
      StepGemmScope step_scope;
      const bool last = t == d.T_q - 1;
      if (lh_next < lh_n && (lh_next == 0 || t < lh_lo[lh_next - 1])) {        // first step inside group lh_next
        MMQG_CUDA(cudaStreamWaitEvent(st, ev_lh(lh_next), 0));
        ++lh_next;
      }
      // t > 0: ONE product per layer, dG_l [W_hh | W_in] -> (d h_rec for step t-1 | d input of this step);
      // t == 0 writes the recurrent part where the text encoder's BPTT picks it up (dh_rec[l]).
      const bool fused = t > 0;
      const int Cd = H + C;
      const long long pc = (long long)B * 2 * H, pc0 = (long long)B * Cd;     // partial strides of dcat[l > 0], dcat[0]
      for (int l = L - 1; l >= 0; --l) {
        const float* acts = w.acts_dec[l] + (size_t)t * B * G;
        b16* dg = w.dg_dec[l] + (size_t)t * B * G;
        // recurrent gradient from step t+1 (always produced by a fused launch)
        const float* dh0 = nullptr; int ld0 = H, n0 = kSplitB; long long s0 = ps;
        if (!last) {
          dh0 = w.dcat[l];
          if (l > 0) { ld0 = 2 * H; s0 = pc; }
          else { ld0 = Cd; s0 = pc0; }
        }
        const float* dh1 = nullptr; int ld1 = H, n1 = 0; long long s1 = ps;
        const float* dh2 = nullptr;
        DropSpec dr;
        if (l == L - 1) {
          dh2 = w.dhtop + (size_t)t * B * H;
          if (!last) { dh1 = w.dq_h; n1 = kSplitB; }
        } else {
          n1 = kSplitB;               // d input of layer l+1, this step
          if (fused) { dh1 = w.dcat[l + 1] + H; ld1 = 2 * H; s1 = pc; }
          else dh1 = w.dx_above;
          if (g_drop_p > 0.f) {   // it is d/d(dropped h_l): back through the mask of layer l's output
            dr.seed = g_drop_seed; dr.ctr = g_drop_ctr; dr.sid = kSidDec + l; dr.base = (unsigned long long)t * B * H; dr.p = g_drop_p;
          }
        }
        MMQG_TRY(lstm_pointwise_bwd_bf16(acts, G, w.cs_dec[l] + (size_t)t * B * H, H, w.cs_dec[l] + (size_t)(t + 1) * B * H,
                                         H, dh0, ld0, n0, s0, dh1, ld1, n1, s1, dh2, H, w.dc[l], H, last ? 1 : 0, dg, G, B,
                                         H, st, dr));
        if (fused) {
          const int Nl = l > 0 ? 2 * H : Cd;
          MMQG_TRY(Tc(dg, G, false, w.wd_cat[l], Nl, true, B, Nl, G, w.dcat[l], Nl).split(kSplitB, l > 0 ? pc : pc0).run(st));
        } else {
          MMQG_TRY(Tc(dg, G, false, w.wd_cat[l], l > 0 ? 2 * H : Cd, true, B, H, G, w.dh_rec[l], H).split(kSplitB, ps).run(st));
          if (l > 0)
            MMQG_TRY(Tc(dg, G, false, w.wd_cat[l] + H, 2 * H, true, B, H, G, w.dx_above, H).split(kSplitB, ps).run(st));
          else
            MMQG_TRY(Tc(dg, G, false, w.wd_cat[0] + H, Cd, true, B, C, G, w.dctx_all + (size_t)t * B * C, C).run(st));
        }
      }
      float* ds = w.ds_all + (size_t)t * B * Sp;
      as.ds16 = w.ds16 + (size_t)t * B * Sp;
      // fused steps hand the context gradient over as split-K partials next to d h_rec; the attention
      // kernel sums them and leaves the sum in dctx_all for the hoisted memory gradient
      float* dctx = w.dctx_all + (size_t)t * B * C;
      if (fused) { as.dctx_parts = kSplitB; as.dctx_part_stride = pc0; as.dctx_sum = dctx; as.lddsum = C; }
      else { as.dctx_parts = 1; as.dctx_sum = nullptr; }
      MMQG_TRY(attn_bwd(w.attn_all + (size_t)t * B * Sp, ds, Sp, fused ? w.dcat[0] + H : dctx, fused ? Cd : C, w.m_txt, w.m_aud,
                        w.m_vid, nullptr, nullptr, as, st));
      MMQG_TRY(Tc(as.ds16, Sp, false, w.wa_h, H, true, B, H, Sp, w.dq_h, H).split(kSplitB, ps).run(st));
    }
    // memory gradients feed the encoders' BPTT
    return attn_dmem(w.attn_all, Sp, w.dctx_all, C, w.dm_txt, w.dm_vid, d.T_q, as, st);
  }

  // hoisted decoder gradients: LSTM weights/biases, attention Linears, decoder-side embedding rows
  int dec_hoisted(cudaStream_t st) {
    for (int l = 0; l < L; ++l) {
      const b16* dG = w.dg_dec[l];
      MMQG_TRY(Tc(dG, G, true, w.hs_dec[l], H, true, G, H, R, Gd.dec_w_hh[l], H).run(st));
      if (l > 0) {
        MMQG_TRY(Tc(dG, G, true, g_drop_p > 0.f ? w.hdrop_dec[l - 1] : w.hs_dec[l - 1] + (size_t)B * H, H, true, G, H, R,
                    Gd.dec_w_ih[l], H).run(st));
      } else {
        MMQG_TRY(Tc(dG, G, true, w.e_dec, Ep, true, G, E, R, Gd.dec_w_ih[0], X0).run(st));
        MMQG_TRY(Tc(dG, G, true, w.ctx16, C, true, G, C, R, Gd.dec_w_ih[0] + E, X0).run(st));
      }
      MMQG_TRY(colsum_bf16(dG, G, Gd.dec_b_ih[l], Gd.dec_b_hh[l], R, G, 0.f, st));
    }
    MMQG_TRY(Tc(w.ds16, Sp, true, w.e_dec, Ep, true, Sp, E, R, w.attn_dw_cat, Q).run(st));
    MMQG_TRY(Tc(w.ds16, Sp, true, w.hs_dec[L - 1], H, true, Sp, H, R, w.attn_dw_cat + E, Q).run(st));
    MMQG_TRY(colsum(w.ds_all, Sp, w.attn_db_cat, nullptr, R, Sp, 0.f, st));
    const int off[3] = {0, d.TM, d.TM + d.AM}, len[3] = {d.TM, d.AM, d.AM};
    for (int i = 0; i < 3; ++i) {
      MMQG_CUDA(cudaMemcpyAsync(Gd.attn_w[i], w.attn_dw_cat + (size_t)off[i] * Q, sizeof(float) * (size_t)len[i] * Q,
                                cudaMemcpyDeviceToDevice, st));
      MMQG_CUDA(cudaMemcpyAsync(Gd.attn_b[i], w.attn_db_cat + off[i], sizeof(float) * len[i], cudaMemcpyDeviceToDevice, st));
    }
    MMQG_TRY(Tc(w.dg_dec[0], G, false, w.wd_e, Ep, true, R, E, G, w.de_dec, E).second(w.ds16, Sp, w.wa_e, Ep, Sp).run(st));
    MMQG_CUDA(cudaMemsetAsync(Gd.emb, 0, sizeof(float) * (size_t)d.V * E, st));
    return embedding_scatter_add(Gd.emb, w.idx_dec, w.de_dec, R, E, d.V, st);
  }

  // video LSTM BPTT (encoder.py:69), dh_ext(t) = dM_vid(:,t,:), and its weight gradients
  int video(cudaStream_t st) {
    const int Hv = d.H_v, Gv = 4 * d.H_v;
    const long long pv = (long long)B * Hv;
    if (persist_video(d)) {
      tl_ktag = 900;
      MMQG_TRY(rec_bwd(w.acts_v, w.cs_v, w.dg_v, w.wvp_b, w.dm_vid, Hv, (long long)d.AM * Hv, nullptr, nullptr, w.flags_v,
                       d.T_v, B, Hv, st, len_video()));
    } else {
      for (int t = d.T_v - 1; t >= 0; --t) {
        StepGemmScope step_scope;
        const bool last = t == d.T_v - 1;
        b16* dg = w.dg_v + (size_t)t * B * Gv;
        MMQG_TRY(lstm_pointwise_bwd_bf16(w.acts_v + (size_t)t * B * Gv, Gv, t > 0 ? w.cs_v + (size_t)t * B * Hv : nullptr, Hv,
                                         w.cs_v + (size_t)(t + 1) * B * Hv, Hv, last ? nullptr : w.dh_rec_vid, Hv, kSplitB, pv,
                                         nullptr, 0, 0, 0, w.dm_vid + (size_t)t * Hv, d.AM * Hv, w.dc_v, Hv, last ? 1 : 0, dg,
                                         Gv, B, Hv, st));
        if (t > 0) MMQG_TRY(Tc(dg, Gv, false, w.wv_hh, Hv, true, B, Hv, Gv, w.dh_rec_vid, Hv).split(kSplitB, pv).run(st));
      }
    }
    if (d.T_v > 1)
      MMQG_TRY(Tc(w.dg_v + (size_t)B * Gv, Gv, true, w.hs_v + (size_t)B * Hv, Hv, true, Gv, Hv, (d.T_v - 1) * B, Gd.vid_w_hh, Hv).run(st));
    else
      MMQG_CUDA(cudaMemsetAsync(Gd.vid_w_hh, 0, sizeof(float) * (size_t)Gv * Hv, st));
    MMQG_TRY(Tc(w.dg_v, Gv, true, w.frames16, d.F_v, true, Gv, d.F_v, d.T_v * B, Gd.vid_w_ih, d.F_v).run(st));
    return colsum_bf16(w.dg_v, Gv, Gd.vid_b_ih, Gd.vid_b_hh, d.T_v * B, Gv, 0.f, st);
  }

  // BPTT of text layer l (encoder.py:95-100) and the input gradient the layer below needs
  int text_bptt(int l, cudaStream_t st) {
    const int I = l == 0 ? E : H, Ip = l == 0 ? Ep : H;
    if (g_drop_p > 0.f && l < L - 1)   // dx_text holds d/d(dropped h_l): back through the mask
      MMQG_TRY(dropout_scale_f32(w.dx_text, 1, 0, (long long)d.T_t * B * H, g_drop_seed, g_drop_ctr, kSidText + l, 0, g_drop_p, st));
    if (persist_text(d)) {
      // d loss / d h_final = the decoder's gradient w.r.t. its initial state (train.py:169) plus,
      // for the top layer, the step-0 attention query; d loss / d c_final sits in dc[l].
      MMQG_TRY(sum_dh_last(l, w.dh_last, st));
      const float* ext = l == L - 1 ? w.dm_txt : w.dx_text;
      const long long ts = l == L - 1 ? H : (long long)B * H, ld = l == L - 1 ? (long long)d.TM * H : H;
      MMQG_TRY(rec_bwd(w.acts_text[l], w.cs_text[l], w.dg_text[l], w.wtp_b[l], ext, ts, ld, w.dh_last, w.dc[l], w.flags,
                       d.T_t, B, H, st, len_text(0, l == L - 1)));
    } else {
      for (int t = d.T_t - 1; t >= 0; --t) {
        StepGemmScope step_scope;
        const bool last = t == d.T_t - 1;
        b16* dg = w.dg_text[l] + (size_t)t * B * G;
        const float* dh0 = last ? w.dh_rec[l] : w.dh_rec_enc;
        const float* dh1 = (last && l == L - 1) ? w.dq_h : nullptr;
        const float* dh2 = l == L - 1 ? w.dm_txt + (size_t)t * H : w.dx_text + (size_t)t * B * H;
        const int ldh2 = l == L - 1 ? d.TM * H : H;
        MMQG_TRY(lstm_pointwise_bwd_bf16(w.acts_text[l] + (size_t)t * B * G, G, t > 0 ? w.cs_text[l] + (size_t)t * B * H : nullptr,
                                         H, w.cs_text[l] + (size_t)(t + 1) * B * H, H, dh0, H, kSplitB, ps, dh1, H, kSplitB, ps,
                                         dh2, ldh2, w.dc[l], H, 0, dg, G, B, H, st));
        if (t > 0) MMQG_TRY(Tc(dg, G, false, w.wt_hh[l], H, true, B, H, G, w.dh_rec_enc, H).split(kSplitB, ps).run(st));
      }
    }
    return Tc(w.dg_text[l], G, false, w.wt_ih[l], Ip, true, d.T_t * B, I, G, w.dx_text, I).run(st);
  }

  int text_hoisted(int l, cudaStream_t st) {
    const int I = l == 0 ? E : H, Ip = l == 0 ? Ep : H;
    const b16* dG = w.dg_text[l];
    const b16* X = l == 0 ? w.x0 : (g_drop_p > 0.f ? w.xdrop_text[l - 1] : w.hs_text[l - 1] + (size_t)B * H);
    // K = T_t*B is long and the output small (64 tiles): split K so that no CTA lives longer than a
    // few microseconds -- these products share the SMs with the latency-bound BPTT kernels, which
    // can only start once 64 SMs are free -- and reduce the partial tiles afterwards.
    static const bool splitw_env = []() { const char* e = getenv("MMQG_SPLITW"); return !(e && e[0] == '0'); }();
    const bool splitw = splitw_env && (long long)d.T_t * B >= 16ll * 64 * kSplitW;      // >= 16 k-blocks per slice
    if (splitw) {
      MMQG_TRY(Tc(dG, G, true, X, Ip, true, G, I, d.T_t * B, w.dw_part, I).split(kSplitW, (long long)G * I).run(st));
      MMQG_TRY(reduce_partials(w.dw_part, kSplitW, (long long)G * I, Gd.text_w_ih[l], I, G, I, st));
      MMQG_TRY(Tc(dG + (size_t)B * G, G, true, w.hs_text[l] + (size_t)B * H, H, true, G, H, (d.T_t - 1) * B, w.dw_part, H)
                   .split(kSplitW, (long long)G * H).run(st));
      MMQG_TRY(reduce_partials(w.dw_part, kSplitW, (long long)G * H, Gd.text_w_hh[l], H, G, H, st));
      return colsum_bf16(dG, G, Gd.text_b_ih[l], Gd.text_b_hh[l], d.T_t * B, G, 0.f, st);
    }
    MMQG_TRY(Tc(dG, G, true, X, Ip, true, G, I, d.T_t * B, Gd.text_w_ih[l], I).run(st));
    if (d.T_t > 1)
      MMQG_TRY(Tc(dG + (size_t)B * G, G, true, w.hs_text[l] + (size_t)B * H, H, true, G, H, (d.T_t - 1) * B, Gd.text_w_hh[l], H).run(st));
    else
      MMQG_CUDA(cudaMemsetAsync(Gd.text_w_hh[l], 0, sizeof(float) * (size_t)G * H, st));
    return colsum_bf16(dG, G, Gd.text_b_ih[l], Gd.text_b_hh[l], d.T_t * B, G, 0.f, st);
  }

  // text layers pipelined over time chunks (mirror of the forward schedule): layer l runs chunk c
  // once layer l+1 has produced the input gradient of that chunk.  Layer l uses stream S(l); the
  // hoisted weight gradients of a finished layer go to `hoist`.
  int text_pipelined(int NC, cudaStream_t st, cudaStream_t hoist, cudaEvent_t const* ready = nullptr) {
    const int CL = (d.T_t + NC - 1) / NC;
    auto S = [&](int l) { return l == L - 1 ? st : g_aux.s[L - 2 - l]; };      // stage k of the reverse pipeline on s[k-1]
    const int n_mt = (B + 127) / 128;
    for (int l = 0; l < L; ++l) MMQG_CUDA(cudaMemsetAsync(w.flags_t[l], 0, sizeof(uint32_t) * (size_t)(d.T_t + 1) * n_mt, st));
    MMQG_CUDA(cudaEventRecord(g_aux.ev[1], st));
    for (int l = 0; l < L - 1; ++l) MMQG_CUDA(cudaStreamWaitEvent(S(l), g_aux.ev[1], 0));
    for (int c = NC - 1; c >= 0; --c) {
      const int t0 = c * CL, nT = (d.T_t - t0 < CL) ? d.T_t - t0 : CL;
      const bool tail = c == NC - 1;
      for (int l = L - 1; l >= 0; --l) {
        cudaStream_t s = S(l);
        const int I = l == 0 ? E : H, Ip = l == 0 ? Ep : H;
        DropSpec dr;
        if (l < L - 1) {
          MMQG_CUDA(cudaStreamWaitEvent(s, ev_bwd(l + 1, c), 0));
          if (g_drop_p > 0.f) {      // dx_text is d/d(dropped h_l): the kernel applies this layer's output mask to it
            dr.seed = g_drop_seed; dr.ctr = g_drop_ctr; dr.sid = kSidText + l; dr.base = (unsigned long long)t0 * B * H; dr.p = g_drop_p;
          }
        }
        if (tail)   // d loss / d h_final of this layer (decoder initial state, + step-0 attention query on top)
          MMQG_TRY(sum_dh_last(l, w.dh_last_l[l], s));
        const float* ext = l == L - 1 ? w.dm_txt + (size_t)t0 * H : w.dx_text + (size_t)t0 * B * H;
        const long long ts = l == L - 1 ? H : (long long)B * H, ld = l == L - 1 ? (long long)d.TM * H : H;
        tl_ktag = l * 16 + c;
        MMQG_TRY(lstm_seq_bwd_persist(w.acts_text[l] + (size_t)t0 * B * G, w.cs_text[l] + (size_t)t0 * B * H,
                                      w.dg_text[l] + (size_t)t0 * B * G, w.wtp_b[l], ext, ts, ld, tail ? w.dh_last_l[l] : nullptr,
                                      w.dc[l], w.flags_t[l] + (size_t)t0 * n_mt, nT, B, H, tail ? 0 : 1, w.dc[l], s, dr, false,
                                      len_text(t0, l == L - 1)));
        // input gradient of this chunk (what the layer below waits for) on the product stream
        float* dx = l == 0 ? w.dx_emb + (size_t)t0 * B * E : w.dx_text + (size_t)t0 * B * H;
        cudaStream_t gs = split_products() ? g_aux.g[l] : s;
        if (gs != s) {
          MMQG_CUDA(cudaEventRecord(ev_gb(l, c), s));
          MMQG_CUDA(cudaStreamWaitEvent(gs, ev_gb(l, c), 0));
        }
        MMQG_TRY(Tc(w.dg_text[l] + (size_t)t0 * B * G, G, false, w.wt_ih[l], Ip, true, nT * B, I, G, dx, I).run(gs));
        MMQG_CUDA(cudaEventRecord(ev_bwd(l, c), gs));
        if (c == 0) {
          MMQG_CUDA(cudaStreamWaitEvent(hoist, ev_bwd(l, 0), 0));
          MMQG_TRY(text_hoisted(l, hoist));
          if (ready && ready[3 + (L - 1 - l)]) MMQG_CUDA(cudaEventRecord(ready[3 + (L - 1 - l)], hoist));   // this layer's gradients are final
        }
      }
    }
    for (int l = 0; l < L; ++l) MMQG_CUDA(cudaStreamWaitEvent(st, ev_bwd(l, 0), 0));   // join the layer / product streams
    return 0;
  }

  int emb_enc(cudaStream_t st, bool chunked) {
    return embedding_scatter_add(Gd.emb, w.idx_ctx, chunked ? w.dx_emb : w.dx_text, d.T_t * B, E, d.V, st);
  }
};

static int train_backward_bf16_on(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                                  size_t workspace_bytes, mmqg_tensors& Gd, int phase, float dropout_p, unsigned long long seed,
                                  cudaStream_t st, cudaEvent_t const* ready);

int train_backward_bf16(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                        size_t workspace_bytes, mmqg_tensors& Gd, int phase, float dropout_p, unsigned long long seed,
                        cudaStream_t user, cudaEvent_t const* ready) {
  MMQG_TRY(check_dims_bf16(d));
  const size_t need = train_workspace_bytes_bf16(d, d.T_q);
  if (need > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, need);
  cudaStream_t st;
  MMQG_TRY(g_aux.enter(user, &st));
  const int rc = train_backward_bf16_on(d, P, bt, workspace, workspace_bytes, Gd, phase, dropout_p, seed, st, ready);
  if (rc != 0) { g_aux.join_after_error(user, st); return rc; }
  return g_aux.leave(user, st);
}

static int train_backward_bf16_on(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                                  size_t workspace_bytes, mmqg_tensors& Gd, int phase, float dropout_p, unsigned long long seed,
                                  cudaStream_t st, cudaEvent_t const* ready) {
  MMQG_TRY(check_dims_bf16(d));
  g_drop_p = dropout_p;
  g_drop_seed = seed;
  g_fwd_only = false;
  Ws16 w = carve16(d, d.T_q, workspace);
  if (w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  g_drop_ctr = w.seed_ctr;      // unchanged since the forward of this step: the same masks
  MMQG_TRY(set_len_state(d, bt, w));      // shift arrays were filled by the forward of this step
  Bwd16 b(d, P, bt, w, Gd);
  const bool lh_def = lh_deferred(d, w);
  if (phase == 1) {
    if (lh_def) MMQG_TRY(b.loss_head_bwd(0, d.T_q, true, st));      // phase-wise callers: serially, in front of the BPTT
    if (b.dec_pers()) MMQG_TRY(b.dec_loop_persist(st));
    else MMQG_TRY(b.dec_loop(st));
    return b.dec_hoisted(st);
  }
  if (phase == 2) return b.video(st);
  const int NC = text_chunks(d);
  if (phase == 3) {
    if (NC > 1) {
      MMQG_TRY(g_aux.init());
      cudaStream_t hoist = g_aux.s[AuxStream::NS - 1];
      MMQG_CUDA(cudaEventRecord(g_aux.ev[2], st));
      MMQG_CUDA(cudaStreamWaitEvent(hoist, g_aux.ev[2], 0));
      MMQG_TRY(b.text_pipelined(NC, st, hoist));
      MMQG_CUDA(cudaEventRecord(g_aux.ev[7], hoist));
      MMQG_CUDA(cudaStreamWaitEvent(st, g_aux.ev[7], 0));
      return b.emb_enc(st, true);
    }
    for (int l = d.L - 1; l >= 0; --l) {
      MMQG_TRY(b.text_bptt(l, st));
      MMQG_TRY(b.text_hoisted(l, st));
    }
    return b.emb_enc(st, false);
  }
  // phase 0: the whole backward with the hoisted products overlapped on an auxiliary stream
  MMQG_TRY(g_aux.init());
  cudaStream_t ax = g_aux.s[AuxStream::NS - 1];
  mark(6, st);
  if (b.dec_pers()) {
    // the persistent BPTT kernel holds 128 SMs: the loss-head backward (0.26 ms on the whole GPU) runs in front of it
    if (lh_def) MMQG_TRY(b.loss_head_bwd(0, d.T_q, true, st));
    if (ready && ready[0]) MMQG_CUDA(cudaEventRecord(ready[0], st));
    MMQG_TRY(b.dec_loop_persist(st));
  } else if (lh_def) {
    // loss-head backward on its own stream under the BPTT loop, one event per group
    cudaStream_t lh = g_aux.s[AuxStream::NS - 2];
    int lo[kMaxLhGroups];
    const int n = b.lh_groups(lo);
    MMQG_CUDA(cudaEventRecord(g_aux.ev[10], st));
    MMQG_CUDA(cudaStreamWaitEvent(lh, g_aux.ev[10], 0));
    for (int k = 0; k < n; ++k) {
      MMQG_TRY(b.loss_head_bwd(lo[k], k == 0 ? d.T_q : lo[k - 1], k == 0, lh));
      MMQG_CUDA(cudaEventRecord(ev_lh(k), lh));
    }
    if (ready && ready[0]) MMQG_CUDA(cudaEventRecord(ready[0], lh));          // loss-head group final
    MMQG_CUDA(cudaEventRecord(g_aux.ev[11], lh));
    MMQG_TRY(b.dec_loop(st, lo, n));
    MMQG_CUDA(cudaStreamWaitEvent(st, g_aux.ev[11], 0));                     // join the loss-head stream
  } else {
    if (ready && ready[0]) MMQG_CUDA(cudaEventRecord(ready[0], st));          // final since the forward call
    MMQG_TRY(b.dec_loop(st));
  }
  mark(7, st);
  MMQG_CUDA(cudaEventRecord(g_aux.ev[0], st));
  MMQG_CUDA(cudaStreamWaitEvent(ax, g_aux.ev[0], 0));
  MMQG_TRY(b.dec_hoisted(ax));
  if (ready && ready[1]) MMQG_CUDA(cudaEventRecord(ready[1], ax));     // decoder group final
  MMQG_TRY(b.video(ax));
  if (ready && ready[2]) MMQG_CUDA(cudaEventRecord(ready[2], ax));     // video group final
  if (NC > 1) {
    MMQG_TRY(b.text_pipelined(NC, st, ax, ready));
  } else {
    for (int l = d.L - 1; l >= 0; --l) {
      MMQG_TRY(b.text_bptt(l, st));
      MMQG_CUDA(cudaEventRecord(g_aux.ev[3 + l], st));
      MMQG_CUDA(cudaStreamWaitEvent(ax, g_aux.ev[3 + l], 0));
      MMQG_TRY(b.text_hoisted(l, ax));
      if (ready && ready[3 + (d.L - 1 - l)]) MMQG_CUDA(cudaEventRecord(ready[3 + (d.L - 1 - l)], ax));
    }
  }
  mark(8, st);
  MMQG_CUDA(cudaEventRecord(g_aux.ev[7], ax));
  MMQG_CUDA(cudaStreamWaitEvent(st, g_aux.ev[7], 0));       // join: everything is ordered on `st` again
  MMQG_TRY(b.emb_enc(st, NC > 1));
  mark(9, st);
  if (ready && ready[3 + d.L]) MMQG_CUDA(cudaEventRecord(ready[3 + d.L], st));     // shared embedding final (both scatter-adds landed)
  return 0;
}

}  // namespace mmqg

// Debug hooks for tools/sections.py: switch the section marks on/off; read the elapsed time (ms)
// from mark 0 to each mark after a synchronised eager step.
extern "C" void mmqg_debug_sections(int enable) { mmqg::g_sec.on = enable != 0; }
extern "C" int mmqg_debug_section_times(float* ms_out) {
  if (!mmqg::g_sec.created) return 1;
  for (int i = 0; i < 10; ++i) {
    ms_out[i] = 0.f;
    if (cudaEventElapsedTime(&ms_out[i], mmqg::g_sec.ev[0], mmqg::g_sec.ev[i]) != cudaSuccess) { cudaGetLastError(); ms_out[i] = -1.f; }
  }
  return 0;
}

// Whole-path orchestration: teacher-forced train step (forward, loss head, backward
// through time) and greedy decode, as sequences of kernel launches on the caller's stream.
//
// Restates the control flow of the reference's per-sample loops (train.py:153-177 and
// train.py:101-110) for a whole batch; layouts are time-major (T,B,*) so that every
// per-timestep operand is a contiguous slab and every hoisted whole-sequence product
// (input projections, weight gradients, input gradients) is ONE GEMM.
//
// Hoisting (SURVEY.md section 7 "Hard parts"):
//   encoder layers : X W_ih^T for all T_t steps in one GEMM; per step only h W_hh^T.
//   decoder        : the embedding columns of W_ih_l0 and of the three attention Linears
//                    are hoisted over all T_q teacher-forced steps; the context columns and
//                    the recurrent parts are per step (they depend on h_top(t-1)).
//   loss head      : h_top W_out^T + log-softmax + NLL in row chunks; with want_grads the
//                    chunk's dlogits feed dH, dW_out, db_out immediately, so full logits
//                    are never stored (train.py:174 + decoder.py:106 + their backward).
//   backward       : per step only the pointwise cell gradient and dG W_hh (+ dG W_ih for
//                    the layer below); all weight gradients are GEMMs over the stored dG.
#include <algorithm>
#include "kernels.h"

namespace mmqg {

static const int kSplitMax = 8; // split-K slices of the skinny (B x H) per-step backward products: 4 (SIMT), 8 (tensor-core parity mode)
static const int kPreSplitMax = 4;

struct Carver {
  char* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Ws {
  unsigned long long* seed_ctr;       // FIRST 8 bytes of the workspace: dropout call counter (same convention as the bf16 mode)
  float *xdrop_text[MMQG_MAX_LAYERS], *hdrop_dec[MMQG_MAX_LAYERS];   // dropped layer outputs = inputs of the next layer
  int64_t *idx_ctx, *idx_dec, *tgt_tm, *idx_cur;
  float *bsum_text[MMQG_MAX_LAYERS], *bsum_dec[MMQG_MAX_LAYERS], *bsum_vid;
  float *attn_w_cat, *attn_b_cat, *attn_dw_cat, *attn_db_cat;
  float *x0_text, *acts_text[MMQG_MAX_LAYERS], *hs_text[MMQG_MAX_LAYERS], *cs_text[MMQG_MAX_LAYERS];
  float *m_txt, *m_aud, *m_vid;
  float *acts_v, *hs_v, *cs_v;
  float *e_dec, *attn_all, *ds_all, *ctx_all, *acts_dec[MMQG_MAX_LAYERS], *hs_dec[MMQG_MAX_LAYERS], *cs_dec[MMQG_MAX_LAYERS];
  float *logits, *nll, *dhtop;
  float *dh_rec[MMQG_MAX_LAYERS], *dc[MMQG_MAX_LAYERS], *dx_above, *dq_h, *dctx_all, *dm_txt, *dm_vid, *de_dec;
  float *dh_rec_enc, *dh_rec_vid, *dx_text, *dc_v;
  float *xcat;
  int *shift_t, *shift_v;             // per-sample lengths (mmqg_batch.ctx_len / n_frames): first live step of each sample
  float *row_w, *frames_tm;           // loss-row weights (T_q*B); right-aligned time-major copy of the frames (T_v*B, F_v)
  void *x3_hs, *x3_dg;                // h_t / dG_t of the current recurrent step as bf16 [hi | lo], written by the cell kernels
  float* pre_part;                    // split-K partial pre-activations of one recurrent step (tensor-core parity mode)
  void *x3_arena, *x3_sa, *x3_sb;     // MMQG_MODE_FP32_TC: split copies (gemm_f32x3.cu)
  size_t x3_arena_bytes, x3_sa_bytes, x3_sb_bytes;
  int S_pad, Rc;
  size_t bytes;
};

static int vocab_chunk_rows32(int R, int V) {
  long long rc = (16ll << 20) / (V > 0 ? V : 1);
  rc = rc / 64 * 64;
  if (rc < 64) rc = 64;
  if (rc > R) rc = R;
  return (int)rc;
}

// T_q here is the number of decoder steps the buffers must hold (max_len for greedy).
static Ws carve(const mmqg_dims& d, int T_q, void* base, bool tc = false) {
  Ws w{};
  Carver c{reinterpret_cast<char*>(base), 0};
  const size_t B = d.B, H = d.H, G = 4 * (size_t)d.H, Gv = 4 * (size_t)d.H_v;
  const int S = d.TM + 2 * d.AM;
  w.S_pad = (S + 3) / 4 * 4;
  const size_t Q = d.E + d.H, C = (size_t)d.H + d.H_a + d.H_v, X0 = d.E + C;
  const size_t R = (size_t)T_q * B;
  w.Rc = vocab_chunk_rows32((int)R, d.V);
  w.seed_ctr = c.take<unsigned long long>(1);
  w.idx_ctx = c.take<int64_t>((size_t)d.T_t * B);
  w.idx_dec = c.take<int64_t>(R);
  w.tgt_tm = c.take<int64_t>(R);
  w.idx_cur = c.take<int64_t>(B);
  for (int l = 0; l < d.L; ++l) { w.bsum_text[l] = c.take<float>(G); w.bsum_dec[l] = c.take<float>(G); }
  w.bsum_vid = c.take<float>(Gv);
  w.attn_w_cat = c.take<float>(w.S_pad * Q);
  w.attn_b_cat = c.take<float>(w.S_pad);
  w.attn_dw_cat = c.take<float>(w.S_pad * Q);
  w.attn_db_cat = c.take<float>(w.S_pad);
  w.x0_text = c.take<float>((size_t)d.T_t * B * d.E);
  for (int l = 0; l < d.L; ++l) {
    w.acts_text[l] = c.take<float>((size_t)d.T_t * B * G);
    w.hs_text[l] = c.take<float>((size_t)(d.T_t + 1) * B * H);
    w.cs_text[l] = c.take<float>((size_t)(d.T_t + 1) * B * H);
  }
  w.m_txt = c.take<float>(B * d.TM * H);
  w.m_aud = c.take<float>(B * d.AM * d.H_a);
  w.m_vid = c.take<float>(B * d.AM * d.H_v);
  w.acts_v = c.take<float>((size_t)d.T_v * B * Gv);
  w.hs_v = c.take<float>((size_t)(d.T_v + 1) * B * d.H_v);
  w.cs_v = c.take<float>((size_t)(d.T_v + 1) * B * d.H_v);
  w.e_dec = c.take<float>(R * d.E);
  w.attn_all = c.take<float>(R * w.S_pad);
  w.ds_all = c.take<float>(R * w.S_pad);
  w.ctx_all = c.take<float>(R * C);
  for (int l = 0; l < d.L; ++l) {
    w.acts_dec[l] = c.take<float>(R * G);
    w.hs_dec[l] = c.take<float>((size_t)(T_q + 1) * B * H);
    w.cs_dec[l] = c.take<float>((size_t)(T_q + 1) * B * H);
  }
  w.logits = c.take<float>((size_t)w.Rc * d.V);
  w.nll = c.take<float>(R);
  w.dhtop = c.take<float>(R * H);
  for (int l = 0; l < d.L; ++l) { w.dh_rec[l] = c.take<float>(kSplitMax * B * H); w.dc[l] = c.take<float>(B * H); }
  w.dx_above = c.take<float>(kSplitMax * B * H);
  w.dq_h = c.take<float>(kSplitMax * B * H);
  w.dctx_all = c.take<float>(R * C);
  w.dm_txt = c.take<float>(B * d.TM * H);
  w.dm_vid = c.take<float>(B * d.AM * d.H_v);
  w.de_dec = c.take<float>(R * d.E);
  w.dh_rec_enc = c.take<float>(kSplitMax * B * H);
  w.dh_rec_vid = c.take<float>(kSplitMax * B * d.H_v);
  w.dx_text = c.take<float>((size_t)d.T_t * B * (d.E > d.H ? d.E : d.H));
  w.dc_v = c.take<float>(B * d.H_v);
  w.xcat = c.take<float>(B * X0);
  for (int l = 0; l + 1 < d.L; ++l) { w.xdrop_text[l] = c.take<float>((size_t)d.T_t * B * H); w.hdrop_dec[l] = c.take<float>(R * H); }
  w.shift_t = c.take<int>(B); w.shift_v = c.take<int>(B);
  w.row_w = c.take<float>(R);
  w.frames_tm = c.take<float>((size_t)d.T_v * B * d.F_v);
  if (tc) {
    // constant operands: every GEMM weight once per layout (forward (N,K) and backward (K,N) roles), column slices padded
    auto ent = [](size_t r, size_t cc) { return 4 * (r + 8) * (cc + 16) + 512; };
    size_t wb = ent(Gv, d.F_v) + ent(Gv, d.H_v) + ent(d.V, H) + ent(w.S_pad, Q) + ent(w.S_pad, 16);
    for (int l = 0; l < d.L; ++l)
      wb += ent(G, l == 0 ? d.E : H) + ent(G, H) + ent(G, l == 0 ? X0 + 16 : H) + ent(G, H);
    w.x3_arena_bytes = 2 * wb + (1u << 20);
    const long long NB = (long long)d.T_t * B > (long long)R ? (long long)d.T_t * B : (long long)R;
    const long long Km = std::max<long long>(std::max<long long>(G + w.S_pad, Gv), std::max<long long>(X0, d.F_v + d.H_v));
    size_t sa = std::max(f32x3_split_bytes(0, NB, Km), f32x3_split_bytes(1, Km, NB));
    sa = std::max(sa, std::max(f32x3_split_bytes(0, w.Rc, d.V), f32x3_split_bytes(1, d.V, w.Rc)));
    w.x3_sa_bytes = sa + 4096;
    size_t sb = std::max(f32x3_split_bytes(1, std::max<long long>(X0, H), NB), f32x3_split_bytes(1, d.F_v, B));
    w.x3_sb_bytes = sb + 4096;
    w.pre_part = c.take<float>((size_t)kPreSplitMax * B * std::max(G, Gv));
    w.x3_hs = c.take<char>((size_t)B * 2 * std::max(H, (size_t)d.H_v) * 2);
    w.x3_dg = c.take<char>((size_t)B * 2 * std::max(G, Gv) * 2);
    w.x3_arena = c.take<char>(w.x3_arena_bytes);
    w.x3_sa = c.take<char>(w.x3_sa_bytes);
    w.x3_sb = c.take<char>(w.x3_sb_bytes);
  }
  w.bytes = align_up(c.off, 256);
  return w;
}

static int check_dims(const mmqg_dims* d) {
  MMQG_REQUIRE(d, "null dims");
  MMQG_REQUIRE(d->B > 0 && d->T_t > 0 && d->T_v > 0 && d->T_q > 0 && d->V > 2, "dims: non-positive size");
  MMQG_REQUIRE(d->E > 0 && d->H > 0 && d->H_a > 0 && d->H_v > 0 && d->F_v > 0, "dims: non-positive width");
  MMQG_REQUIRE(d->L >= 1 && d->L <= MMQG_MAX_LAYERS, "dims: L=%d not in [1,%d]", d->L, MMQG_MAX_LAYERS);
  MMQG_REQUIRE(d->T_t <= d->TM && d->T_v <= d->AM, "dims: T_t<=TM and T_v<=AM required (%d,%d,%d,%d)", d->T_t, d->TM,
               d->T_v, d->AM);
  return 0;
}

static int check_tensors(const mmqg_dims& d, const mmqg_tensors* t, const char* what) {
  MMQG_REQUIRE(t, "%s: null", what);
  bool ok = t->emb && t->vid_w_ih && t->vid_w_hh && t->vid_b_ih && t->vid_b_hh && t->out_w && t->out_b;
  for (int i = 0; i < 3; ++i) ok = ok && t->attn_w[i] && t->attn_b[i];
  for (int l = 0; l < d.L; ++l)
    ok = ok && t->text_w_ih[l] && t->text_w_hh[l] && t->text_b_ih[l] && t->text_b_hh[l] && t->dec_w_ih[l] &&
         t->dec_w_hh[l] && t->dec_b_ih[l] && t->dec_b_hh[l];
  MMQG_REQUIRE(ok, "%s: null tensor pointer", what);
  return 0;
}

// GEMM call helper ------------------------------------------------------------------------
// MMQG_MODE_FP32_TC: the contractions of the current call run on the tensor cores (bf16 split, gemm_f32x3.cu)
static thread_local bool g32_tc = false;

struct GemmCall {
  mmqg_gemm_args a;
  bool b_const = false;
  const void* a_split = nullptr;
  GemmCall(const float* A, int lda, bool tA, const float* Bm, int ldb, bool tB, int M, int N, int K, float* C, int ldc) {
    a = mmqg_gemm_args{};
    a.A = A; a.lda = lda; a.transA = tA; a.B = Bm; a.ldb = ldb; a.transB = tB;
    a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.alpha = 1.f; a.beta = 0.f; a.split_k = 1;
  }
  GemmCall& second(const float* A2, int lda2, const float* B2, int ldb2, int K2) {
    a.A2 = A2; a.lda2 = lda2; a.B2 = B2; a.ldb2 = ldb2; a.K2 = K2; return *this;
  }
  GemmCall& add(const float* Cin, int ldcin, float beta = 1.f) { a.Cin = Cin; a.ldcin = ldcin; a.beta = beta; return *this; }
  GemmCall& accumulate(bool on) { if (on) { a.Cin = a.C; a.ldcin = a.ldc; a.beta = 1.f; } return *this; }
  GemmCall& bias(const float* b) { a.bias = b; return *this; }
  GemmCall& split(int s, long long stride) { a.split_k = s; a.c_split_stride = stride; return *this; }
  GemmCall& w() { b_const = true; return *this; }      // B (and B2) are parameters: constant during the call
  GemmCall& presplit(const void* p) { a_split = p; return *this; }   // tensor-core parity mode: A already split by its producer kernel
  int run(cudaStream_t st) {
    if (!g32_tc) return gemm_f32(a, st);
    if (a_split) f32x3_presplit_a(a_split);
    return gemm_f32x3(a, b_const, st);
  }
};

// Tensor-core parity mode: the (B x 4H) pre-activation product of one recurrent step is issued split-K so that ~128 SMs
// pull operands instead of ceil(B/128) * 4H/64; the partial sums go to `part` and the cell kernel adds them
// (to the hoisted input projection already in `gates`, or to `bias`).  Returns the PreSpec for lstm_pointwise_fwd.
static int step_pre_tc(GemmCall& g, float* part, int M, int N, const float* bias, bool add_gates, PreSpec* ps, cudaStream_t st) {
  const int base = ceil_div(M, 128) * ceil_div(N, 64);
  int S = 148 / base;
  S = S < 1 ? 1 : (S > kPreSplitMax ? kPreSplitMax : S);
  g.a.C = part; g.a.ldc = N; g.a.Cin = nullptr; g.a.beta = 0.f; g.a.bias = nullptr;
  g.split(S, (long long)M * N).w();
  MMQG_TRY(g.run(st));
  *ps = PreSpec{};
  ps->part = part; ps->n_part = S; ps->ld = N; ps->stride = (long long)M * N; ps->bias = bias; ps->add_gates = add_gates ? 1 : 0;
  return 0;
}

// inter-layer dropout of the current call (0 = off); stream ids as in engine_bf16.cu
static thread_local float g32_drop_p = 0.f;
static thread_local unsigned long long g32_drop_seed = 0;
static const int kSidText32 = 10, kSidDec32 = 20;

// per-sample lengths of the current call (mmqg_batch.ctx_len / tgt_len / n_frames, NULL = uniform): sequences are right-aligned
// in time inside the fixed (T, B) layout, exactly as in the bf16 mode
static thread_local bool g32_len = false;
static LenSpec len32(const int* shift, int t, int mem_shift) {
  LenSpec l;
  if (g32_len) { l.shift = shift; l.t_base = t; l.mem_shift = mem_shift; }
  return l;
}

static AttnShape attn_shape(const mmqg_dims& d) { return AttnShape{d.B, d.TM, d.AM, d.H, d.H_a, d.H_v, d.T_t, d.T_v}; }

// ----------------------------------------------------------------------------------------
// Encoder forward: video LSTM + padding (reference encoder.py:69, train.py:155-157) and the
// text LSTM stack (encoder.py:95-100, train.py:159-166).
static int encoder_forward(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, Ws& w, cudaStream_t st) {
  const int B = d.B, H = d.H, G = 4 * d.H, Hv = d.H_v, Gv = 4 * d.H_v;
  // Memories are (B,TM|AM,*) like the reference's padded tensors (train.py:155-160), but the
  // padding rows are never read (the context sums stop at T_t / T_v), so they are not zeroed.
  if (g32_len) {      // rows beyond a sample's own length ARE read (the context sums run over T_t / T_v rows): zero padding of train.py:155-160
    MMQG_TRY(audio_pad(bt.audio, w.m_aud, bt.n_frames, B, d.T_v, d.AM, d.H_a, st));
    MMQG_CUDA(cudaMemsetAsync(w.m_vid, 0, sizeof(float) * (size_t)B * d.AM * Hv, st));
    MMQG_CUDA(cudaMemsetAsync(w.m_txt, 0, sizeof(float) * (size_t)B * d.TM * H, st));
    MMQG_TRY(frames_to_time_major_f32(bt.frames, w.frames_tm, w.shift_v, B, d.T_v, d.F_v, st));
  } else
  MMQG_CUDA(cudaMemcpy2DAsync(w.m_aud, sizeof(float) * (size_t)d.AM * d.H_a, bt.audio,
                              sizeof(float) * (size_t)d.T_v * d.H_a, sizeof(float) * (size_t)d.T_v * d.H_a, B,
                              cudaMemcpyDeviceToDevice, st));
  // video LSTM
  MMQG_TRY(add2(P.vid_b_ih, P.vid_b_hh, w.bsum_vid, Gv, st));
  for (int t = 0; t < d.T_v; ++t) {
    StepGemmScope step_scope;
    PdlScope pdl_scope(g32_tc && pdl_enabled());
    float* acts = w.acts_v + (size_t)t * B * Gv;
    GemmCall g(g32_len ? w.frames_tm + (size_t)t * B * d.F_v : bt.frames + (size_t)t * d.F_v, g32_len ? d.F_v : d.T_v * d.F_v, false,
               P.vid_w_ih, d.F_v, true, B, Gv, d.F_v, acts, Gv);
    g.bias(w.bsum_vid);
    if (t > 0) g.second(w.hs_v + (size_t)t * B * Hv, Hv, P.vid_w_hh, Hv, Hv);
    PreSpec ps;
    if (g32_tc) MMQG_TRY(step_pre_tc(g, w.pre_part, B, Gv, w.bsum_vid, false, &ps, st));
    else MMQG_TRY(g.w().run(st));
    MMQG_TRY(lstm_pointwise_fwd(acts, Gv, t > 0 ? w.cs_v + (size_t)t * B * Hv : nullptr, Hv,
                                w.cs_v + (size_t)(t + 1) * B * Hv, Hv, w.hs_v + (size_t)(t + 1) * B * Hv, Hv,
                                g32_len ? w.m_vid : w.m_vid + (size_t)t * Hv, d.AM * Hv, B, Hv, st, ps, len32(w.shift_v, t, 1)));
  }
  // text LSTM stack
  MMQG_TRY(embedding_gather(P.emb, w.idx_ctx, w.x0_text, d.E, d.T_t * B, d.E, d.V, st));
  for (int l = 0; l < d.L; ++l) {
    MMQG_TRY(add2(P.text_b_ih[l], P.text_b_hh[l], w.bsum_text[l], G, st));
    if (l > 0 && g32_drop_p > 0.f)      // encoder.py:91: dropout between the LSTM layers (train mode)
      MMQG_TRY(dropout_f32(w.hs_text[l - 1] + (size_t)B * H, w.xdrop_text[l - 1], (long long)d.T_t * B * H, g32_drop_seed, w.seed_ctr,
                           kSidText32 + l - 1, 0, g32_drop_p, st));
    const float* X = l == 0 ? w.x0_text : (g32_drop_p > 0.f ? w.xdrop_text[l - 1] : w.hs_text[l - 1] + (size_t)B * H);
    const int I = l == 0 ? d.E : H;
    MMQG_TRY(GemmCall(X, I, false, P.text_w_ih[l], I, true, d.T_t * B, G, I, w.acts_text[l], G)
                 .bias(w.bsum_text[l]).w().run(st));
    for (int t = 0; t < d.T_t; ++t) {
      StepGemmScope step_scope;
      PdlScope pdl_scope(g32_tc && pdl_enabled());
      float* acts = w.acts_text[l] + (size_t)t * B * G;
      PreSpec ps;
      if (t > 0) {
        GemmCall g(w.hs_text[l] + (size_t)t * B * H, H, false, P.text_w_hh[l], H, true, B, G, H, acts, G);
        if (g32_tc && H % 8 == 0) g.presplit(w.x3_hs);      // h_{t-1} as [hi | lo], written by the cell kernel of step t-1
        if (g32_tc) MMQG_TRY(step_pre_tc(g, w.pre_part, B, G, nullptr, true, &ps, st));
        else MMQG_TRY(g.accumulate(true).w().run(st));
      }
      if (g32_tc && H % 8 == 0 && t + 1 < d.T_t) ps.h_split = w.x3_hs;
      MMQG_TRY(lstm_pointwise_fwd(acts, G, t > 0 ? w.cs_text[l] + (size_t)t * B * H : nullptr, H,
                                  w.cs_text[l] + (size_t)(t + 1) * B * H, H, w.hs_text[l] + (size_t)(t + 1) * B * H, H,
                                  l == d.L - 1 ? (g32_len ? w.m_txt : w.m_txt + (size_t)t * H) : nullptr, d.TM * H, B, H, st, ps,
                                  len32(w.shift_t, t, 1)));
    }
  }
  return 0;
}

static int pack_attention(const mmqg_dims& d, const mmqg_tensors& P, Ws& w, cudaStream_t st) {
  const size_t Q = d.E + d.H;
  MMQG_CUDA(cudaMemsetAsync(w.attn_w_cat, 0, sizeof(float) * w.S_pad * Q, st));
  MMQG_CUDA(cudaMemsetAsync(w.attn_b_cat, 0, sizeof(float) * w.S_pad, st));
  const int off[3] = {0, d.TM, d.TM + d.AM}, len[3] = {d.TM, d.AM, d.AM};
  for (int i = 0; i < 3; ++i) {
    MMQG_CUDA(cudaMemcpyAsync(w.attn_w_cat + (size_t)off[i] * Q, P.attn_w[i], sizeof(float) * len[i] * Q,
                              cudaMemcpyDeviceToDevice, st));
    MMQG_CUDA(cudaMemcpyAsync(w.attn_b_cat + off[i], P.attn_b[i], sizeof(float) * len[i], cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// decoder state slab 0 := encoder final state (train.py:169)
static int handoff_state(const mmqg_dims& d, Ws& w, cudaStream_t st) {
  const size_t n = (size_t)d.B * d.H;
  for (int l = 0; l < d.L; ++l) {
    MMQG_CUDA(cudaMemcpyAsync(w.hs_dec[l], w.hs_text[l] + (size_t)d.T_t * n, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    MMQG_CUDA(cudaMemcpyAsync(w.cs_dec[l], w.cs_text[l] + (size_t)d.T_t * n, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

}  // namespace mmqg

using namespace mmqg;

extern "C" {

size_t mmqg_train_workspace_bytes(const mmqg_dims* d, int mode) {
  if (check_dims(d) != 0) return 0;
  if (mode == MMQG_MODE_BF16) return check_dims_bf16(*d) == 0 ? train_workspace_bytes_bf16(*d, d->T_q) : 0;
  return carve(*d, d->T_q, nullptr, mode == MMQG_MODE_FP32_TC).bytes;
}

size_t mmqg_greedy_workspace_bytes(const mmqg_dims* d, int max_len, int mode) {
  if (check_dims(d) != 0 || max_len <= 0) return 0;
  if (mode == MMQG_MODE_BF16) return check_dims_bf16(*d) == 0 ? greedy_workspace_bytes_bf16(*d, max_len) : 0;
  return carve(*d, max_len, nullptr, mode == MMQG_MODE_FP32_TC).bytes;
}

int mmqg_train_forward(const mmqg_dims* dp, const mmqg_tensors* params, const mmqg_batch* batch, void* workspace,
                       size_t workspace_bytes, float* loss_out, int want_grads, mmqg_tensors* grads, float grad_scale,
                       float dropout_p, unsigned long long seed, int mode, void* stream) {
  MMQG_TRY(check_dims(dp));
  const mmqg_dims& d = *dp;
  MMQG_TRY(check_tensors(d, params, "params"));
  MMQG_REQUIRE(batch && batch->context && batch->target && batch->frames && batch->audio, "batch: null pointer");
  MMQG_REQUIRE(workspace && loss_out, "null workspace / loss_out");
  MMQG_REQUIRE(mode == MMQG_MODE_FP32 || mode == MMQG_MODE_BF16 || mode == MMQG_MODE_FP32_TC, "unknown mode %d", mode);
  MMQG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p=%g not in [0,1)", dropout_p);
  if (want_grads) MMQG_TRY(check_tensors(d, grads, "grads"));
  if (mode == MMQG_MODE_BF16)
    return train_forward_bf16(d, *params, *batch, workspace, workspace_bytes, loss_out, want_grads, grads, grad_scale,
                              dropout_p, seed, as_stream(stream));
  g32_drop_p = dropout_p;
  g32_drop_seed = seed;
  g32_len = batch->ctx_len || batch->tgt_len || batch->n_frames;
  g32_tc = mode == MMQG_MODE_FP32_TC;
  Ws w = carve(d, d.T_q, workspace, g32_tc);
  if (g32_tc) f32x3_bind(w.x3_arena, w.x3_arena_bytes, w.x3_sa, w.x3_sa_bytes, w.x3_sb, w.x3_sb_bytes);
  if (w.bytes > workspace_bytes)
    return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  cudaStream_t st = as_stream(stream);
  const mmqg_tensors& P = *params;
  const int B = d.B, H = d.H, G = 4 * d.H, E = d.E, Q = d.E + d.H, C = d.H + d.H_a + d.H_v, X0 = E + C;
  const int R = d.T_q * B, Sp = w.S_pad;

  if (dropout_p > 0.f) MMQG_TRY(bump_counter(w.seed_ctr, st));     // this call's masks: seed + (calls so far)
  if (g32_len) MMQG_TRY(prep_lengths(batch->ctx_len, batch->tgt_len, batch->n_frames, w.shift_t, w.shift_v, w.row_w, B, d.T_t, d.T_v, d.T_q, st));
  MMQG_TRY(build_indices(batch->context, batch->target, w.idx_ctx, w.idx_dec, w.tgt_tm, B, d.T_t, d.T_q, st, g32_len ? w.shift_t : nullptr));
  MMQG_TRY(encoder_forward(d, P, *batch, w, st));
  MMQG_TRY(handoff_state(d, w, st));
  MMQG_TRY(pack_attention(d, P, w, st));

  // decoder: hoisted embedding-column products over all teacher-forced steps
  MMQG_TRY(embedding_gather(P.emb, w.idx_dec, w.e_dec, E, R, E, d.V, st));
  for (int l = 0; l < d.L; ++l) MMQG_TRY(add2(P.dec_b_ih[l], P.dec_b_hh[l], w.bsum_dec[l], G, st));
  MMQG_TRY(GemmCall(w.e_dec, E, false, P.dec_w_ih[0], X0, true, R, G, E, w.acts_dec[0], G).bias(w.bsum_dec[0]).w().run(st));
  MMQG_TRY(GemmCall(w.e_dec, E, false, w.attn_w_cat, Q, true, R, Sp, E, w.attn_all, Sp).bias(w.attn_b_cat).w().run(st));
  const AttnShape as = attn_shape(d);
  for (int t = 0; t < d.T_q; ++t) {
    StepGemmScope step_scope;
    PdlScope pdl_scope(g32_tc && pdl_enabled());
    const float* htop_prev = w.hs_dec[d.L - 1] + (size_t)t * B * H;
    float* sc = w.attn_all + (size_t)t * B * Sp;
    float* ctx = w.ctx_all + (size_t)t * B * C;
    MMQG_TRY(GemmCall(htop_prev, H, false, w.attn_w_cat + E, Q, true, B, Sp, H, sc, Sp).accumulate(true).w().run(st));
    MMQG_TRY(attn_fwd(sc, Sp, w.m_txt, w.m_aud, w.m_vid, ctx, C, as, st));
    for (int l = 0; l < d.L; ++l) {
      float* acts = w.acts_dec[l] + (size_t)t * B * G;
      const float* hprev = w.hs_dec[l] + (size_t)t * B * H;
      PreSpec ps;
      if (l == 0) {
        GemmCall g(ctx, C, false, P.dec_w_ih[0] + E, X0, true, B, G, C, acts, G);
        g.second(hprev, H, P.dec_w_hh[0], H, H);
        if (g32_tc) MMQG_TRY(step_pre_tc(g, w.pre_part, B, G, nullptr, true, &ps, st));
        else MMQG_TRY(g.accumulate(true).w().run(st));
      } else {
        const float* xin = g32_drop_p > 0.f ? w.hdrop_dec[l - 1] + (size_t)t * B * H : w.hs_dec[l - 1] + (size_t)(t + 1) * B * H;
        GemmCall g(xin, H, false, P.dec_w_ih[l], H, true, B, G, H, acts, G);
        g.second(hprev, H, P.dec_w_hh[l], H, H);
        if (g32_tc) MMQG_TRY(step_pre_tc(g, w.pre_part, B, G, w.bsum_dec[l], false, &ps, st));
        else MMQG_TRY(g.bias(w.bsum_dec[l]).w().run(st));
      }
      MMQG_TRY(lstm_pointwise_fwd(acts, G, w.cs_dec[l] + (size_t)t * B * H, H, w.cs_dec[l] + (size_t)(t + 1) * B * H, H,
                                  w.hs_dec[l] + (size_t)(t + 1) * B * H, H, nullptr, 0, B, H, st, ps));
      if (g32_drop_p > 0.f && l + 1 < d.L)      // decoder.py:69: dropout between the LSTM layers
        MMQG_TRY(dropout_f32(w.hs_dec[l] + (size_t)(t + 1) * B * H, w.hdrop_dec[l] + (size_t)t * B * H, (long long)B * H, g32_drop_seed,
                             w.seed_ctr, kSidDec32 + l, (unsigned long long)t * B * H, g32_drop_p, st));
    }
  }
  // loss head in row chunks (rows r = t*B + b of h_top)
  const float* htop = w.hs_dec[d.L - 1] + (size_t)B * H;
  const float dscale = want_grads ? grad_scale / (float)B : 0.f;
  for (int r0 = 0, first = 1; r0 < R; r0 += w.Rc, first = 0) {
    const int rc = R - r0 < w.Rc ? R - r0 : w.Rc;
    MMQG_TRY(GemmCall(htop + (size_t)r0 * H, H, false, P.out_w, H, true, rc, d.V, H, w.logits, d.V).bias(P.out_b).w().run(st));
    MMQG_TRY(nll_rows(w.logits, d.V, w.tgt_tm + r0, 1, w.nll + r0, rc, d.V, dscale, st, g32_len ? w.row_w + r0 : nullptr));
    if (want_grads) {
      MMQG_TRY(GemmCall(w.logits, d.V, false, P.out_w, H, false, rc, H, d.V, w.dhtop + (size_t)r0 * H, H).w().run(st));
      MMQG_TRY(GemmCall(w.logits, d.V, true, htop + (size_t)r0 * H, H, false, d.V, H, rc, grads->out_w, H)
                   .accumulate(!first).run(st));
      MMQG_TRY(colsum(w.logits, d.V, grads->out_b, nullptr, rc, d.V, first ? 0.f : 1.f, st));
    }
  }
  MMQG_TRY(sum_scale(w.nll, R, 1.0f / (float)B, loss_out, st));
  return 0;
}

int mmqg_train_backward_events(const mmqg_dims* dp, const mmqg_tensors* params, const mmqg_batch* batch, void* workspace,
                               size_t workspace_bytes, mmqg_tensors* grads, void* const* ready_events, float dropout_p,
                               unsigned long long seed, int mode, void* stream) {
  MMQG_REQUIRE(ready_events, "ready_events: null pointer (use mmqg_train_backward phase 0)");
  MMQG_REQUIRE(dp && dp->L >= 1 && dp->L <= MMQG_MAX_LAYERS, "dims: L out of range");
  const int n_ev = 4 + dp->L;         // loss head, decoder, video, text layers L-1 .. 0, shared embedding
  cudaEvent_t ev[4 + MMQG_MAX_LAYERS];
  for (int i = 0; i < n_ev; ++i) ev[i] = reinterpret_cast<cudaEvent_t>(ready_events[i]);
  if (mode == MMQG_MODE_BF16) {
    MMQG_TRY(check_dims(dp));
    MMQG_TRY(check_tensors(*dp, params, "params"));
    MMQG_TRY(check_tensors(*dp, grads, "grads"));
    MMQG_REQUIRE(batch && batch->frames, "batch: null pointer");
    MMQG_REQUIRE(workspace, "null workspace");
    MMQG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p=%g not in [0,1)", dropout_p);
    return train_backward_bf16(*dp, *params, *batch, workspace, workspace_bytes, *grads, 0, dropout_p, seed, as_stream(stream), ev);
  }
  if (ev[0]) MMQG_CUDA(cudaEventRecord(ev[0], as_stream(stream)));      // fp32 mode: the loss-head gradients are final since the forward call
  for (int ph = 1; ph <= 3; ++ph) {     // fp32 parity mode: phases in order on the caller's stream
    MMQG_TRY(mmqg_train_backward(dp, params, batch, workspace, workspace_bytes, grads, ph, dropout_p, seed, mode, stream));
    if (ph < 3) {
      if (ev[ph]) MMQG_CUDA(cudaEventRecord(ev[ph], as_stream(stream)));
    } else {
      for (int i = 3; i < n_ev; ++i)
        if (ev[i]) MMQG_CUDA(cudaEventRecord(ev[i], as_stream(stream)));
    }
  }
  return 0;
}

int mmqg_train_backward(const mmqg_dims* dp, const mmqg_tensors* params, const mmqg_batch* batch, void* workspace,
                        size_t workspace_bytes, mmqg_tensors* grads, int phase, float dropout_p,
                        unsigned long long seed, int mode, void* stream) {
  MMQG_TRY(check_dims(dp));
  const mmqg_dims& d = *dp;
  MMQG_TRY(check_tensors(d, params, "params"));
  MMQG_TRY(check_tensors(d, grads, "grads"));
  MMQG_REQUIRE(batch && batch->frames, "batch: null pointer");
  MMQG_REQUIRE(workspace, "null workspace");
  MMQG_REQUIRE(mode == MMQG_MODE_FP32 || mode == MMQG_MODE_BF16 || mode == MMQG_MODE_FP32_TC, "unknown mode %d", mode);
  MMQG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p=%g not in [0,1)", dropout_p);
  MMQG_REQUIRE(phase >= 0 && phase <= 3, "phase %d not in 0..3", phase);
  if (mode == MMQG_MODE_BF16)
    return train_backward_bf16(d, *params, *batch, workspace, workspace_bytes, *grads, phase, dropout_p, seed, as_stream(stream));
  g32_drop_p = dropout_p;
  g32_drop_seed = seed;
  g32_len = batch->ctx_len || batch->tgt_len || batch->n_frames;      // shift_t / shift_v / row_w were prepared by the forward call
  g32_tc = mode == MMQG_MODE_FP32_TC;
  Ws w = carve(d, d.T_q, workspace, g32_tc);
  if (g32_tc) f32x3_bind(w.x3_arena, w.x3_arena_bytes, w.x3_sa, w.x3_sa_bytes, w.x3_sb, w.x3_sb_bytes);
  if (w.bytes > workspace_bytes)
    return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  cudaStream_t st = as_stream(stream);
  const mmqg_tensors& P = *params;
  mmqg_tensors& Gd = *grads;
  const int B = d.B, H = d.H, G = 4 * d.H, E = d.E, Q = d.E + d.H, C = d.H + d.H_a + d.H_v, X0 = E + C;
  const int R = d.T_q * B, Sp = w.S_pad, L = d.L;
  const int kSplit = g32_tc ? kSplitMax : 4;
  const bool x3dg = g32_tc && d.H % 2 == 0;      // the cell-gradient kernels also write dG as bf16 [hi | lo] (4H % 8 == 0)
  const long long ps = (long long)B * H;       // split-K partial stride for (B,H) products
  const AttnShape as = attn_shape(d);

  if (phase == 0) {      // whole backward; the fp32 mode simply runs the phases in order
    for (int ph = 1; ph <= 3; ++ph)
      MMQG_TRY(mmqg_train_backward(dp, params, batch, workspace, workspace_bytes, grads, ph, dropout_p, seed, mode, stream));
    return 0;
  }
  if (phase == 1) {
    // ---- decoder BPTT (reverse of decoder.py:74-107 for t = T_q-1 .. 0) ----
    // pad columns of dS (slots S..S_pad) take part in the K-loops below against zero weight
    // rows, so they must be finite: clear the buffer once.
    if (Sp != d.TM + 2 * d.AM) MMQG_CUDA(cudaMemsetAsync(w.ds_all, 0, sizeof(float) * (size_t)R * Sp, st));
    for (int t = d.T_q - 1; t >= 0; --t) {
      StepGemmScope step_scope;
      PdlScope pdl_scope(g32_tc && pdl_enabled());
      const bool last = t == d.T_q - 1;
      for (int l = L - 1; l >= 0; --l) {
        float* acts = w.acts_dec[l] + (size_t)t * B * G;
        const float* dh0 = last ? nullptr : w.dh_rec[l];
        const float* dh1 = nullptr; int n1 = 0;
        const float* dh2 = nullptr;
        if (l == L - 1) {
          dh2 = w.dhtop + (size_t)t * B * H;
          if (!last) { dh1 = w.dq_h; n1 = kSplit; }
        } else {
          dh1 = w.dx_above; n1 = kSplit;
          if (g32_drop_p > 0.f)      // dx_above is d/d(dropped h_l(t)): back through the mask of layer l's output
            MMQG_TRY(dropout_scale_f32(w.dx_above, kSplit, ps, (long long)B * H, g32_drop_seed, w.seed_ctr, kSidDec32 + l,
                                       (unsigned long long)t * B * H, g32_drop_p, st));
        }
        MMQG_TRY(lstm_pointwise_bwd(acts, G, w.cs_dec[l] + (size_t)t * B * H, H, w.cs_dec[l] + (size_t)(t + 1) * B * H, H,
                                    dh0, H, kSplit, ps, dh1, H, n1, ps, dh2, H, w.dc[l], H, last ? 1 : 0, B, H, st, x3dg ? w.x3_dg : nullptr));
        const void* dgs = x3dg ? w.x3_dg : nullptr;       // dG_l(t) as [hi | lo], shared by the products below
        MMQG_TRY(GemmCall(acts, G, false, P.dec_w_hh[l], H, false, B, H, G, w.dh_rec[l], H).split(kSplit, ps).w().presplit(dgs).run(st));
        if (l > 0)
          MMQG_TRY(GemmCall(acts, G, false, P.dec_w_ih[l], H, false, B, H, G, w.dx_above, H).split(kSplit, ps).w().presplit(dgs).run(st));
        else
          MMQG_TRY(GemmCall(acts, G, false, P.dec_w_ih[0] + E, X0, false, B, C, G, w.dctx_all + (size_t)t * B * C, C).w().presplit(dgs).run(st));
      }
      // dS(t) goes to its own buffer: the softmax weights are needed again by the hoisted
      // memory gradients (attn_dmem) after the loop.
      float* ds = w.ds_all + (size_t)t * B * Sp;
      MMQG_TRY(attn_bwd(w.attn_all + (size_t)t * B * Sp, ds, Sp, w.dctx_all + (size_t)t * B * C, C, w.m_txt, w.m_aud,
                        w.m_vid, nullptr, nullptr, as, st));
      MMQG_TRY(GemmCall(ds, Sp, false, w.attn_w_cat + E, Q, false, B, H, Sp, w.dq_h, H).split(kSplit, ps).w().run(st));
    }
    // ---- hoisted decoder weight gradients over all steps ----
    for (int l = 0; l < L; ++l) {
      const float* dG = w.acts_dec[l];
      MMQG_TRY(GemmCall(dG, G, true, w.hs_dec[l], H, false, G, H, R, Gd.dec_w_hh[l], H).run(st));
      if (l > 0) {
        MMQG_TRY(GemmCall(dG, G, true, g32_drop_p > 0.f ? w.hdrop_dec[l - 1] : w.hs_dec[l - 1] + (size_t)B * H, H, false, G, H, R, Gd.dec_w_ih[l], H).run(st));
      } else {
        MMQG_TRY(GemmCall(dG, G, true, w.e_dec, E, false, G, E, R, Gd.dec_w_ih[0], X0).run(st));
        MMQG_TRY(GemmCall(dG, G, true, w.ctx_all, C, false, G, C, R, Gd.dec_w_ih[0] + E, X0).run(st));
      }
      MMQG_TRY(colsum(dG, G, Gd.dec_b_ih[l], Gd.dec_b_hh[l], R, G, 0.f, st));
    }
    // memory gradients: dM(b,j,:) = sum_t a_t(b,j) dctx_t(b,:)  (rows j < T_t / T_v)
    MMQG_TRY(attn_dmem(w.attn_all, Sp, w.dctx_all, C, w.dm_txt, w.dm_vid, d.T_q, as, st));
    // attention Linears: dW = dS^T [E_dec | h_top(t-1)], db = colsum(dS)
    MMQG_TRY(GemmCall(w.ds_all, Sp, true, w.e_dec, E, false, Sp, E, R, w.attn_dw_cat, Q).run(st));
    MMQG_TRY(GemmCall(w.ds_all, Sp, true, w.hs_dec[L - 1], H, false, Sp, H, R, w.attn_dw_cat + E, Q).run(st));
    MMQG_TRY(colsum(w.ds_all, Sp, w.attn_db_cat, nullptr, R, Sp, 0.f, st));
    const int off[3] = {0, d.TM, d.TM + d.AM}, len[3] = {d.TM, d.AM, d.AM};
    for (int i = 0; i < 3; ++i) {
      MMQG_CUDA(cudaMemcpyAsync(Gd.attn_w[i], w.attn_dw_cat + (size_t)off[i] * Q, sizeof(float) * (size_t)len[i] * Q,
                                cudaMemcpyDeviceToDevice, st));
      MMQG_CUDA(cudaMemcpyAsync(Gd.attn_b[i], w.attn_db_cat + off[i], sizeof(float) * len[i], cudaMemcpyDeviceToDevice, st));
    }
    // embedding gradient, decoder side: dE = dG0 W_ih_l0[:, :E] + dS W_attn[:, :E]
    MMQG_TRY(GemmCall(w.acts_dec[0], G, false, P.dec_w_ih[0], X0, false, R, E, G, w.de_dec, E)
                 .second(w.ds_all, Sp, w.attn_w_cat, Q, Sp).w().run(st));
    MMQG_CUDA(cudaMemsetAsync(Gd.emb, 0, sizeof(float) * (size_t)d.V * E, st));
    MMQG_TRY(embedding_scatter_add(Gd.emb, w.idx_dec, w.de_dec, R, E, d.V, st));
    return 0;
  }

  if (phase == 2) {
    // ---- video LSTM BPTT (encoder.py:69); dh_ext(t) = dM_vid(:,t,:) ----
    const int Hv = d.H_v, Gv = 4 * d.H_v;
    const long long pv = (long long)B * Hv;
    for (int t = d.T_v - 1; t >= 0; --t) {
      StepGemmScope step_scope;
      PdlScope pdl_scope(g32_tc && pdl_enabled());
      const bool last = t == d.T_v - 1;
      float* acts = w.acts_v + (size_t)t * B * Gv;
      MMQG_TRY(lstm_pointwise_bwd(acts, Gv, t > 0 ? w.cs_v + (size_t)t * B * Hv : nullptr, Hv,
                                  w.cs_v + (size_t)(t + 1) * B * Hv, Hv, last ? nullptr : w.dh_rec_vid, Hv, kSplit, pv,
                                  nullptr, 0, 0, 0, g32_len ? w.dm_vid : w.dm_vid + (size_t)t * Hv, d.AM * Hv, w.dc_v, Hv, last ? 1 : 0, B, Hv,
                                  st, nullptr, len32(w.shift_v, t, 1)));
      if (t > 0)
        MMQG_TRY(GemmCall(acts, Gv, false, P.vid_w_hh, Hv, false, B, Hv, Gv, w.dh_rec_vid, Hv).split(kSplit, pv).w().run(st));
    }
    if (d.T_v > 1)
      MMQG_TRY(GemmCall(w.acts_v + (size_t)B * Gv, Gv, true, w.hs_v + (size_t)B * Hv, Hv, false, Gv, Hv, (d.T_v - 1) * B,
                        Gd.vid_w_hh, Hv).run(st));
    else
      MMQG_CUDA(cudaMemsetAsync(Gd.vid_w_hh, 0, sizeof(float) * (size_t)Gv * Hv, st));
    for (int t = 0; t < d.T_v; ++t)
      MMQG_TRY(GemmCall(w.acts_v + (size_t)t * B * Gv, Gv, true, g32_len ? w.frames_tm + (size_t)t * B * d.F_v : batch->frames + (size_t)t * d.F_v,
                        g32_len ? d.F_v : d.T_v * d.F_v, false, Gv, d.F_v, B, Gd.vid_w_ih, d.F_v).accumulate(t > 0).run(st));
    MMQG_TRY(colsum(w.acts_v, Gv, Gd.vid_b_ih, Gd.vid_b_hh, d.T_v * B, Gv, 0.f, st));
    return 0;
  }

  // ---- phase 3: text LSTM stack BPTT (encoder.py:95-100) + encoder-side embedding gradient ----
  for (int l = L - 1; l >= 0; --l) {
    const int I = l == 0 ? E : H;
    if (g32_drop_p > 0.f && l < L - 1)   // dx_text holds d/d(dropped h_l): back through the mask
      MMQG_TRY(dropout_scale_f32(w.dx_text, 1, 0, (long long)d.T_t * B * H, g32_drop_seed, w.seed_ctr, kSidText32 + l, 0, g32_drop_p, st));
    for (int t = d.T_t - 1; t >= 0; --t) {
      StepGemmScope step_scope;
      PdlScope pdl_scope(g32_tc && pdl_enabled());
      const bool last = t == d.T_t - 1;
      float* acts = w.acts_text[l] + (size_t)t * B * G;
      // At the last encoder step the recurrent gradient is the decoder's gradient w.r.t. its
      // initial state (train.py:169), plus, for the top layer, the step-0 attention query.
      const float* dh0 = last ? w.dh_rec[l] : w.dh_rec_enc;
      const float* dh1 = (last && l == L - 1) ? w.dq_h : nullptr;
      const float* dh2 = l == L - 1 ? (g32_len ? w.dm_txt : w.dm_txt + (size_t)t * H) : w.dx_text + (size_t)t * B * H;
      const int ldh2 = l == L - 1 ? d.TM * H : H;
      MMQG_TRY(lstm_pointwise_bwd(acts, G, t > 0 ? w.cs_text[l] + (size_t)t * B * H : nullptr, H,
                                  w.cs_text[l] + (size_t)(t + 1) * B * H, H, dh0, H, kSplit, ps, dh1, H, kSplit, ps, dh2,
                                  ldh2, w.dc[l], H, 0, B, H, st, x3dg && t > 0 ? w.x3_dg : nullptr, len32(w.shift_t, t, l == L - 1 ? 1 : 0)));
      if (t > 0)
        MMQG_TRY(GemmCall(acts, G, false, P.text_w_hh[l], H, false, B, H, G, w.dh_rec_enc, H).split(kSplit, ps).w()
                     .presplit(x3dg ? w.x3_dg : nullptr).run(st));
    }
    const float* dG = w.acts_text[l];
    const float* X = l == 0 ? w.x0_text : (g32_drop_p > 0.f ? w.xdrop_text[l - 1] : w.hs_text[l - 1] + (size_t)B * H);
    MMQG_TRY(GemmCall(dG, G, true, X, I, false, G, I, d.T_t * B, Gd.text_w_ih[l], I).run(st));
    if (d.T_t > 1)
      MMQG_TRY(GemmCall(dG + (size_t)B * G, G, true, w.hs_text[l] + (size_t)B * H, H, false, G, H, (d.T_t - 1) * B,
                        Gd.text_w_hh[l], H).run(st));
    else
      MMQG_CUDA(cudaMemsetAsync(Gd.text_w_hh[l], 0, sizeof(float) * (size_t)G * H, st));
    MMQG_TRY(colsum(dG, G, Gd.text_b_ih[l], Gd.text_b_hh[l], d.T_t * B, G, 0.f, st));
    // input gradient for the layer below (or the embedding rows)
    MMQG_TRY(GemmCall(dG, G, false, P.text_w_ih[l], I, false, d.T_t * B, I, G, w.dx_text, I).w().run(st));
  }
  MMQG_TRY(embedding_scatter_add(Gd.emb, w.idx_ctx, w.dx_text, d.T_t * B, E, d.V, st));
  return 0;
}

static int decode_impl(const mmqg_dims* dp, const mmqg_tensors* params, const mmqg_batch* batch, void* workspace,
                       size_t workspace_bytes, int64_t* tokens_out, int max_len, int mode, void* stream, int sample,
                       unsigned long long seed);

int mmqg_greedy_decode(const mmqg_dims* dp, const mmqg_tensors* params, const mmqg_batch* batch, void* workspace,
                       size_t workspace_bytes, int64_t* tokens_out, int max_len, int mode, void* stream) {
  return decode_impl(dp, params, batch, workspace, workspace_bytes, tokens_out, max_len, mode, stream, 0, 0);
}

int mmqg_sample_decode(const mmqg_dims* dp, const mmqg_tensors* params, const mmqg_batch* batch, void* workspace,
                       size_t workspace_bytes, int64_t* tokens_out, int max_len, unsigned long long seed, int mode,
                       void* stream) {
  return decode_impl(dp, params, batch, workspace, workspace_bytes, tokens_out, max_len, mode, stream, 1, seed);
}

static int decode_impl(const mmqg_dims* dp, const mmqg_tensors* params, const mmqg_batch* batch, void* workspace,
                       size_t workspace_bytes, int64_t* tokens_out, int max_len, int mode, void* stream, int sample,
                       unsigned long long seed) {
  MMQG_TRY(check_dims(dp));
  const mmqg_dims& d = *dp;
  MMQG_TRY(check_tensors(d, params, "params"));
  MMQG_REQUIRE(batch && batch->context && batch->frames && batch->audio, "batch: null pointer");
  MMQG_REQUIRE(workspace && tokens_out && max_len > 0, "greedy: bad args");
  MMQG_REQUIRE(mode == MMQG_MODE_FP32 || mode == MMQG_MODE_BF16 || mode == MMQG_MODE_FP32_TC, "unknown mode %d", mode);
  if (mode == MMQG_MODE_BF16)
    return greedy_decode_bf16(d, *params, *batch, workspace, workspace_bytes, tokens_out, max_len, as_stream(stream), sample, seed);
  g32_len = batch->ctx_len || batch->n_frames;
  g32_drop_p = 0.f;      // decoding is eval mode: no dropout
  g32_tc = mode == MMQG_MODE_FP32_TC;
  Ws w = carve(d, max_len, workspace, g32_tc);
  if (g32_tc) f32x3_bind(w.x3_arena, w.x3_arena_bytes, w.x3_sa, w.x3_sa_bytes, w.x3_sb, w.x3_sb_bytes);
  if (w.bytes > workspace_bytes)
    return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  cudaStream_t st = as_stream(stream);
  const mmqg_tensors& P = *params;
  const int B = d.B, H = d.H, G = 4 * d.H, E = d.E, Q = d.E + d.H, C = d.H + d.H_a + d.H_v, X0 = E + C, Sp = w.S_pad;

  if (g32_len) MMQG_TRY(prep_lengths(batch->ctx_len, nullptr, batch->n_frames, w.shift_t, w.shift_v, w.row_w, B, d.T_t, d.T_v, 1, st));
  MMQG_TRY(build_indices(batch->context, nullptr, w.idx_ctx, nullptr, nullptr, B, d.T_t, 0, st, g32_len ? w.shift_t : nullptr));
  MMQG_TRY(encoder_forward(d, P, *batch, w, st));
  MMQG_TRY(handoff_state(d, w, st));
  MMQG_TRY(pack_attention(d, P, w, st));
  for (int l = 0; l < d.L; ++l) MMQG_TRY(add2(P.dec_b_ih[l], P.dec_b_hh[l], w.bsum_dec[l], G, st));
  const AttnShape as = attn_shape(d);
  for (int t = 0; t < max_len; ++t) {
    // x = [emb(word) | contexts]  (decoder.py:75,99); word = <start> at t=0 else previous argmax
    if (t == 0) MMQG_TRY(fill_i64(w.idx_cur, B, 1, st));           // <start>, train.py:84
    const int64_t* words = w.idx_cur;
    MMQG_TRY(embedding_gather(P.emb, words, w.xcat, X0, B, E, d.V, st));
    const float* htop_prev = w.hs_dec[d.L - 1] + (size_t)t * B * H;
    float* sc = w.attn_all + (size_t)t * B * Sp;
    MMQG_TRY(GemmCall(w.xcat, X0, false, w.attn_w_cat, Q, true, B, Sp, E, sc, Sp)
                 .second(htop_prev, H, w.attn_w_cat + E, Q, H).bias(w.attn_b_cat).w().run(st));
    MMQG_TRY(attn_fwd(sc, Sp, w.m_txt, w.m_aud, w.m_vid, w.xcat + E, X0, as, st));
    for (int l = 0; l < d.L; ++l) {
      float* acts = w.acts_dec[l] + (size_t)t * B * G;
      const float* hprev = w.hs_dec[l] + (size_t)t * B * H;
      const float* x = l == 0 ? w.xcat : w.hs_dec[l - 1] + (size_t)(t + 1) * B * H;
      const int I = l == 0 ? X0 : H;
      PreSpec ps;
      GemmCall g(x, I, false, P.dec_w_ih[l], I, true, B, G, I, acts, G);
      g.second(hprev, H, P.dec_w_hh[l], H, H);
      if (g32_tc) MMQG_TRY(step_pre_tc(g, w.pre_part, B, G, w.bsum_dec[l], false, &ps, st));
      else MMQG_TRY(g.bias(w.bsum_dec[l]).w().run(st));
      MMQG_TRY(lstm_pointwise_fwd(acts, G, w.cs_dec[l] + (size_t)t * B * H, H, w.cs_dec[l] + (size_t)(t + 1) * B * H, H,
                                  w.hs_dec[l] + (size_t)(t + 1) * B * H, H, nullptr, 0, B, H, st, ps));
    }
    const float* htop = w.hs_dec[d.L - 1] + (size_t)(t + 1) * B * H;
    for (int r0 = 0; r0 < B; r0 += w.Rc) {
      const int rc = B - r0 < w.Rc ? B - r0 : w.Rc;
      MMQG_TRY(GemmCall(htop + (size_t)r0 * H, H, false, P.out_w, H, true, rc, d.V, H, w.logits, d.V).bias(P.out_b).w().run(st));
      if (sample)
        MMQG_TRY(sample_rows(w.logits, d.V, tokens_out + (size_t)r0 * max_len + t, max_len, w.idx_cur + r0, rc, d.V, seed,
                             (unsigned long long)t, r0, B, st));
      else
        MMQG_TRY(argmax_rows(w.logits, d.V, tokens_out + (size_t)r0 * max_len + t, max_len, w.idx_cur + r0, rc, d.V, st));
    }
  }
  return 0;
}

}  // extern "C"

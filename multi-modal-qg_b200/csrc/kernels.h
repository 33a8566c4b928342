// Internal (non-ABI) declarations shared by the translation units of libmmqg.so.
#pragma once
#include "common.cuh"

namespace mmqg {

int gemm_f32(const mmqg_gemm_args& a, cudaStream_t st);
int gemm_bf16(const mmqg_gemm_bf16_args& a, cudaStream_t st);
// fp32 contraction on the tensor cores by bf16 splitting (gemm_f32x3.cu); b_const: B/B2 are constant during the current
// C-ABI call and their split copy is cached in the arena bound by f32x3_bind (which also forgets the previous call's copies)
int gemm_f32x3(const mmqg_gemm_args& a, bool b_const, cudaStream_t st);
void f32x3_bind(void* arena, size_t arena_bytes, void* sa, size_t sa_bytes, void* sb, size_t sb_bytes);
// the NEXT gemm_f32x3 call takes its A operand already split ([hi | lo], K contiguous, pitch 2*K, K % 8 == 0, no second pair)
void f32x3_presplit_a(const void* a_split);
size_t f32x3_split_bytes(int mn, long long n, long long K);

// Variable-length batches in the recurrent kernels (persistent kernels: rows of a launch; fp32 cell kernels: t_base = the step): sample (row) m is right-aligned in time,
// its steps t_base + t < shift[m] are masked (state held at zero, no gradient); with mem_shift the
// batch-major memory output (forward) / external gradient input (backward) is indexed by the
// sample's own position t_base + t - shift[m] instead of the time step.
struct LenSpec {
  const int* shift = nullptr;
  int t_base = 0;
  int mem_shift = 0;
};
// Pre-activations handed to the forward cell kernel as split-K partial sums: pre = (gates if
// add_gates) + bias + sum_k part[k*stride + b*ld + .]; the activated gates still land in `gates`.
struct PreSpec {
  const float* part = nullptr;
  int n_part = 0, ld = 0;
  long long stride = 0;
  const float* bias = nullptr;
  int add_gates = 0;
  void* h_split = nullptr;   // fp32 cell kernel (tensor-core parity mode): h also as bf16 [hi(H) | lo(H)] rows of pitch 2H = the next
                             // step's pre-split A operand (gemm_f32x3.cu), saving that step's split launch
  int no_save = 0;      // forward-only caller: the activated gates need not be written back (honoured by the vectorised bf16 kernel)
};
int lstm_pointwise_fwd(float* gates, int ldg, const float* c_prev, int ldcp, float* c_out, int ldc, float* h_out,
                       int ldh, float* h2, int ldh2, int B, int H, cudaStream_t st, PreSpec ps = PreSpec(), LenSpec len = LenSpec());
int lstm_pointwise_bwd(float* acts, int ldg, const float* c_prev, int ldcp, const float* c_new, int ldc,
                       const float* dh0, int ldh0, int n0, long long s0, const float* dh1, int ldh1, int n1, long long s1,
                       const float* dh2, int ldh2, float* dc, int lddc, int dc_is_zero, int B, int H,
                       cudaStream_t st, void* dg_split = nullptr,       // dg_split: d pre-activations also as bf16 [hi(4H) | lo(4H)], pitch 8H
                       LenSpec len = LenSpec());
int embedding_gather(const float* emb, const int64_t* idx, float* out, int ldo, int N, int E, int V, cudaStream_t st);
int embedding_scatter_add(float* demb, const int64_t* idx, const float* dx, int N, int E, int V, cudaStream_t st);
int nll_rows(float* logits, int ldl, const int64_t* targets, long long tgt_stride, float* nll, int R, int V,
             float scale, cudaStream_t st, const float* row_w = nullptr);     // row_w: per-row loss weight (0 beyond a sample's target length)
int frames_to_time_major_f32(const float* frames, float* out, const int* shift_v, int B, int T_v, int F, cudaStream_t st);
int argmax_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int64_t* tokens2, int R, int V,
                cudaStream_t st);
int sample_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int64_t* tokens2, int R, int V,
                unsigned long long seed, unsigned long long step, int row0, int R_total, cudaStream_t st);
int colsum(const float* X, int ldx, float* out, float* out2, int M, int N, float beta, cudaStream_t st);
int add2(const float* a, const float* b, float* y, int n, cudaStream_t st);
int build_indices(const int64_t* ctx, const int64_t* tgt, int64_t* idx_ctx, int64_t* idx_dec, int64_t* tgt_tm, int B,
                  int T_t, int T_q, cudaStream_t st, const int* shift_t = nullptr);
int sum_scale(const float* x, int n, float scale, float* out, cudaStream_t st);
int fill_i64(int64_t* p, int n, int64_t v, cudaStream_t st);

struct AttnShape {
  int B, TM, AM, H, H_a, H_v, T_t, T_v;
  // bf16 mode: optional bf16 copies of the outputs that feed tensor-core GEMMs
  void* ctx16; int ldctx16;     // attn_fwd: contexts
  void* ds16; int ldds16;       // attn_bwd: d loss / d scores
  // bf16 mode: bf16 copies of the text / video memories (same shapes); read instead of the fp32 ones
  const void* m_txt16; const void* m_vid16;
  // attn_bwd: dctx may arrive as `dctx_parts` split-K partial sums `dctx_part_stride` floats apart
  // (0/1 = a single tensor); the summed rows are then also written to dctx_sum (row pitch lddsum)
  int dctx_parts; long long dctx_part_stride; float* dctx_sum; int lddsum;
};
bool attn_fast_ok(const AttnShape& s, const void* M_txt, const void* M_aud, const void* M_vid);
int attn_fwd_fast(float* scores, int lds, const void* M_txt, const float* M_aud, const void* M_vid, bool mem_bf16, float* ctx,
                  int ldctx, const AttnShape& s, cudaStream_t st);
int attn_bwd_fast(const float* attn, float* ds_out, int lds, const float* dctx, int lddctx, const void* M_txt, const float* M_aud,
                  const void* M_vid, bool mem_bf16, const AttnShape& s, cudaStream_t st);

// bf16-mode elementwise kernels (pointwise_bf16.cu)
int cvt_f32_bf16_2d(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols,
                    int cols_dst, cudaStream_t st);
int embedding_gather_bf16(const float* emb, const int64_t* idx, void* out, int ldo, int N, int E, int E_pad, int V,
                          cudaStream_t st);
// Inter-layer dropout fused into the cell kernels: forward writes a dropped bf16 copy of h to
// `out` (row pitch ld); backward multiplies the dh1 term by the same mask (p > 0).  Element
// (b, j) of the call uses counter base + b*H + j of stream sid.
struct DropSpec {
  void* out = nullptr;
  int ld = 0;
  unsigned long long seed = 0, base = 0;
  const unsigned long long* ctr = nullptr;   // device call counter added to seed (fresh masks per step, also under graph replay)
  int sid = 0;
  float p = 0.f;
};
int lstm_pointwise_fwd_bf16(float* gates, int ldg, const float* c_prev, int ldcp, float* c_out, int ldc, void* h_out,
                            int ldh, float* h2, int ldh2, int B, int H, cudaStream_t st, DropSpec dr = DropSpec(),
                            PreSpec ps = PreSpec());
int lstm_pointwise_bwd_bf16(const float* acts, int ldg, const float* c_prev, int ldcp, const float* c_new, int ldc,
                            const float* dh0, int ldh0, int n0, long long s0, const float* dh1, int ldh1, int n1,
                            long long s1, const float* dh2, int ldh2, float* dc, int lddc, int dc_is_zero, void* dg,
                            int lddg, int B, int H, cudaStream_t st, DropSpec dr = DropSpec());
int colsum_bf16(const void* X, int ldx, float* out, float* out2, int M, int N, float beta, cudaStream_t st);
int dropout_bf16(const void* x, void* y, long long n, unsigned long long seed, const unsigned long long* ctr, int sid,
                 unsigned long long base, float p, cudaStream_t st);
int dropout_f32(const float* x, float* y, long long n, unsigned long long seed, const unsigned long long* ctr, int sid,
                unsigned long long base, float p, cudaStream_t st);
int bump_counter(unsigned long long* ctr, cudaStream_t st);
int dropout_scale_f32(float* x, int n_part, long long stride, long long n, unsigned long long seed, const unsigned long long* ctr, int sid, unsigned long long base,
                      float p, cudaStream_t st);
int reduce_partials(const float* part, int n_part, long long stride, float* out, int ldo, long long rows, int N, cudaStream_t st);
// variable-length batches (mmqg_batch.ctx_len / tgt_len / n_frames), pointwise_bf16.cu
int prep_lengths(const int* ctx_len, const int* tgt_len, const int* n_frames, int* shift_t, int* shift_v, float* row_w, int B,
                 int T_t, int T_v, int T_q, cudaStream_t st);
int frames_to_time_major_bf16(const float* frames, void* out, const int* shift_v, int B, int T_v, int F, cudaStream_t st);
int audio_pad(const float* audio, float* m_aud, const int* n_frames, int B, int T_v, int AM, int H_a, cudaStream_t st);
int dropout_mask(float* out, long long n, unsigned long long seed, int sid, float p, cudaStream_t st);
// persistent recurrent-cell kernels (lstm_persist.cu)
bool lstm_persist_ok(int B, int H);
bool lstm_persist_fwd_ok(int B, int H);     // forward kernel alone (two m-tiles per CTA)
int lstm_persist_fwd_ctas(int B, int H);    // CTAs of one forward launch
int device_sms();
int pack_whh(const float* w_hh, void* fwd_packed, void* bwd_packed, int H, cudaStream_t st);
extern thread_local int tl_ktag;      // debug: tag of the next persistent recurrent launch (lstm_persist.cu)
int sum_partials(const float* a, int na, const float* b, int nb, long long stride, float* y, int n, cudaStream_t st);
int lstm_seq_fwd_persist(float* gates, float* cs, void* hs, const void* wp_fwd, float* mem, void* mem16, long long mem_ld,
                         uint32_t* flags, int T, int B, int H, int load_c0, cudaStream_t st, DropSpec dr = DropSpec(),
                         bool zero_flags = true, LenSpec len = LenSpec());
int lstm_seq_bwd_persist(const float* acts, const float* cs, void* dg, const void* wp_bwd, const float* dh_ext,
                         long long ext_ts, long long ext_ld, const float* dh_last, const float* dc_last, uint32_t* flags,
                         int T, int B, int H, int has_next, float* dc_out, cudaStream_t st, DropSpec dr = DropSpec(),
                         bool zero_flags = true, LenSpec len = LenSpec());
// persistent decoder-step kernel, forward (dec_persist.cu)
struct DecPersistShape { int B, H, C, Sp, TM, AM, T_t, T_v, H_a, H_v, L; };
struct DecPersistArgs {
  DecPersistShape shape;
  int Tq;                                   // decoder steps of the whole sequence (sizes of the buffers below)
  float* attn_all; void* ctx16;             // (Tq*B, Sp) fp32, (Tq*B, C) bf16
  float* acts[3]; float* cs[3]; void* hs[3]; void* hdrop[3]; const float* bias[3];
  const void* w_hh[3]; const void* w_in[3]; // 16-unit gate-slice row order (pack_whh forward layout / pack_rows_gate16), bf16
  const void* wa_h;                         // (Sp, H) bf16
  const void* m_txt16; const void* m_vid16; const float* m_aud;
  uint32_t* flags;                          // dec_persist_flag_words() words, zeroed before the first launch of a sequence
  float drop_p; unsigned long long seed; const unsigned long long* ctr; int sid0;
};
bool dec_persist_ok(const DecPersistShape& s);
int dec_persist_waves(const DecPersistShape& s);     // launches per call (1 = all groups co-resident)
size_t dec_persist_flag_words(const DecPersistShape& s, int Tq);
int pack_rows_gate16(const float* w, int ld, int K, int Kp, void* out, int H, cudaStream_t st);
int dec_seq_fwd_persist(const DecPersistArgs& a, int t0, int T, cudaStream_t st);
int dec_persist_rows(int B, int H);                  // batch rows per group (64 / 128), 0 = shape not supported
// persistent decoder-step kernel, backward through time (dec_persist_bwd.cu)
struct DecPersistBwdArgs {
  DecPersistShape shape;
  int Tq;
  const float* attn_all; float* ds_all; void* ds16; float* dctx_all;     // (Tq*B, Sp) fp32 in, (Tq*B, Sp) fp32 / bf16 out, (Tq*B, C) fp32 out
  const float* dhtop;                                                   // (Tq*B, H) fp32 from the loss head
  const float* acts[3]; const float* cs[3]; void* dg[3];                // saved activations / cell states; dG_l (Tq*B, 4H) bf16 out
  float* dc[3]; float* dh_rec[3];                                       // (B, H) fp32 state handed over between calls / to the encoders
  const void* wT[3];                                                    // packed transposed bf16 weights per layer (pack_decb_weights)
  const void* wahT; int Spp;                                            // attention Linears, state columns, transposed: (H, Spp), Spp % 64 == 0
  const void* m_txt16; const void* m_vid16; const float* m_aud;
  uint32_t* flags;                                                      // dec_bwd_persist_flag_words() words, zeroed before the first call
  float drop_p; unsigned long long seed; const unsigned long long* ctr; int sid0;
};
bool dec_bwd_persist_ok(const DecPersistShape& s);
size_t dec_bwd_persist_flag_words(const DecPersistShape& s, int Tq);
int dec_seq_bwd_persist(const DecPersistBwdArgs& a, int t_lo, int t_hi, cudaStream_t st);
int transpose_f32_bf16(const float* src, long long ld_src, int R, int Cc, void* dst, long long ld_dst, cudaStream_t st);
size_t dec_bwd_weight_rows(int H, int l);
int pack_decb_weights(const float* w_hh, int H, const float* w_in, long long ld_in, int n_in_total, int l, void* out, cudaStream_t st);
// 3x3 / stride-1 fast path of the conv stack for the reference's channel chain (convstack3.cu)
bool conv3_fast_ok(int Cin, int Cout, int K, int stride, int Win, int max_win);      // max_win <= 0: any width
int conv3_relu_fwd(const float* x, const float* in_scale, const float* in_shift, const float* w, const float* b, float* y, float* stats,
                   int N, int Cin, int Hin, int Win, int Cout, int parts, cudaStream_t st);
int conv3_bwd_x(const float* dz, const float* w, float* dxn, int N, int Cin, int Hin, int Win, int Cout, cudaStream_t st);
int conv3_bwd_w(const float* x, const float* in_scale, const float* in_shift, const float* dz, float* dw, float* db, int N, int Cin,
                int Hin, int Win, int Cout, cudaStream_t st);
// bf16-mode orchestration (engine_bf16.cu)
size_t train_workspace_bytes_bf16(const mmqg_dims& d, int T_q);
int check_dims_bf16(const mmqg_dims& d);
int train_forward_bf16(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                       size_t workspace_bytes, float* loss_out, int want_grads, mmqg_tensors* grads, float grad_scale,
                       float dropout_p, unsigned long long seed, cudaStream_t st);
int greedy_decode_bf16(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace, size_t workspace_bytes,
                       int64_t* tokens_out, int max_len, cudaStream_t st, int sample = 0, unsigned long long seed = 0);
size_t greedy_workspace_bytes_bf16(const mmqg_dims& d, int max_len);
int train_backward_bf16(const mmqg_dims& d, const mmqg_tensors& P, const mmqg_batch& bt, void* workspace,
                        size_t workspace_bytes, mmqg_tensors& Gd, int phase, float dropout_p, unsigned long long seed,
                        cudaStream_t st, cudaEvent_t const* ready = nullptr);
// fused vocabulary projection + log-softmax + NLL, backward, greedy arg-max (vocab_nll.cu)
size_t vocab_stat_floats(int R, int V);
int vocab_chunk_rows(int R, int Vp);
int vocab_nll_fwd(const void* X, int ldx, const void* W, int ldw, const float* bias, const int64_t* targets, const float* row_w, int R,
                  int V, int H, float dscale, float* nll, float* lse, float* row_scale, float* stat_a, float* stat_b, float* tgt_logit,
                  cudaStream_t st);
int vocab_dlogits(const void* X, int ldx, const void* W, int ldw, const float* bias, const int64_t* targets, const float* lse,
                  const float* row_scale, int R, int V, int H, void* dlogits, int lddl, cudaStream_t st);
int vocab_nll_bwd(const void* X, int ldx, const void* W, int ldw, const float* bias, const int64_t* targets, const float* lse,
                  const float* row_scale, int R, int V, int H, void* dl, int Vp, int rc_rows, float* part, float* dH, int lddh, float* dW,
                  float* db, bool accumulate, cudaStream_t st);
int vocab_argmax(const void* X, int ldx, const void* W, int ldw, const float* bias, int R, int V, int H, float* stat_a, int* stat_i,
                 int64_t* tokens, long long tok_stride, int64_t* tokens2, cudaStream_t st);
int nll_rows_bf16(const float* logits, int ldl, const int64_t* targets, float* nll, int R, int V, float scale,
                  void* dlogits, int lddl, cudaStream_t st, const float* row_w = nullptr);
int attn_fwd(float* scores, int lds, const float* M_txt, const float* M_aud, const float* M_vid, float* ctx,
             int ldctx, const AttnShape& s, cudaStream_t st);
int attn_bwd(const float* attn, float* ds_out, int lds, const float* dctx, int lddctx, const float* M_txt,
             const float* M_aud, const float* M_vid, float* dM_txt, float* dM_vid, const AttnShape& s,
             cudaStream_t st);
// Hoisted memory gradients: dM_txt(b,j,:) = sum_t attn(t,b,j) dctx(t,b,txt part), same for video.
int attn_dmem(const float* attn_all, int lds, const float* dctx_all, int lddctx, float* dM_txt, float* dM_vid, int T_q,
              const AttnShape& s, cudaStream_t st);

}  // namespace mmqg

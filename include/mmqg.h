/*
 * mmqg.h -- C ABI of libmmqg.so: the B200 (sm_100a) implementation of the
 * multi-modal-qg training hot path.
 *
 * The reference has no native layer and no FFI: its "plugin API"
 * for this path is the set of torch.nn.Module classes in model/encoder.py and
 * model/decoder.py.  This header is therefore the boundary a maintainer of the
 * reference would bind from Python with ctypes (INTEGRATION.md shows the stub); each
 * entry point names the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer into caller-owned memory
 *     (e.g. tensor.data_ptr()); the library never allocates, frees or retains them.
 *   - scratch memory comes from a caller-provided workspace whose size is returned by
 *     the matching *_workspace_bytes() function.
 *   - `stream` is a cudaStream_t passed as void*.  All work is enqueued asynchronously
 *     and is ordered on it: there are no hidden synchronisations.  The bf16 train path forks
 *     library-owned auxiliary streams from `stream` and joins them back with events before
 *     the call returns, so callers (and CUDA-graph capture) see single-stream semantics.
 *   - one process drives one GPU (the data-parallel layout of SURVEY.md section 8e): the helper
 *     streams and events are per process, calls are not re-entrant across host threads.
 *   - return value: 0 = enqueued OK, otherwise an mmqg_status; mmqg_last_error()
 *     gives a thread-local message.
 *   - parameters are fp32 in PyTorch layout: LSTM weights (4H, in) row-major with gate
 *     blocks i,f,g,o; Linear weights (out, in) (SURVEY.md section 8b state_dict contract).
 *   - tokens are int64 (reference utils/custom_transforms.py:25).
 */
#ifndef MMQG_H
#define MMQG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMQG_ABI_VERSION 3
#define MMQG_MAX_LAYERS 4

typedef enum {
  MMQG_OK = 0,
  MMQG_ERR_BAD_ARG = 1,      /* null pointer, non-positive size, unsupported shape */
  MMQG_ERR_WORKSPACE = 2,    /* workspace too small */
  MMQG_ERR_CUDA = 3,         /* a CUDA runtime call failed; see mmqg_last_error() */
  MMQG_ERR_ARCH = 4          /* device is not sm_100 */
} mmqg_status;

/* precision modes (SURVEY.md section 8d "Precision modes") */
#define MMQG_MODE_FP32 0     /* fp32 storage, fp32 FMA accumulation: the parity mode   */
#define MMQG_MODE_BF16 1     /* bf16 operands on tcgen05 tensor cores, fp32 accumulate */
#define MMQG_MODE_FP32_TC 2  /* the fp32 parity path with every contraction on the tcgen05 tensor cores: operands split
                              * into bf16 (hi, lo), three products hi*lo + lo*hi + hi*hi accumulated in fp32 (~2^-16
                              * relative per product); fp32 storage, pointwise math and reductions as MMQG_MODE_FP32 */

/* Shapes (SURVEY.md section 8 symbol table; reference config.py:64-87). */
typedef struct {
  int B, T_t, T_v, T_q, V;
  int E, H, L, H_a, H_v, F_v, TM, AM;
} mmqg_dims;

/* All learnable tensors on the path.  Index order of attn_*: 0=text, 1=audio, 2=video
 * (the order AttnDecoder.forward returns the weights, reference decoder.py:107). */
typedef struct {
  float* emb;                                   /* (V,E)   train.py:236, shared          */
  float* text_w_ih[MMQG_MAX_LAYERS];            /* (4H,E|H)   encoder.py:91              */
  float* text_w_hh[MMQG_MAX_LAYERS];            /* (4H,H)                                */
  float* text_b_ih[MMQG_MAX_LAYERS];            /* (4H)                                  */
  float* text_b_hh[MMQG_MAX_LAYERS];
  float* vid_w_ih;                              /* (4H_v,F_v) encoder.py:54              */
  float* vid_w_hh;                              /* (4H_v,H_v)                            */
  float* vid_b_ih;
  float* vid_b_hh;
  float* attn_w[3];                             /* (TM|AM|AM, E+H) decoder.py:64-66      */
  float* attn_b[3];
  float* dec_w_ih[MMQG_MAX_LAYERS];             /* (4H, E+H+H_a+H_v | H) decoder.py:69   */
  float* dec_w_hh[MMQG_MAX_LAYERS];
  float* dec_b_ih[MMQG_MAX_LAYERS];
  float* dec_b_hh[MMQG_MAX_LAYERS];
  float* out_w;                                 /* (V,H)   decoder.py:70                 */
  float* out_b;                                 /* (V)                                   */
} mmqg_tensors;                                 /* used for parameters and for gradients */

/* One uniform-length batch (features in place of raw media). */
typedef struct {
  const int64_t* context;    /* (B,T_t)      token ids            train.py:164-165 */
  const int64_t* target;     /* (B,T_q)      question + <end>     train.py:174-175 */
  const float* frames;       /* (B,T_v,F_v)  frame features       encoder.py:69    */
  const float* audio;        /* (B,T_v,H_a)  VGGish features      encoder.py:122   */
  /* Optional per-sample lengths (device int32 (B), or NULL = every sample uses the full T):
   * sample b is context[b,:ctx_len[b]], target[b,:tgt_len[b]], frames/audio[b,:n_frames[b]],
   * 1 <= len <= T; what lies beyond is ignored.  Same result as running the reference's
   * per-sample loop (train.py:153-177) on each sample with its own lengths: encoder states
   * start from zero at the sample's first real token, memory rows beyond the length are the
   * zero padding of train.py:155-160, the loss sums over the sample's own target steps.
   * All three modes (MMQG_MODE_BF16: shapes the persistent recurrent kernels take). */
  const int* ctx_len;
  const int* tgt_len;
  const int* n_frames;
} mmqg_batch;

int mmqg_abi_version(void);
const char* mmqg_last_error(void);
/* 0 if device `dev` can run this library (compute capability 10.x), else MMQG_ERR_ARCH. */
int mmqg_device_ok(int dev);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
unsigned long long mmqg_launch_count(void);

/* Timing probe for bench.py's roofline figure: CUDA events are recorded around every launch
 * of ONE kernel class on the launching stream.  Classes: 1 per-timestep recurrent products,
 * 2 hoisted whole-sequence products, 3 LSTM cell pointwise, 4 attention step, 5 loss rows,
 * 6 embedding gather/scatter.  mmqg_probe_stop() waits for the recorded events and returns
 * the summed kernel time, the number of launches and the algorithmic FLOPs / bytes the
 * launches were booked with.  Not for use under CUDA-graph capture. */
int mmqg_probe_start(int kernel_class);
int mmqg_probe_stop(double* total_ms, unsigned long long* launches, double* flops, double* bytes);

/* ---- whole-path entry points (what bench.py and the sequence-level modules call) ---- */

size_t mmqg_train_workspace_bytes(const mmqg_dims* d, int mode);

/* Teacher-forced forward over the whole batch: encoder (reference encoder.py:95-100,69;
 * train.py:153-166), T_q decoder steps (decoder.py:74-107) and the summed batch-mean
 * cross entropy (train.py:171-175).  loss_out: device float[1] = sum_t mean_b NLL.
 * If want_grads != 0 the loss head also produces d loss/d h_top, d out_w, d out_b
 * (logits are produced and consumed in row chunks and never stored in full);
 * grad_scale multiplies every gradient (use B_local/B_global for data parallelism,
 * 1.0 otherwise).  dropout_p is torch.nn.LSTM's inter-layer dropout (0 disables; both modes).
 * Its masks are a function of seed + the number of forward calls with dropout made on
 * this workspace so far: the FIRST 8 BYTES of the bf16-mode workspace are that call counter
 * (uint64; the caller zeroes the workspace once, everything else in it is scratch), advanced on
 * the device, so every step -- also every replay of a captured CUDA graph -- draws fresh masks
 * and the backward call of the step regenerates the same ones. */
int mmqg_train_forward(const mmqg_dims* d, const mmqg_tensors* params, const mmqg_batch* batch,
                       void* workspace, size_t workspace_bytes, float* loss_out,
                       int want_grads, mmqg_tensors* grads, float grad_scale,
                       float dropout_p, unsigned long long seed, int mode, void* stream);

/* NOTE on the loss-head gradients (d out_w, d out_b): in MMQG_MODE_BF16 with the persistent decoder kernel they are
 * produced by the BACKWARD call of the step (phase 0 / 1 or mmqg_train_backward_events), not by mmqg_train_forward:
 * callers must not read or reduce them between the two calls.  The forward call still needs `grads` (unchanged ABI).
 *
 * Backward of the same step (reference train.py:177).  `phase` selects which gradient
 * groups to produce, in the order they become final, so the caller can start the
 * all-reduce of one group while the next is computed:
 *   1 decoder (attention Linears, decoder LSTM; also d/d memories, d/d encoder state)
 *   2 video LSTM      3 text LSTM + shared embedding
 * Phases must be called in order 1,2,3 (2 and 3 are independent of each other).
 * phase 0 runs the whole backward in one call; in bf16 mode the hoisted weight-gradient
 * products then run on an internal auxiliary stream beside the persistent BPTT kernels and
 * are joined back onto `stream` before the call returns (CUDA-graph capturable).
 * Every gradient tensor is overwritten, not accumulated (the reference zeroes grads
 * every iteration, train.py:149-151). */
int mmqg_train_backward(const mmqg_dims* d, const mmqg_tensors* params, const mmqg_batch* batch,
                        void* workspace, size_t workspace_bytes, mmqg_tensors* grads,
                        int phase, float dropout_p, unsigned long long seed, int mode, void* stream);

/* Same decode loop with the reference's 'sampling' strategy (evaluate.py:84-90): every step draws
 * the next token from softmax(logits) instead of taking the arg-max ('topk' in the reference is
 * topk(1), i.e. greedy).  Deterministic in (seed, step, row). */
int mmqg_sample_decode(const mmqg_dims* d, const mmqg_tensors* params, const mmqg_batch* batch,
                       void* workspace, size_t workspace_bytes, int64_t* tokens_out, int max_len,
                       unsigned long long seed, int mode, void* stream);

/* The whole backward (as phase 0, with its internal overlap) for data-parallel callers.
 * ready_events: array of 4 + L cudaEvent_t created by the caller (entries may be NULL), one per gradient group;
 * each is recorded at the point -- and on whichever internal stream -- where that group becomes final, so the
 * caller can make its communication stream wait on the event and start the group's all-reduce under the rest
 * of the backward:
 *   [0] loss head (out_layer): final since the forward call, or -- when the persistent decoder kernel is in use and
 *       the backward half of the loss head is deferred to this call -- once that half has run under the decoder BPTT
 *   [1] decoder (attention Linears + decoder LSTM)      [2] video LSTM
 *   [3 + k] text LSTM layer L-1-k, k = 0 .. L-1 (the top layer's BPTT and weight gradients finish first)
 *   [3 + L] the shared embedding (decoder- and encoder-side scatter-adds both landed: last of all).
 * All work is joined back onto `stream` before the call returns.  Data-parallel callers give every rank its
 * own dropout `seed`, otherwise local sample b draws the same masks on every rank. */
int mmqg_train_backward_events(const mmqg_dims* d, const mmqg_tensors* params, const mmqg_batch* batch,
                               void* workspace, size_t workspace_bytes, mmqg_tensors* grads,
                               void* const* ready_events, float dropout_p, unsigned long long seed,
                               int mode, void* stream);

size_t mmqg_greedy_workspace_bytes(const mmqg_dims* d, int max_len, int mode);

/* Greedy decode (reference train.py:101-110, evaluate.py:70-80): encoder pass, then
 * max_len decoder steps feeding back argmax(logits) (lowest index on ties).
 * tokens_out: device int64 (B,max_len).  d->T_q is ignored.  MMQG_MODE_FP32 reproduces the
 * reference's tokens exactly (fp32-accurate logits); MMQG_MODE_BF16 runs the tensor-core path
 * (about 6x faster) and can leave the fp32 token path where the top-2 logit margin is below
 * bf16 resolution. */
int mmqg_greedy_decode(const mmqg_dims* d, const mmqg_tensors* params, const mmqg_batch* batch,
                       void* workspace, size_t workspace_bytes, int64_t* tokens_out, int max_len,
                       int mode, void* stream);

/* Fused multi-tensor Adam over flat fp32 buffers of identical layout (reference train.py:179-181
 * with the three torch.optim.Adam(lr=1e-4) of train.py:265-267; no weight decay, no amsgrad).
 * Elements [rep_lo, rep_hi) receive the update twice -- the shared embedding is registered with
 * two of the reference's optimisers (train.py:236,245,255).  `state`: 3 device floats, zeroed once
 * by the caller: the step count (as int bits) and the two bias-correction scalars the call derives
 * from it on the device, so a captured CUDA graph replays correctly.  n must be a multiple of 4. */
int mmqg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                   long long rep_lo, long long rep_hi, float lr, float beta1, float beta2, float eps,
                   float* state, void* stream);

/* The multiplicative inter-layer dropout mask (0 or 1/(1-p)) the train path applies for a given
 * seed: stream `sid` = 10+l for the output of text-encoder layer l, 20+l for decoder layer l;
 * element index (t*B + b)*H + j.  For tests (the mask is counter-based, not ATen's Philox). */
int mmqg_dropout_mask(float* out, long long n, unsigned long long seed, int sid, float p, void* stream);

/* ---- building blocks (what the step-wise drop-in modules call) ---- */

/* C(M,N) = alpha * op(A) op(B) [+ op(A2) op(B2)] + beta * Cin + bias(N).
 * Row-major with leading dimensions.  transA: A is stored (K,M); transB: B is stored
 * (N,K) (the PyTorch Linear / LSTM weight layout).  A2/B2 (K2 > 0) extend the
 * reduction with a second operand pair, e.g. x W_ih^T + h W_hh^T (replaces the addmm
 * pairs inside aten::lstm and decoder.py:78,84,92,106). */
typedef struct {
  const float* A; const float* B; int lda, ldb, K;
  const float* A2; const float* B2; int lda2, ldb2, K2;
  float* C; int ldc;
  const float* Cin; int ldcin;
  const float* bias;
  int M, N;
  float alpha, beta;
  int transA, transB;
  int split_k; long long c_split_stride;   /* split_k>1: slice z writes C + z*stride */
} mmqg_gemm_args;
int mmqg_gemm_f32(const mmqg_gemm_args* a, void* stream);

/* The same contraction on the 5th-generation tensor cores (tcgen05.mma, bf16 operands, fp32
 * accumulation in tensor memory, operands staged by TMA).  A, B (and A2, B2) are bf16.
 *   a_mn_major = 0: A stored (M,K) row-major;  1: A stored (K,M) row-major
 *   b_mn_major = 0: B stored (N,K) row-major (PyTorch weight layout);  1: B stored (K,N)
 * Base pointers must be 16-byte aligned and leading dimensions multiples of 8 elements.
 * C is fp32 (c_bf16 = 0) or bf16 (c_bf16 = 1); Cin and bias are fp32.  split_k > 1 writes
 * fp32 partial results C + z * c_split_stride that the caller sums. */
typedef struct {
  const void* A; const void* B; int lda, ldb, K;
  const void* A2; const void* B2; int lda2, ldb2, K2;
  int a_mn_major, b_mn_major;
  void* C; int ldc; int c_bf16;
  const float* Cin; int ldcin;
  const float* bias;
  int M, N;
  float alpha, beta;
  int split_k; long long c_split_stride;
} mmqg_gemm_bf16_args;
int mmqg_gemm_bf16(const mmqg_gemm_bf16_args* a, void* stream);

/* out(n,:) = emb(idx(n),:) for n < N (embedding lookup, encoder.py:96, decoder.py:75). */
int mmqg_embedding_gather(const float* emb, const int64_t* idx, float* out, int N, int E, int V, void* stream);
/* demb(idx(n),:) += dx(n,:)  (dense embedding gradient). */
int mmqg_embedding_scatter_add(float* demb, const int64_t* idx, const float* dx, int N, int E, int V, void* stream);

/* LSTM cell pointwise part: gates (B,4H) hold the pre-activations x W_ih^T + h W_hh^T + b
 * (blocks i,f,g,o) and are overwritten with the activated gates; c_prev may be NULL (zeros).
 * c' = s(f) c + s(i) tanh(g);  h' = s(o) tanh(c').  h2 (optional) receives a second copy
 * of h' with its own leading dimension (the attention-memory row, train.py:166). */
int mmqg_lstm_pointwise_fwd(float* gates, int ldg, const float* c_prev, int ldcp, float* c_out, int ldc,
                            float* h_out, int ldh, float* h2, int ldh2, int B, int H, void* stream);

/* Backward of the pointwise part.  dh = sum of up to three addends (dh0 with n0 split-K
 * partials spaced s0 apart, dh1, dh2; NULL = absent).  dc (B,H) is read as d loss/d c'
 * (NULL = zeros) and overwritten with d loss/d c_prev.  acts (activated gates) are
 * overwritten with d loss/d pre-activations. */
int mmqg_lstm_pointwise_bwd(float* acts, int ldg, const float* c_prev, int ldcp, const float* c_new, int ldc,
                            const float* dh0, int ldh0, int n0, long long s0,
                            const float* dh1, int ldh1, int n1, long long s1, const float* dh2, int ldh2,
                            float* dc, int lddc, int dc_is_zero, int B, int H, void* stream);

/* Three location-attention heads of one decoder step (decoder.py:78-97): scores (B,ld)
 * hold q W^T + b for the slots [text TM | audio AM | video AM]; they are overwritten
 * with the three softmaxes (over ALL slots of each head: the reference's length mask
 * is a no-op).  ctx (B, H+H_a+H_v) = [a_txt M_txt ; a_aud M_aud ;
 * a_vid M_vid] (decoder.py:99 order).  Memories are (B,TM,H), (B,AM,H_a), (B,AM,H_v);
 * only the first T_t / T_v rows are read (the rest are zero padding, train.py:155-160). */
int mmqg_attn_fwd(float* scores, int lds, const float* M_txt, const float* M_aud, const float* M_vid,
                  float* ctx, int ldctx, int B, int TM, int AM, int H, int H_a, int H_v, int T_t, int T_v,
                  void* stream);
/* Backward: attn (B,ld) softmax outputs are overwritten with d loss/d scores; dM_txt /
 * dM_vid (same shapes as the memories) are accumulated into (+=).  */
int mmqg_attn_bwd(float* attn, int lds, const float* dctx, int lddctx,
                  const float* M_txt, const float* M_aud, const float* M_vid,
                  float* dM_txt, float* dM_vid, int B, int TM, int AM, int H, int H_a, int H_v,
                  int T_t, int T_v, void* stream);

/* Row-wise cross entropy of logits (R,V): nll(r) = logsumexp - logits(r,target(r))
 * (train.py:174 CrossEntropyLoss).  If dlogits_scale != 0 the logits are overwritten
 * with dlogits_scale * (softmax - onehot). */
int mmqg_nll_rows(float* logits, int ldl, const int64_t* targets, long long tgt_stride, float* nll,
                  int R, int V, float dlogits_scale, void* stream);
/* tokens(r) = argmax_v logits(r,v), lowest index on ties (train.py:107-108). */
int mmqg_argmax_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int R, int V, void* stream);
/* tokens(r) ~ softmax(logits(r,:)): the reference's 'sampling' strategy (evaluate.py:84-90,
 * np.random.choice(V, p=softmax)) as an inverse-CDF draw with the uniform u(seed, step, r) that
 * mmqg_sample_uniform returns (n = R entries of one step) -- for tests and other hosts. */
int mmqg_sample_rows(const float* logits, int ldl, int64_t* tokens, long long tok_stride, int R, int V,
                     unsigned long long seed, unsigned long long step, void* stream);
int mmqg_sample_uniform(float* out, int n, unsigned long long seed, unsigned long long step, void* stream);
/* ---- one LSTM layer over a whole sequence (K1 + K2 / K5): SURVEY.md section 8b minimum export set ----
 * What the reference does with T per-token aten::lstm calls (encoder.py:95-100 driven by train.py:164-166;
 * encoder.py:69; decoder.py:25-34) and autograd's T backward twins, as one forward and one backward call:
 * hoisted input projection (one tcgen05 GEMM over T*B rows) + persistent recurrent kernel, persistent BPTT
 * kernel + hoisted weight-gradient products.  bf16 operands, fp32 accumulation / cell state / results.
 * x (T*B, I) rows t*B + b; weights in PyTorch layout (4H, I) / (4H, H), gate blocks i,f,g,o; h0, c0 (B, H) or
 * NULL (zeros).  y (T*B, H) = h_t; hn, cn (B, H) may be NULL.  Needs H % 64 == 0, H <= 512 and
 * (H/16) * ceil(B/128) CTAs co-resident (MMQG_ERR_BAD_ARG otherwise: use the per-step building blocks).
 * The workspace carries the saved activations from _fwd to _bwd: keep it untouched in between. */
size_t mmqg_lstm_seq_workspace_bytes(int T, int B, int I, int H);
int mmqg_lstm_seq_fwd(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                      const float* h0, const float* c0, int T, int B, int I, int H, void* workspace,
                      size_t workspace_bytes, float* y, float* hn, float* cn, void* stream);
/* dy (T*B, H), dhn, dcn (B, H): gradients w.r.t. y, hn, cn (any may be NULL = zeros).  Outputs (overwritten):
 * dx (T*B, I) (NULL to skip), dw_ih, dw_hh, db_ih, db_hh (identical, as autograd gives both biases the same
 * gradient; db_hh may be NULL), dh0, dc0 (B, H) (NULL to skip). */
int mmqg_lstm_seq_bwd(const float* dy, const float* dhn, const float* dcn, const float* w_hh, int T, int B, int I, int H,
                      void* workspace, size_t workspace_bytes, float* dx, float* dw_ih, float* dw_hh, float* db_ih,
                      float* db_hh, float* dh0, float* dc0, void* stream);

/* ---- fused loss head (K4) and greedy arg-max (K6): SURVEY.md section 8b minimum export set ---- */

/* Scratch for the three calls below for R rows, vocabulary V, hidden size H (multiple of 8). */
size_t mmqg_vocab_workspace_bytes(int R, int V, int H);

/* Vocabulary projection + log-softmax + NLL (decoder.py:106 out_layer, train.py:174 CrossEntropyLoss) without
 * materialising the logits: h (R,H) and w (V,H) are bf16 row-major, bias fp32 (V); the logits tile
 * h w^T + bias lives in tensor memory, its epilogue keeps per-row (max, sum exp, target logit).
 * nll(r) = row_w(r) * (lse(r) - logit(r, targets(r)))   (row_w may be NULL = 1).
 * lse (R) and row_scale (R) = grad_scale * row_w are what mmqg_vocab_nll_bwd needs (either may be NULL). */
int mmqg_vocab_nll_fwd(const void* h_bf16, const void* w_bf16, const float* bias, const int64_t* targets, const float* row_w,
                       int R, int V, int H, float grad_scale, void* workspace, size_t workspace_bytes, float* nll, float* lse,
                       float* row_scale, void* stream);

/* Backward: d logits(r,:) = row_scale(r) * (softmax(logits(r,:)) - onehot(targets(r))) is re-formed tile by tile
 * (bf16, in L2-sized row chunks inside the workspace) and contracted: dH (R,H) = dZ w, dW (V,H) = dZ^T h,
 * db (V) = column sums of dZ; all fp32.  accumulate != 0 adds into dW / db instead of overwriting them. */
int mmqg_vocab_nll_bwd(const void* h_bf16, const void* w_bf16, const float* bias, const int64_t* targets, const float* lse,
                       const float* row_scale, int R, int V, int H, void* workspace, size_t workspace_bytes, float* dH,
                       float* dW, float* db, int accumulate, void* stream);

/* Greedy step (train.py:106-108): tokens(r * tok_stride) = argmax_v (h w^T + bias)(r, v), lowest index on ties;
 * the arg-max is taken in the projection's epilogue, logits are not stored. */
int mmqg_decode_step_argmax(const void* h_bf16, const void* w_bf16, const float* bias, int R, int V, int H, void* workspace,
                            size_t workspace_bytes, int64_t* tokens, long long tok_stride, void* stream);

/* ---- f2: conv stack of VideoConvLstmEncoder on raw frames (reference model/encoder.py:40-52, :58-67) -----------------
 * 4 x [Conv2d(k, stride) -> ReLU -> BatchNorm2d], MaxPool2d(k, k) behind the 2nd and 4th; NCHW fp32, channels <= 16, k <= 5.
 * Train-mode BatchNorm is split and folded into its neighbours (see csrc/convstack.cu):
 *   conv_relu_fwd : y = relu(conv2d(x * in_scale[c] + in_shift[c], w) + b)  (in_scale/in_shift = the previous BatchNorm, NULL = none);
 *                   stats (mmqg_conv_stats_parts() x 2*Cout floats, NULL = skip) receives per-block partial sums and sums of
 *                   squares of y per channel (no same-address atomics)
 *   bn_finalize   : the `nparts` partials, `count` = N*H*W elements -> scale = gamma*invstd, shift = beta - mean*scale, saved mean / invstd,
 *                   running_mean / running_var updated like torch.nn.BatchNorm2d (momentum, unbiased variance; NULL = skip);
 *                   with nparts > 1, scale and shift must be ONE contiguous (2, C) buffer (shift == scale + C): it holds the
 *                   partial sums while they are reduced
 *   bn_maxpool_fwd: out = maxpool_K(y * scale + shift) (kernel = stride = K, floor), idx = position of the first maximum in the window
 * Backward:
 *   maxpool_bwd   : dense gradient w.r.t. the BatchNorm output from the pooled gradient
 *   bn_relu_bwd   : dz = gradient w.r.t. the conv output (through BatchNorm, batch statistics included, and ReLU);
 *                   sums (2*C floats) receives d beta = sum d and d gamma = sum d*xhat; sums == NULL: eval-mode BatchNorm
 *   conv_bwd_w    : dw (Cout,Cin,K,K), db (Cout) from dz and the (re-normalised) layer input
 *   conv_bwd_x    : gradient w.r.t. the normalised layer input = the previous BatchNorm's output */
int mmqg_conv_relu_fwd(const float* x, const float* in_scale, const float* in_shift, const float* w, const float* b, float* y,
                       float* stats, int N, int Cin, int Hin, int Win, int Cout, int K, int stride, void* stream);
int mmqg_conv_stats_parts(int N, int Hin, int Win, int K, int stride);
int mmqg_bn_finalize(const float* stats_parts, int nparts, long long count, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                     float* invstd, int C, void* stream);
int mmqg_bn_maxpool_fwd(const float* y, const float* scale, const float* shift, float* out, unsigned char* idx, int N, int C, int H,
                        int W, int K, void* stream);
int mmqg_maxpool_bwd(const float* dpool, const unsigned char* idx, float* dbn, int N, int C, int H, int W, int K, void* stream);
/* mmqg_maxpool_bwd followed by mmqg_bn_relu_bwd in one pass: the gradient w.r.t. the BatchNorm output is taken from the pooled
 * gradient and the saved arg-max on the fly, the dense (N,C,H,W) intermediate is never written (layers followed by a MaxPool2d). */
int mmqg_bn_relu_pool_bwd(const float* y, const float* mean, const float* invstd, const float* gamma, const float* dpool,
                          const unsigned char* idx, int K, float* dz, float* sums, int N, int C, int H, int W, void* stream);
int mmqg_bn_relu_bwd(const float* y, const float* mean, const float* invstd, const float* gamma, const float* dbn, float* dz,
                     float* sums, int N, int C, int H, int W, void* stream);
int mmqg_conv_bwd_w(const float* x, const float* in_scale, const float* in_shift, const float* dz, float* dw, float* db, int N,
                    int Cin, int Hin, int Win, int Cout, int K, int stride, void* stream);
int mmqg_conv_bwd_x(const float* dz, const float* w, float* dxn, int N, int Cin, int Hin, int Win, int Cout, int K, int stride,
                    void* stream);

/* Gradient exchange in bf16 (multi-GPU; the reference has no parallelism): a flat fp32 gradient
 * bucket is rounded to bf16 before the NCCL all-reduce and widened back afterwards, halving the bytes on NVLink.
 * pack: dst[i] = bf16(src[i]);  unpack: dst[i] = float(src[i]).  n elements, any alignment of n. */
int mmqg_pack_bf16(const float* src, void* dst_bf16, long long n, void* stream);
int mmqg_unpack_bf16(const void* src_bf16, float* dst, long long n, void* stream);

/* In-place all-reduce (sum) of a gradient bucket that lives in NVSwitch multicast memory (the reference has
 * no exchange step).  mc = MULTICAST address of the bucket (the same symmetric allocation mapped on every rank, 16-byte
 * aligned), n floats (multiple of 4); signal_pads = device array of `world` pointers to the ranks' signal pads (peer-mapped,
 * >= ctas * world * 4 bytes each, zero between calls).  Every rank makes the same sequence of calls with the same n and ctas.
 * One kernel of `ctas` CTAs: barrier across ranks, multimem.ld_reduce + multimem.st over the rank's slice, barrier. */
int mmqg_allreduce_multimem(float* mc, long long n, void* const* signal_pads, int rank, int world, int ctas, void* stream);

/* out(n) = sum_m X(m,n)  (bias gradients). */
int mmqg_colsum(const float* X, int ldx, float* out, int M, int N, float beta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMQG_H */

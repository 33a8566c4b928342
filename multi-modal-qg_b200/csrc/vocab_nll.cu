// K4: fused vocabulary projection + log-softmax + NLL, its backward, and the greedy arg-max
// (reference decoder.py:106 `out_layer`, train.py:174 CrossEntropyLoss, train.py:107-108 argmax).
//
// Forward.  logits = H_top W_out^T + b is formed tile by tile (128 rows x 256 vocabulary columns) in
// tensor memory by the persistent tcgen05 GEMM (gemm_tc_persist.cu) and is NEVER written anywhere:
// the epilogue reduces every row of the tile to (max, sum exp(x - max), target logit) in registers
// (VE_STATS); nll_merge_kernel folds the ceil(V/256) partials of a row into lse and the NLL.
// HBM traffic: H_top (R x H bf16) + W_out (V x H bf16) in, 2 x ceil(V/256) floats per row out.
//
// Backward.  d logits = scale * (softmax - onehot) needs the row's lse, i.e. the whole forward pass,
// so the logits tile is RECOMPUTED (one more pass of the same GEMM) and its epilogue (VE_DLOGITS)
// emits the bf16 d-logits operand for the two remaining products dH = dZ W_out and
// dW_out = dZ^T H_top (+ db = column sums).  The operand is produced in row chunks small enough to
// stay in the 126 MB L2 (<= MMQG_LH_BYTES, default 24 MB), written once and consumed by three
// kernels straight from L2 before the next chunk overwrites the same addresses.  A single kernel
// that also feeds the dH / dW_out MMAs from tensor memory was costed and rejected:
// it has to re-form the logits tile once per output block of BOTH products (4 x the S = H W^T
// work), is bound by the L2->SM ingest of those operands, and would hold 2 x 148 CTAs of ~200 KB
// shared memory for ~0.4 ms beside the latency-bound decoder loop.
#include <cuda_bf16.h>
#include "kernels.h"
#include "tc_common.cuh"

namespace mmqg {

typedef __nv_bfloat16 bf16;

__global__ void nll_merge_kernel(const float* __restrict__ stat_a, const float* __restrict__ stat_b, const float* __restrict__ tgt_logit,
                                 const float* __restrict__ row_w, int tiles_n, int R, float dscale, float* __restrict__ nll,
                                 float* __restrict__ lse_out, float* __restrict__ row_scale) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float m = -INFINITY;
  for (int i = 0; i < tiles_n; ++i) m = fmaxf(m, stat_a[(size_t)i * R + r]);
  float s = 0.f;
  for (int i = 0; i < tiles_n; ++i) {
    const float a = stat_a[(size_t)i * R + r];
    if (a > -INFINITY) s += stat_b[(size_t)i * R + r] * expf(a - m);
  }
  const float lse = m + logf(s);
  const float w = row_w ? row_w[r] : 1.f;        // 0 for target steps beyond the sample's own length
  nll[r] = w * (lse - tgt_logit[r]);
  if (lse_out) lse_out[r] = lse;
  if (row_scale) row_scale[r] = dscale * w;
}

__global__ void argmax_merge_kernel(const float* __restrict__ stat_a, const int* __restrict__ stat_i, int tiles_n, int R,
                                    int64_t* __restrict__ tokens, long long tok_stride, int64_t* __restrict__ tokens2) {
  // one warp per row, lanes over the column tiles (a thread per row walked its ~40 tiles as one chain of dependent loads:
  // 14.6 us per greedy step at B = 1024)
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < tiles_n; i += 32) {     // tiles in increasing column order: strict > keeps the lowest index on ties
    const float a = stat_a[(size_t)i * R + r];
    if (a > bv) { bv = a; bi = stat_i[(size_t)i * R + r]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane != 0) return;
  const int64_t w = bi == 0x7fffffff ? 0 : bi;
  tokens[(size_t)r * tok_stride] = w;
  if (tokens2) tokens2[r] = w;
}

static int vocab_maps(CUtensorMap* ta, CUtensorMap* tb, const void* X, int ldx, const void* W, int ldw, int R, int V, int H) {
  MMQG_TRY(make_tmap_bf16_2d(ta, X, R, H, ldx, 128, 64));
  return make_tmap_bf16_2d(tb, W, V, H, ldw, 256, 64);
}

size_t vocab_stat_floats(int R, int V) { return (size_t)ceil_div(V, 256) * R; }

int vocab_nll_fwd(const void* X, int ldx, const void* W, int ldw, const float* bias, const int64_t* targets, const float* row_w, int R,
                  int V, int H, float dscale, float* nll, float* lse, float* row_scale, float* stat_a, float* stat_b, float* tgt_logit,
                  cudaStream_t st) {
  MMQG_REQUIRE(X && W && targets && nll && stat_a && stat_b && tgt_logit && R > 0 && V > 0 && H > 0, "vocab_nll_fwd: bad args");
  CUtensorMap ta, tb;
  MMQG_TRY(vocab_maps(&ta, &tb, X, ldx, W, ldw, R, V, H));
  TcGemmP p{};
  p.M = R; p.N = V; p.nk1 = ceil_div(H, 64); p.nk2 = 0; p.alpha = 1.f; p.split_k = 1;
  p.bias = bias; p.targets = reinterpret_cast<const long long*>(targets);
  p.stat_a = stat_a; p.stat_b = stat_b; p.tgt_logit = tgt_logit;
  {
    const int prev = tl_gemm_class;
    tl_gemm_class = KC_GEMM_SEQ;
    MMQG_PROBE(KC_GEMM_SEQ, 2.0 * R * V * (double)H, 2.0 * ((double)R + V) * H + 8.0 * vocab_stat_floats(R, V));
    const int rc = gemm_tc_vocab_launch(ta, tb, p, VE_STATS, st);
    tl_gemm_class = prev;
    MMQG_TRY(rc);
  }
  MMQG_PROBE(KC_LOSS, 0, 8.0 * vocab_stat_floats(R, V) + 16.0 * R);
  nll_merge_kernel<<<ceil_div(R, 128), 128, 0, st>>>(stat_a, stat_b, tgt_logit, row_w, ceil_div(V, 256), R, dscale, nll, lse, row_scale);
  MMQG_LAUNCH_CHECK();
  return 0;
}

int vocab_dlogits(const void* X, int ldx, const void* W, int ldw, const float* bias, const int64_t* targets, const float* lse,
                  const float* row_scale, int R, int V, int H, void* dlogits, int lddl, cudaStream_t st) {
  MMQG_REQUIRE(X && W && targets && lse && row_scale && dlogits && lddl % 8 == 0 && lddl >= V, "vocab_dlogits: bad args");
  CUtensorMap ta, tb;
  MMQG_TRY(vocab_maps(&ta, &tb, X, ldx, W, ldw, R, V, H));
  TcGemmP p{};
  p.M = R; p.N = V; p.nk1 = ceil_div(H, 64); p.nk2 = 0; p.alpha = 1.f; p.split_k = 1;
  p.C = dlogits; p.ldc = lddl; p.c_bf16 = 1;
  p.bias = bias; p.targets = reinterpret_cast<const long long*>(targets); p.lse = lse; p.row_scale = row_scale;
  MMQG_PROBE(KC_GEMM_SEQ, 2.0 * R * V * (double)H, 2.0 * ((double)R + V) * H + 2.0 * R * V);
  return gemm_tc_vocab_launch(ta, tb, p, VE_DLOGITS, st);
}

int vocab_argmax(const void* X, int ldx, const void* W, int ldw, const float* bias, int R, int V, int H, float* stat_a, int* stat_i,
                 int64_t* tokens, long long tok_stride, int64_t* tokens2, cudaStream_t st) {
  MMQG_REQUIRE(X && W && stat_a && stat_i && tokens && R > 0 && V > 0 && H > 0, "vocab_argmax: bad args");
  CUtensorMap ta, tb;
  MMQG_TRY(vocab_maps(&ta, &tb, X, ldx, W, ldw, R, V, H));
  TcGemmP p{};
  p.M = R; p.N = V; p.nk1 = ceil_div(H, 64); p.nk2 = 0; p.alpha = 1.f; p.split_k = 1;
  p.bias = bias; p.stat_a = stat_a; p.stat_i = stat_i;
  MMQG_PROBE(KC_GEMM_SEQ, 2.0 * R * V * (double)H, 2.0 * ((double)R + V) * H + 8.0 * vocab_stat_floats(R, V));
  MMQG_TRY(gemm_tc_vocab_launch(ta, tb, p, VE_ARGMAX, st));
  argmax_merge_kernel<<<ceil_div(R, 8), 256, 0, st>>>(stat_a, stat_i, ceil_div(V, 256), R, tokens, tok_stride, tokens2);
  MMQG_LAUNCH_CHECK();
  return 0;
}

// rows of one d-logits chunk: MMQG_LH_BYTES (default 32 MB) of bf16, a multiple of 128; never below
// MMQG_LH_MINROWS (default 1024): every chunk re-reads and re-writes all of dW_out (V x H fp32), so short chunks
// at a large vocabulary cost more in that traffic than the L2 residency of the operand returns (measured at
// cfg-4, V = 50k, B200: 256 / 512 / 1024 rows -> 19.0 / 17.9 / 17.2 ms per step).
int vocab_chunk_rows(int R, int Vp) {
  const char* e = getenv("MMQG_LH_BYTES");      // read per call: tests switch the chunking inside one process
  const char* m = getenv("MMQG_LH_MINROWS");
  long long budget = e ? atoll(e) : (32ll << 20);
  if (budget < (1 << 16)) budget = 1 << 16;
  long long min_rows = m ? atoll(m) : 1024;
  if (e && !m) min_rows = 128;                  // an explicit byte budget is honoured down to one m-tile
  long long rc = budget / (2ll * (Vp > 0 ? Vp : 1));
  rc = rc / 128 * 128;
  if (rc < min_rows) rc = min_rows;
  if (rc < 128) rc = 128;
  if (rc > R) rc = R;
  return (int)rc;
}

// split-K factor of dH = dZ (rc x V) . W (V x H): enough slices to fill the SMs, at least 4 k-blocks each
int vocab_dh_split(int rc, int V, int H) {
  const int tiles = ceil_div(rc, 128) * ceil_div(H, 128);
  int s = tiles >= 96 ? 1 : 128 / tiles;
  const int nk = ceil_div(V, 64);
  while (s > 1 && nk / s < 4) --s;
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}

// Backward of the loss head over rows [0, R): per chunk of rows, d logits -> dH (R x H fp32), dW (V x H, += unless
// first chunk and !accumulate) and db (V).  `dl` is the chunk buffer (rc x Vp bf16), `part` the split-K scratch
// (split x rc x H fp32, may be null when vocab_dh_split() == 1).
int vocab_nll_bwd(const void* X, int ldx, const void* W, int ldw, const float* bias, const int64_t* targets, const float* lse,
                  const float* row_scale, int R, int V, int H, void* dl, int Vp, int rc_rows, float* part, float* dH, int lddh, float* dW,
                  float* db, bool accumulate, cudaStream_t st) {
  const bf16* Xb = reinterpret_cast<const bf16*>(X);
  bool first = !accumulate;
  for (int q0 = 0; q0 < R; q0 += rc_rows, first = false) {
    const int rc = R - q0 < rc_rows ? R - q0 : rc_rows;
    MMQG_TRY(vocab_dlogits(Xb + (size_t)q0 * ldx, ldx, W, ldw, bias, targets + q0, lse + q0, row_scale + q0, rc, V, H, dl, Vp, st));
    const int split = part ? vocab_dh_split(rc, V, H) : 1;
    mmqg_gemm_bf16_args a{};
    // dH(rc,H) = dZ(rc,V) . W(V,H): A K-major, B MN-major
    a.A = dl; a.lda = Vp; a.a_mn_major = 0; a.B = W; a.ldb = ldw; a.b_mn_major = 1; a.M = rc; a.N = H; a.K = V;
    a.alpha = 1.f; a.beta = 0.f;
    if (split > 1) {
      a.C = part; a.ldc = H; a.split_k = split; a.c_split_stride = (long long)rc * H;
      MMQG_TRY(gemm_bf16(a, st));
      MMQG_TRY(reduce_partials(part, split, (long long)rc * H, dH + (size_t)q0 * lddh, lddh, rc, H, st));
    } else {
      a.C = dH + (size_t)q0 * lddh; a.ldc = lddh; a.split_k = 1;
      MMQG_TRY(gemm_bf16(a, st));
    }
    // dW(V,H) (+)= dZ^T(V,rc) . X(rc,H): both MN-major
    mmqg_gemm_bf16_args b{};
    b.A = dl; b.lda = Vp; b.a_mn_major = 1; b.B = Xb + (size_t)q0 * ldx; b.ldb = ldx; b.b_mn_major = 1; b.M = V; b.N = H; b.K = rc;
    b.C = dW; b.ldc = H; b.alpha = 1.f; b.split_k = 1;
    if (!first) { b.Cin = dW; b.ldcin = H; b.beta = 1.f; }
    MMQG_TRY(gemm_bf16(b, st));
    MMQG_TRY(colsum_bf16(dl, Vp, db, nullptr, rc, V, first ? 0.f : 1.f, st));
  }
  return 0;
}

}  // namespace mmqg

// ---- C ABI ------------------------------------------------------------------------------------
using namespace mmqg;

namespace {
struct VocabWs {
  float *stat_a, *stat_b, *tgt_logit, *part;
  int* stat_i;
  void* dl;
  int Vp, rc;
  size_t bytes;
};
VocabWs carve_vocab(int R, int V, int H, void* base) {
  VocabWs w{};
  char* b = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t n) { off = align_up(off, 256); char* p = b ? b + off : nullptr; off += n; return p; };
  const size_t ns = vocab_stat_floats(R, V);
  w.Vp = (V + 7) / 8 * 8;
  w.rc = vocab_chunk_rows(R, w.Vp);
  w.stat_a = reinterpret_cast<float*>(take(ns * 4));
  w.stat_b = reinterpret_cast<float*>(take(ns * 4));
  w.stat_i = reinterpret_cast<int*>(w.stat_b);                    // arg-max partials share the sum-exp slot
  w.tgt_logit = reinterpret_cast<float*>(take((size_t)R * 4));
  w.dl = take((size_t)w.rc * w.Vp * 2);
  w.part = reinterpret_cast<float*>(take((size_t)16 * w.rc * H * 4));
  w.bytes = align_up(off, 256);
  return w;
}
}  // namespace

extern "C" {

size_t mmqg_vocab_workspace_bytes(int R, int V, int H) { return carve_vocab(R, V, H, nullptr).bytes; }

int mmqg_vocab_nll_fwd(const void* h_bf16, const void* w_bf16, const float* bias, const int64_t* targets, const float* row_w, int R, int V,
                       int H, float grad_scale, void* workspace, size_t workspace_bytes, float* nll, float* lse, float* row_scale,
                       void* stream) {
  if (H % 8 != 0) return set_err(MMQG_ERR_BAD_ARG, "vocab_nll_fwd: H must be a multiple of 8");
  VocabWs w = carve_vocab(R, V, H, workspace);
  if (!workspace || w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  return vocab_nll_fwd(h_bf16, H, w_bf16, H, bias, targets, row_w, R, V, H, grad_scale, nll, lse, row_scale, w.stat_a, w.stat_b,
                       w.tgt_logit, as_stream(stream));
}

int mmqg_vocab_nll_bwd(const void* h_bf16, const void* w_bf16, const float* bias, const int64_t* targets, const float* lse,
                       const float* row_scale, int R, int V, int H, void* workspace, size_t workspace_bytes, float* dH, float* dW,
                       float* db, int accumulate, void* stream) {
  if (H % 8 != 0) return set_err(MMQG_ERR_BAD_ARG, "vocab_nll_bwd: H must be a multiple of 8");
  VocabWs w = carve_vocab(R, V, H, workspace);
  if (!workspace || w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  return vocab_nll_bwd(h_bf16, H, w_bf16, H, bias, targets, lse, row_scale, R, V, H, w.dl, w.Vp, w.rc, w.part, dH, H, dW, db,
                       accumulate != 0, as_stream(stream));
}

int mmqg_decode_step_argmax(const void* h_bf16, const void* w_bf16, const float* bias, int R, int V, int H, void* workspace,
                            size_t workspace_bytes, int64_t* tokens, long long tok_stride, void* stream) {
  if (H % 8 != 0) return set_err(MMQG_ERR_BAD_ARG, "decode_step_argmax: H must be a multiple of 8");
  VocabWs w = carve_vocab(R, V, H, workspace);
  if (!workspace || w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  return vocab_argmax(h_bf16, H, w_bf16, H, bias, R, V, H, w.stat_a, w.stat_i, tokens, tok_stride, nullptr, as_stream(stream));
}

}  // extern "C"

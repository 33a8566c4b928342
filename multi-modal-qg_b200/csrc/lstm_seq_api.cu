// Sequence-level LSTM layer through the C ABI (SURVEY.md section 8b minimum export set: lstm_seq_fwd / lstm_seq_bwd).
// One call runs ALL T steps of one torch.nn.LSTM layer (reference encoder.py:69,98 and decoder.py:20 as the
// whole-sequence call of encoder.py:95-100 / decoder.py:25-34) on the tensor-core path the train engine uses:
//   forward : x -> bf16, hoisted input projection X W_ih^T + (b_ih + b_hh) as ONE tcgen05 GEMM over T*B rows,
//             then the persistent recurrent kernel (lstm_persist.cu: W_hh resident in shared memory, cell state
//             in registers, one launch for all steps);
//   backward: persistent BPTT kernel (split-K 4-CTA clusters), then the hoisted products dX = dG W_ih,
//             dW_ih = dG^T X, dW_hh = dG^T H_prev, db = column sums of dG, dh0 = dG_0 W_hh.
// Activations saved for the backward pass live in the caller's workspace (the caller keeps it alive between the
// two calls); parameters and results are fp32 in PyTorch layout, products run in bf16 with fp32 accumulation.
#include <cuda_bf16.h>
#include "kernels.h"

namespace mmqg {
namespace {

typedef uint16_t b16;

struct SeqWs {
  b16 *x16, *w_ih16, *wp_f, *wp_b, *hs, *dg;
  float *bsum, *acts, *cs, *dh_last, *dc_tmp, *dw_part;
  uint32_t* flags;
  int Ip;
  size_t bytes;
};

SeqWs carve_seq(int T, int B, int I, int H, void* base) {
  SeqWs w{};
  char* b = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t n) { off = align_up(off, 256); char* p = b ? b + off : nullptr; off += n; return p; };
  const size_t G = 4 * (size_t)H, R = (size_t)T * B;
  w.Ip = (I + 7) / 8 * 8;
  w.x16 = reinterpret_cast<b16*>(take(R * w.Ip * 2));
  w.w_ih16 = reinterpret_cast<b16*>(take(G * w.Ip * 2));
  w.wp_f = reinterpret_cast<b16*>(take(G * H * 2));
  w.wp_b = reinterpret_cast<b16*>(take(G * H * 2));
  w.hs = reinterpret_cast<b16*>(take((R + B) * H * 2));
  w.dg = reinterpret_cast<b16*>(take(R * G * 2));
  w.bsum = reinterpret_cast<float*>(take(G * 4));
  w.acts = reinterpret_cast<float*>(take(R * G * 4));
  w.cs = reinterpret_cast<float*>(take((R + B) * H * 4));
  w.dh_last = reinterpret_cast<float*>(take((size_t)B * H * 4));
  w.dc_tmp = reinterpret_cast<float*>(take((size_t)B * H * 4));
  w.flags = reinterpret_cast<uint32_t*>(take((size_t)(T + 1) * ceil_div(B, 128) * 4));
  w.bytes = align_up(off, 256);
  return w;
}

__global__ void cvt_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}
__global__ void copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src ? src[i] : 0.f;
}

int check_seq(int T, int B, int I, int H) {
  MMQG_REQUIRE(T >= 1 && B >= 1 && I >= 1 && H >= 1, "lstm_seq: bad sizes T=%d B=%d I=%d H=%d", T, B, I, H);
  MMQG_REQUIRE(lstm_persist_ok(B, H),
               "lstm_seq: the persistent recurrent kernels need H a multiple of 64, H <= 512 and (H/16)*ceil(B/128) CTAs "
               "co-resident on the device; got B=%d H=%d (use the per-step building blocks for other shapes)", B, H);
  return 0;
}

}  // namespace
}  // namespace mmqg

using namespace mmqg;

extern "C" {

size_t mmqg_lstm_seq_workspace_bytes(int T, int B, int I, int H) { return carve_seq(T, B, I, H, nullptr).bytes; }

int mmqg_lstm_seq_fwd(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* h0,
                      const float* c0, int T, int B, int I, int H, void* workspace, size_t workspace_bytes, float* y, float* hn,
                      float* cn, void* stream) {
  MMQG_TRY(check_seq(T, B, I, H));
  MMQG_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && y && workspace, "lstm_seq_fwd: null pointer");
  SeqWs w = carve_seq(T, B, I, H, workspace);
  if (w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  cudaStream_t st = as_stream(stream);
  const int G = 4 * H;
  const long long R = (long long)T * B, n = (long long)B * H;
  MMQG_TRY(cvt_f32_bf16_2d(x, I, w.x16, w.Ip, R, I, w.Ip, st));
  MMQG_TRY(cvt_f32_bf16_2d(w_ih, I, w.w_ih16, w.Ip, G, I, w.Ip, st));
  MMQG_TRY(add2(b_ih, b_hh, w.bsum, G, st));
  MMQG_TRY(pack_whh(w_hh, w.wp_f, w.wp_b, H, st));
  mmqg_gemm_bf16_args a{};
  a.A = w.x16; a.lda = w.Ip; a.B = w.w_ih16; a.ldb = w.Ip; a.K = w.Ip; a.M = (int)R; a.N = G; a.C = w.acts; a.ldc = G;
  a.alpha = 1.f; a.bias = w.bsum; a.split_k = 1;
  MMQG_TRY(gemm_bf16(a, st));
  // slab 0 of the state sequences = initial state
  if (h0) MMQG_TRY(cvt_f32_bf16_2d(h0, H, w.hs, H, B, H, H, st));
  else MMQG_CUDA(cudaMemsetAsync(w.hs, 0, sizeof(b16) * (size_t)n, st));
  copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c0, w.cs, n);
  MMQG_LAUNCH_CHECK();
  MMQG_TRY(lstm_seq_fwd_persist(w.acts, w.cs, w.hs, w.wp_f, nullptr, nullptr, 0, w.flags, T, B, H, 1, st));
  cvt_bf16_f32_kernel<<<(unsigned)((R * H + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(w.hs) + n, y, R * H);
  MMQG_LAUNCH_CHECK();
  if (hn) MMQG_CUDA(cudaMemcpyAsync(hn, y + (size_t)(T - 1) * n, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  if (cn) MMQG_CUDA(cudaMemcpyAsync(cn, w.cs + (size_t)T * n, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int mmqg_lstm_seq_bwd(const float* dy, const float* dhn, const float* dcn, const float* w_hh, int T, int B, int I, int H,
                      void* workspace, size_t workspace_bytes, float* dx, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                      float* dh0, float* dc0, void* stream) {
  MMQG_TRY(check_seq(T, B, I, H));
  MMQG_REQUIRE(workspace && dw_ih && dw_hh && db_ih && w_hh, "lstm_seq_bwd: null pointer");
  SeqWs w = carve_seq(T, B, I, H, workspace);
  if (w.bytes > workspace_bytes) return set_err(MMQG_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, w.bytes);
  cudaStream_t st = as_stream(stream);
  const int G = 4 * H;
  const long long R = (long long)T * B, n = (long long)B * H;
  // the last step's external gradient is dy[T-1] + dhn: the kernel adds dh_ext(t) for every t and dh_last at t = T-1
  copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dhn, w.dh_last, n);
  MMQG_LAUNCH_CHECK();
  copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dcn, w.dc_tmp, n);
  MMQG_LAUNCH_CHECK();
  MMQG_TRY(lstm_seq_bwd_persist(w.acts, w.cs, w.dg, w.wp_b, dy, n, H, w.dh_last, w.dc_tmp, w.flags, T, B, H, 0, dc0 ? dc0 : w.dc_tmp, st));
  mmqg_gemm_bf16_args a{};
  if (dx) {       // dX (R,I) = dG (R,4H) . W_ih (4H,I)
    a = mmqg_gemm_bf16_args{};
    a.A = w.dg; a.lda = G; a.B = w.w_ih16; a.ldb = w.Ip; a.b_mn_major = 1; a.M = (int)R; a.N = I; a.K = G; a.C = dx; a.ldc = I;
    a.alpha = 1.f; a.split_k = 1;
    MMQG_TRY(gemm_bf16(a, st));
  }
  // dW_ih (4H,I) = dG^T X ;  dW_hh (4H,H) = dG^T H_prev (slabs 0 .. T-1 of the state sequence)
  a = mmqg_gemm_bf16_args{};
  a.A = w.dg; a.lda = G; a.a_mn_major = 1; a.B = w.x16; a.ldb = w.Ip; a.b_mn_major = 1; a.M = G; a.N = I; a.K = (int)R; a.C = dw_ih; a.ldc = I;
  a.alpha = 1.f; a.split_k = 1;
  MMQG_TRY(gemm_bf16(a, st));
  a = mmqg_gemm_bf16_args{};
  a.A = w.dg; a.lda = G; a.a_mn_major = 1; a.B = w.hs; a.ldb = H; a.b_mn_major = 1; a.M = G; a.N = H; a.K = (int)R; a.C = dw_hh; a.ldc = H;
  a.alpha = 1.f; a.split_k = 1;
  MMQG_TRY(gemm_bf16(a, st));
  MMQG_TRY(colsum_bf16(w.dg, G, db_ih, nullptr, (int)R, G, 0.f, st));       // one reduction, copied: both biases get the SAME gradient
  if (db_hh) MMQG_CUDA(cudaMemcpyAsync(db_hh, db_ih, sizeof(float) * (size_t)G, cudaMemcpyDeviceToDevice, st));
  if (dh0) {      // dh0 (B,H) = dG_0 (B,4H) . W_hh (4H,H): wp_b holds W_hh^T (H,4H) = the (N,K) K-major operand
    a = mmqg_gemm_bf16_args{};
    a.A = w.dg; a.lda = G; a.B = w.wp_b; a.ldb = G; a.M = B; a.N = H; a.K = G; a.C = dh0; a.ldc = H; a.alpha = 1.f; a.split_k = 1;
    MMQG_TRY(gemm_bf16(a, st));
  }
  return 0;
}

}  // extern "C"

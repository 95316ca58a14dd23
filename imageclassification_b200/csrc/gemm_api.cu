// gemm_api.cu — extern "C" entry points of the MLP GEMMs: bf16 -> tcgen05 kernels (gemm_tc.cu),
// fp32 (or CNX_GEMM_FORCE_SIMT) -> CUDA-core kernels (gemm_simt.cu).  There is no CPU path.
#include "common.cuh"
#include "epilogue.cuh"

namespace cnx {
template <typename TIN, typename TOUT, int KIND>
int gemm_tn_simt(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s);
template <int KIND, typename TOUT>
int gemm_tn_tc(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s);
template <typename TIN>
int gemm_wgrad_simt(const void* X, const void* Y, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                    float* colsum_x, void* workspace, int64_t workspace_bytes, cudaStream_t s);
int gemm_wgrad_tc(const void* X, const void* Y, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                  float* colsum_x, void* workspace, int64_t workspace_bytes, cudaStream_t s, int64_t ldx = 0, int64_t ldy = 0,
                  int passes = 1);
int mlp_fused_fwd_x3_tc(const void* xn2, const void* W1x3, const float* b1, const void* W2x3, const float* b2, const float* gamma,
                        const float* dp, int64_t rows_per_sample, const float* shortcut, float* out, int64_t M, int64_t C,
                        cudaStream_t s);
int mlp_fused_fwd_tc(const void* xn, const void* W1, const float* b1, const void* W2, const float* b2, const float* gamma,
                     const float* dp, int64_t rows_per_sample, const void* shortcut, void* out, int64_t M, int64_t C,
                     cudaStream_t s);
int gemm_dgelu_recompute_tc(const void* dz, const void* Bt, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s);
int64_t wgrad_workspace_bytes_simt(int64_t M, int64_t N1, int64_t N2);
int64_t wgrad_workspace_bytes_tc(int64_t M, int64_t N1, int64_t N2);
}  // namespace cnx

using namespace cnx;

#define CNX_GEMM_ARGS_OK(name)                                                                        \
  CNX_REQUIRE(M > 0 && N > 0 && K > 0, CNX_E_BADARG, name ": bad shape");                             \
  CNX_REQUIRE(dtype_ok(dtype), CNX_E_BADARG, name ": bad dtype");                                     \
  CNX_REQUIRE(N % 8 == 0, CNX_E_SHAPE, name ": N=%lld must be a multiple of 8", (long long)N)

extern "C" {

int cnx_gemm_bias_gelu_fwd(const void* A, const void* W1, const float* b1, int64_t M, int64_t N, int64_t K,
                           void* gprime_out, void* g_out, int dtype, int flags, void* stream) {
  CNX_REQUIRE(A && W1 && b1 && g_out, CNX_E_BADARG, "gemm_bias_gelu_fwd: null pointer");
  CNX_GEMM_ARGS_OK("gemm_bias_gelu_fwd");
  EpiParams ep = {b1, nullptr, nullptr, 1, nullptr, gprime_out, g_out, N};
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CNX_F32) return gemm_tn_simt<float, float, EPI_BIAS_GELU>(A, W1, M, N, K, ep, s);
  if (flags & CNX_GEMM_FORCE_SIMT) return gemm_tn_simt<bf16, bf16, EPI_BIAS_GELU>(A, W1, M, N, K, ep, s);
  return gemm_tn_tc<EPI_BIAS_GELU, bf16>(A, W1, M, N, K, ep, s);
}

int cnx_gemm_bias_gelu_fwd_x3(const void* A3, const void* W3, const float* b1, int64_t M, int64_t N, int64_t K3, void* g3,
                              int a_segments, void* stream) {
  CNX_REQUIRE(A3 && W3 && b1 && g3, CNX_E_BADARG, "gemm_bias_gelu_fwd_x3: null pointer");
  CNX_REQUIRE(M > 0 && N > 0 && K3 > 0 && K3 % 24 == 0, CNX_E_BADARG, "gemm_bias_gelu_fwd_x3: bad shape (K3 = 3K, K %% 8 == 0)");
  CNX_REQUIRE(N % 32 == 0, CNX_E_SHAPE, "gemm_bias_gelu_fwd_x3: N=%lld must be a multiple of 32", (long long)N);
  CNX_REQUIRE(a_segments == 3 || (a_segments == 2 && (K3 / 3) % 32 == 0), CNX_E_SHAPE,
              "gemm_bias_gelu_fwd_x3: a_segments must be 3, or 2 with K3/3 a multiple of 32 (K3=%lld)", (long long)K3);
  EpiParams ep = {b1, nullptr, nullptr, 1, nullptr, nullptr, g3, N};
  if (a_segments == 2) ep.a_wrap = (int32_t)(2 * (K3 / 3));      // the K loop's third segment re-reads A's hi columns
  return gemm_tn_tc<EPI_BIAS_GELU3, bf16>(A3, W3, M, N, K3, ep, (cudaStream_t)stream);
}

int cnx_gemm_bias_gelu_fwd_x3_train(const void* A3, const void* W3, const float* b1, int64_t M, int64_t N, int64_t K3, void* g3,
                                    float* gprime, int a_segments, void* stream) {
  CNX_REQUIRE(A3 && W3 && b1 && g3 && gprime, CNX_E_BADARG, "gemm_bias_gelu_fwd_x3_train: null pointer");
  CNX_REQUIRE(M > 0 && N > 0 && K3 > 0 && K3 % 24 == 0, CNX_E_BADARG, "gemm_bias_gelu_fwd_x3_train: bad shape (K3 = 3K, K %% 8 == 0)");
  CNX_REQUIRE(N % 32 == 0, CNX_E_SHAPE, "gemm_bias_gelu_fwd_x3_train: N=%lld must be a multiple of 32", (long long)N);
  CNX_REQUIRE(a_segments == 3 || (a_segments == 2 && (K3 / 3) % 32 == 0), CNX_E_SHAPE,
              "gemm_bias_gelu_fwd_x3_train: a_segments must be 3, or 2 with K3/3 a multiple of 32 (K3=%lld)", (long long)K3);
  EpiParams ep = {b1, nullptr, nullptr, 1, nullptr, gprime, g3, N};
  if (a_segments == 2) ep.a_wrap = (int32_t)(2 * (K3 / 3));
  return gemm_tn_tc<EPI_BIAS_GELU3, bf16>(A3, W3, M, N, K3, ep, (cudaStream_t)stream);
}

int cnx_gemm_bias_scale_residual_fwd(const void* A, const void* W2, const float* b2, const float* gamma,
                                     const float* dp, int64_t rows_per_sample, const void* shortcut, void* out,
                                     int stream_dtype, int64_t M, int64_t N, int64_t K, int dtype, int flags,
                                     void* stream) {
  CNX_REQUIRE(A && W2 && out, CNX_E_BADARG, "gemm_bias_scale_residual_fwd: null pointer");
  CNX_GEMM_ARGS_OK("gemm_bias_scale_residual_fwd");
  CNX_REQUIRE(dtype_ok(stream_dtype) && rows_per_sample > 0, CNX_E_BADARG, "gemm_bias_scale_residual_fwd: bad argument");
  EpiParams ep = {b2, gamma, dp, rows_per_sample, shortcut, out, nullptr, N};
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & CNX_GEMM_A_SPLIT2) {
    // A is the [hi | mid] split operand written by cnx_gemm_bias_gelu_fwd_x3 (2K/3 columns); K counts all three segments
    CNX_REQUIRE(dtype == CNX_BF16 && !(flags & CNX_GEMM_FORCE_SIMT) && K % 3 == 0 && (K / 3) % 64 == 0, CNX_E_SHAPE,
                "gemm_bias_scale_residual_fwd: CNX_GEMM_A_SPLIT2 needs bf16 operands and K/3 a multiple of 64 (K=%lld)", (long long)K);
    ep.a_wrap = (int32_t)(2 * (K / 3));
  }
  if (dtype == CNX_F32) {
    CNX_REQUIRE(stream_dtype == CNX_F32, CNX_E_BADARG, "fp32 GEMM needs an fp32 residual stream");
    return gemm_tn_simt<float, float, EPI_SCALE_RES>(A, W2, M, N, K, ep, s);
  }
  if (flags & CNX_GEMM_FORCE_SIMT) {
    if (stream_dtype == CNX_F32) return gemm_tn_simt<bf16, float, EPI_SCALE_RES>(A, W2, M, N, K, ep, s);
    return gemm_tn_simt<bf16, bf16, EPI_SCALE_RES>(A, W2, M, N, K, ep, s);
  }
  if (stream_dtype == CNX_F32) return gemm_tn_tc<EPI_SCALE_RES, float>(A, W2, M, N, K, ep, s);
  return gemm_tn_tc<EPI_SCALE_RES, bf16>(A, W2, M, N, K, ep, s);
}

int cnx_mlp_fused_fwd(const void* xn, const void* W1, const float* b1, const void* W2, const float* b2, const float* gamma,
                      const float* dp, int64_t rows_per_sample, const void* shortcut, void* out, int64_t M, int64_t C,
                      void* stream) {
  CNX_REQUIRE(xn && W1 && b1 && W2 && b2 && shortcut && out, CNX_E_BADARG, "mlp_fused_fwd: null pointer");
  CNX_REQUIRE(M > 0 && rows_per_sample > 0, CNX_E_BADARG, "mlp_fused_fwd: bad shape");
  return mlp_fused_fwd_tc(xn, W1, b1, W2, b2, gamma, dp, rows_per_sample, shortcut, out, M, C, (cudaStream_t)stream);
}

int cnx_mlp_fused_fwd_x3(const void* xn2, const void* W1x3, const float* b1, const void* W2x3, const float* b2, const float* gamma,
                         const float* dp, int64_t rows_per_sample, const float* shortcut, float* out, int64_t M, int64_t C,
                         void* stream) {
  CNX_REQUIRE(xn2 && W1x3 && b1 && W2x3 && b2 && shortcut && out, CNX_E_BADARG, "mlp_fused_fwd_x3: null pointer");
  CNX_REQUIRE(M > 0 && rows_per_sample > 0, CNX_E_BADARG, "mlp_fused_fwd_x3: bad shape");
  CNX_REQUIRE(((uintptr_t)shortcut & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)b2 & 15) == 0 &&
                  (gamma == nullptr || ((uintptr_t)gamma & 15) == 0) && ((uintptr_t)b1 & 15) == 0,
              CNX_E_SHAPE, "mlp_fused_fwd_x3: shortcut, out, biases and gamma must be 16-byte aligned");
  return mlp_fused_fwd_x3_tc(xn2, W1x3, b1, W2x3, b2, gamma, dp, rows_per_sample, shortcut, out, M, C, (cudaStream_t)stream);
}

int cnx_gemm_dgrad_gelu_bwd(const void* dz, const void* Bt, const void* gprime, void* dh, int64_t M, int64_t N,
                            int64_t K, int dtype, int flags, void* stream) {
  CNX_REQUIRE(dz && Bt && gprime && dh, CNX_E_BADARG, "gemm_dgrad_gelu_bwd: null pointer");
  CNX_GEMM_ARGS_OK("gemm_dgrad_gelu_bwd");
  EpiParams ep = {nullptr, nullptr, nullptr, 1, gprime, dh, nullptr, N};
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CNX_F32) return gemm_tn_simt<float, float, EPI_DGELU>(dz, Bt, M, N, K, ep, s);
  if (flags & CNX_GEMM_FORCE_SIMT) return gemm_tn_simt<bf16, bf16, EPI_DGELU>(dz, Bt, M, N, K, ep, s);
  return gemm_tn_tc<EPI_DGELU, bf16>(dz, Bt, M, N, K, ep, s);
}

int cnx_gemm_dgrad_gelu_recompute_bwd(const void* dz, const void* Bt, const void* xn, const void* W1, const float* b1, void* dh,
                                      int64_t M, int64_t N, int64_t K, int dtype, void* stream) {
  CNX_REQUIRE(dz && Bt && xn && W1 && b1 && dh, CNX_E_BADARG, "gemm_dgrad_gelu_recompute_bwd: null pointer");
  CNX_REQUIRE(M > 0 && N > 0 && K > 0, CNX_E_BADARG, "gemm_dgrad_gelu_recompute_bwd: bad shape");
  CNX_REQUIRE(dtype == CNX_BF16, CNX_E_BADARG, "gemm_dgrad_gelu_recompute_bwd: bf16 operands only (tcgen05 kernel)");
  EpiParams ep = {b1, nullptr, nullptr, 1, nullptr, dh, nullptr, N};
  ep.a2 = xn;
  ep.b2 = W1;
  return gemm_dgelu_recompute_tc(dz, Bt, M, N, K, ep, (cudaStream_t)stream);
}

int cnx_gemm_dgrad_gelu_bwd_x3(const void* dz2, const void* Bt3, const float* gprime, void* dh2, int64_t M, int64_t N, int64_t K3,
                               int a_segments, void* stream) {
  CNX_REQUIRE(dz2 && Bt3 && gprime && dh2, CNX_E_BADARG, "gemm_dgrad_gelu_bwd_x3: null pointer");
  CNX_REQUIRE(M > 0 && N > 0 && K3 > 0 && K3 % 24 == 0, CNX_E_BADARG, "gemm_dgrad_gelu_bwd_x3: bad shape (K3 = 3K, K %% 8 == 0)");
  CNX_REQUIRE(N % 32 == 0, CNX_E_SHAPE, "gemm_dgrad_gelu_bwd_x3: N=%lld must be a multiple of 32", (long long)N);
  CNX_REQUIRE(a_segments == 3 || (a_segments == 2 && (K3 / 3) % 32 == 0), CNX_E_SHAPE,
              "gemm_dgrad_gelu_bwd_x3: a_segments must be 3, or 2 with K3/3 a multiple of 32 (K3=%lld)", (long long)K3);
  EpiParams ep = {nullptr, nullptr, nullptr, 1, gprime, dh2, nullptr, N};
  if (a_segments == 2) ep.a_wrap = (int32_t)(2 * (K3 / 3));
  return gemm_tn_tc<EPI_DGELU3, bf16>(dz2, Bt3, M, N, K3, ep, (cudaStream_t)stream);
}

int cnx_gemm_plain(const void* A, const void* B, const float* bias, void* out, int out_dtype, int64_t M,
                   int64_t N, int64_t K, int dtype, int flags, void* stream) {
  CNX_REQUIRE(A && B && out, CNX_E_BADARG, "gemm_plain: null pointer");
  CNX_GEMM_ARGS_OK("gemm_plain");
  CNX_REQUIRE(dtype_ok(out_dtype), CNX_E_BADARG, "gemm_plain: bad out dtype");
  EpiParams ep = {bias, nullptr, nullptr, 1, nullptr, out, nullptr, N};
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & CNX_GEMM_A_SPLIT2) {
    CNX_REQUIRE(dtype == CNX_BF16 && !(flags & CNX_GEMM_FORCE_SIMT) && K % 3 == 0 && (K / 3) % 32 == 0, CNX_E_SHAPE,
                "gemm_plain: CNX_GEMM_A_SPLIT2 needs bf16 operands and K/3 a multiple of 32 (K=%lld)", (long long)K);
    ep.a_wrap = (int32_t)(2 * (K / 3));
  }
  if (flags & CNX_GEMM_OUT_ROUND_BF16) {
    CNX_REQUIRE(dtype == CNX_BF16 && out_dtype == CNX_F32, CNX_E_BADARG,
                "gemm_plain: CNX_GEMM_OUT_ROUND_BF16 is for bf16 operands with an fp32 output");
    ep.round_bf16 = 1;
  }
  if (dtype == CNX_F32) {
    CNX_REQUIRE(out_dtype == CNX_F32, CNX_E_BADARG, "gemm_plain: fp32 operands need an fp32 output");
    return gemm_tn_simt<float, float, EPI_PLAIN>(A, B, M, N, K, ep, s);
  }
  if (flags & CNX_GEMM_FORCE_SIMT) {
    if (out_dtype == CNX_F32) return gemm_tn_simt<bf16, float, EPI_PLAIN>(A, B, M, N, K, ep, s);
    return gemm_tn_simt<bf16, bf16, EPI_PLAIN>(A, B, M, N, K, ep, s);
  }
  if (out_dtype == CNX_F32) return gemm_tn_tc<EPI_PLAIN, float>(A, B, M, N, K, ep, s);
  return gemm_tn_tc<EPI_PLAIN, bf16>(A, B, M, N, K, ep, s);
}

int64_t cnx_gemm_wgrad_workspace_bytes(int64_t M, int64_t N1, int64_t N2, int dtype, int flags) {
  if (M <= 0 || N1 <= 0 || N2 <= 0) return 0;
  if (dtype == CNX_F32 || (flags & CNX_GEMM_FORCE_SIMT)) return wgrad_workspace_bytes_simt(M, N1, N2);
  return wgrad_workspace_bytes_tc(M, N1, N2);
}

int cnx_gemm_wgrad(const void* X, const void* Y, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                   float* colsum_x, void* workspace, int64_t workspace_bytes, int dtype, int flags, void* stream) {
  CNX_REQUIRE(X && Y && out && workspace, CNX_E_BADARG, "gemm_wgrad: null pointer");
  CNX_REQUIRE(M > 0 && N1 > 0 && N2 > 0 && dtype_ok(dtype), CNX_E_BADARG, "gemm_wgrad: bad shape/dtype");
  CNX_REQUIRE(N1 % 8 == 0 && N2 % 8 == 0, CNX_E_SHAPE, "gemm_wgrad: N1, N2 must be multiples of 8");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CNX_F32) return gemm_wgrad_simt<float>(X, Y, M, N1, N2, accumulate, out, colsum_x, workspace, workspace_bytes, s);
  if (flags & CNX_GEMM_FORCE_SIMT) return gemm_wgrad_simt<bf16>(X, Y, M, N1, N2, accumulate, out, colsum_x, workspace, workspace_bytes, s);
  return gemm_wgrad_tc(X, Y, M, N1, N2, accumulate, out, colsum_x, workspace, workspace_bytes, s);
}

/* fp32-accurate weight gradient on the tensor cores: X2 [M, 2*N1] = [hi | mid] and Y2 [M, 2*N2] = [hi | mid] (cnx_split3 with
 * segments = 2, or the split outputs of the x3 kernels); out = Xhi^T.Yhi + Xmid^T.Yhi + Xhi^T.Ymid (three bf16 wgrad GEMMs over
 * column blocks of the split tensors, accumulated in fp32), colsum_x = column sums of Xhi + Xmid. */
int cnx_gemm_wgrad_x3_one_loop(const void* X2, const void* Y2, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                               float* colsum_x, void* workspace, int64_t workspace_bytes, void* stream) {
  CNX_REQUIRE(X2 && Y2 && out && workspace, CNX_E_BADARG, "gemm_wgrad_x3_one_loop: null pointer");
  CNX_REQUIRE(M > 0 && N1 > 0 && N2 > 0, CNX_E_BADARG, "gemm_wgrad_x3_one_loop: bad shape");
  CNX_REQUIRE(N1 % 8 == 0 && N2 % 8 == 0, CNX_E_SHAPE, "gemm_wgrad_x3_one_loop: N1, N2 must be multiples of 8");
  // ONE launch whose K loop walks the three products: one set of split-K partials, one reduction — and accumulation chains in
  // TMEM three times as long as cnx_gemm_wgrad_x3's (see include/cnx.h for what that costs in accuracy)
  return gemm_wgrad_tc(X2, Y2, M, N1, N2, accumulate, out, colsum_x, workspace, workspace_bytes, (cudaStream_t)stream, 2 * N1, 2 * N2, 3);
}

int cnx_gemm_wgrad_x3(const void* X2, const void* Y2, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                      float* colsum_x, void* workspace, int64_t workspace_bytes, void* stream) {
  CNX_REQUIRE(X2 && Y2 && out && workspace, CNX_E_BADARG, "gemm_wgrad_x3: null pointer");
  CNX_REQUIRE(M > 0 && N1 > 0 && N2 > 0, CNX_E_BADARG, "gemm_wgrad_x3: bad shape");
  CNX_REQUIRE(N1 % 8 == 0 && N2 % 8 == 0, CNX_E_SHAPE, "gemm_wgrad_x3: N1, N2 must be multiples of 8");
  cudaStream_t s = (cudaStream_t)stream;
  const bf16* Xh = (const bf16*)X2;
  const bf16* Xm = Xh + N1;
  const bf16* Yh = (const bf16*)Y2;
  const bf16* Ym = Yh + N2;
  if (int rc = gemm_wgrad_tc(Xh, Yh, M, N1, N2, accumulate, out, colsum_x, workspace, workspace_bytes, s, 2 * N1, 2 * N2)) return rc;
  if (int rc = gemm_wgrad_tc(Xm, Yh, M, N1, N2, 1, out, colsum_x, workspace, workspace_bytes, s, 2 * N1, 2 * N2)) return rc;
  return gemm_wgrad_tc(Xh, Ym, M, N1, N2, 1, out, nullptr, workspace, workspace_bytes, s, 2 * N1, 2 * N2);
}

}  // extern "C"

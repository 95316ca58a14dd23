// dwconv.cu — a1/a2: C-ABI entry points of the depthwise 7x7 conv (+fused channels-last LayerNorm) forward, backward-data
// (+residual-gradient add) and backward-weights, channels-last, plus the two small helper kernels (tap-major weight
// transpose, wgrad partial reduction).  The conv kernels themselves are in dwconv2.cu (design notes: dwconv_common.cuh):
//   * halo tiles arrive by TMA (cp.async.bulk.tensor.4d over a {C, W, H, N} tensor map); out-of-bounds box elements are
//     zero-filled by the TMA unit, which IS the conv's padding=3 — no per-element address math, no bounds branches;
//   * producer-warp mbarrier ring, persistent CTAs, a half-warp owns 32 channels as 16 channel PAIRS: every FMA is the
//     packed fma.rn.f32x2; the 49 tap pairs and a CPW x TH output strip live in registers.
// (The first-generation kernels that used to live here were removed in round 2: nothing dispatched to them any more.)
#include "common.cuh"

namespace cnx {

__global__ void __launch_bounds__(256) dwconv7_wgrad_finalize_kernel(const float* __restrict__ partial, int P, int64_t C,
                                                                     int accumulate, float* __restrict__ dw,
                                                                     float* __restrict__ db) {
  // one thread per (tap t, channel c), reading partial[p][t][c] coalesced over c
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  pdl_wait();
  if (i >= 50 * C) return;
  int64_t t = i / C, c = i - t * C;
  float s0 = 0.f, s1 = 0.f;
  int p = 0;
  for (; p + 2 <= P; p += 2) {
    s0 += partial[((int64_t)p * 50 + t) * C + c];
    s1 += partial[((int64_t)(p + 1) * 50 + t) * C + c];
  }
  if (p < P) s0 += partial[((int64_t)p * 50 + t) * C + c];
  float s = s0 + s1;
  if (t < 49) {
    float* o = dw + c * 49 + t;
    *o = accumulate ? *o + s : s;
  } else {
    float* o = db + c;
    *o = accumulate ? *o + s : s;
  }
}

// w [C][49] -> wt [49][C]
__global__ void __launch_bounds__(256) weight_transpose_kernel(const float* __restrict__ w, int64_t C, float* __restrict__ wt) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= 49 * C) return;
  int64_t t = i / C, c = i - t * C;
  wt[i] = w[c * 49 + t];
}

// dwconv2.cu
int dwconv7_ln_fwd_v2(const void* x, int x_dtype, const float* wt, const float* bias, const float* ln_w, const float* ln_b,
                      float eps, int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn, int act_dtype, float* mean,
                      float* rstd, cudaStream_t s);
int dwconv7_dgrad_v2(const void* dy, int dy_dtype, const float* wt, const void* dres, void* dx, int stream_dtype, int64_t N,
                     int64_t H, int64_t W, int64_t C, void* dz_up, const float* dp_up, cudaStream_t s);
int dwconv7_ln_fwd_x3_v2(const void* x, const float* wt, const float* bias, const float* ln_w, const float* ln_b, float eps,
                         int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn3, int segments, float* mean, float* rstd,
                         cudaStream_t s);
int dwconv7_wgrad_v2(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W, int64_t C,
                     float* partial, int P, cudaStream_t s);
}  // namespace cnx

using namespace cnx;

extern "C" {

int cnx_dwconv7_weight_prep(const float* w, int64_t C, float* wt, void* stream) {
  CNX_REQUIRE(w && wt && C > 0, CNX_E_BADARG, "dwconv7_weight_prep: bad argument");
  weight_transpose_kernel<<<(unsigned)((49 * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, C, wt);
  return check_launch("dwconv7_weight_prep");
}

int cnx_dwconv7_ln_fwd(const void* x, int x_dtype, const float* wt, const float* bias, const float* ln_w,
                       const float* ln_b, float eps, int64_t N, int64_t H, int64_t W, int64_t C, void* y,
                       void* xn, int act_dtype, float* mean, float* rstd, void* stream) {
  CNX_REQUIRE(x && wt && bias && ln_w && ln_b && y && xn && mean && rstd, CNX_E_BADARG, "dwconv7_ln_fwd: null pointer");
  CNX_REQUIRE(dtype_ok(x_dtype) && dtype_ok(act_dtype), CNX_E_BADARG, "dwconv7_ln_fwd: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_ln_fwd: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_ln_fwd: C=%lld must be a multiple of 32", (long long)C);
  return dwconv7_ln_fwd_v2(x, x_dtype, wt, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, act_dtype, mean, rstd, (cudaStream_t)stream);
}

int cnx_dwconv7_ln_fwd_x3(const float* x, const float* wt, const float* bias, const float* ln_w, const float* ln_b, float eps,
                          int64_t N, int64_t H, int64_t W, int64_t C, float* y_scratch, void* xn3, float* mean, float* rstd,
                          int segments, void* stream) {
  CNX_REQUIRE(x && wt && bias && ln_w && ln_b && y_scratch && xn3 && mean && rstd, CNX_E_BADARG, "dwconv7_ln_fwd_x3: null pointer");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_ln_fwd_x3: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_ln_fwd_x3: C=%lld must be a multiple of 32", (long long)C);
  CNX_REQUIRE(segments == 2 || segments == 3, CNX_E_BADARG, "dwconv7_ln_fwd_x3: segments must be 2 ([hi | mid]) or 3 ([hi | mid | hi])");
  return dwconv7_ln_fwd_x3_v2(x, wt, bias, ln_w, ln_b, eps, N, H, W, C, y_scratch, xn3, segments, mean, rstd, (cudaStream_t)stream);
}

int cnx_dwconv7_dgrad(const void* dy, int dy_dtype, const float* wt, const void* dres, void* dx, int stream_dtype,
                      int64_t N, int64_t H, int64_t W, int64_t C, void* stream) {
  CNX_REQUIRE(dy && wt && dx, CNX_E_BADARG, "dwconv7_dgrad: null pointer");
  CNX_REQUIRE(dtype_ok(dy_dtype) && dtype_ok(stream_dtype), CNX_E_BADARG, "dwconv7_dgrad: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_dgrad: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_dgrad: C=%lld must be a multiple of 32", (long long)C);
  return dwconv7_dgrad_v2(dy, dy_dtype, wt, dres, dx, stream_dtype, N, H, W, C, nullptr, nullptr, (cudaStream_t)stream);
}

int cnx_dwconv7_dgrad_dz(const void* dy, const float* wt, const void* dres, void* dx, int stream_dtype, int64_t N, int64_t H,
                         int64_t W, int64_t C, void* dz_up, const float* dp_up, void* stream) {
  CNX_REQUIRE(dy && wt && dx && dz_up, CNX_E_BADARG, "dwconv7_dgrad_dz: null pointer");
  CNX_REQUIRE(dtype_ok(stream_dtype), CNX_E_BADARG, "dwconv7_dgrad_dz: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_dgrad_dz: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_dgrad_dz: C=%lld must be a multiple of 32", (long long)C);
  return dwconv7_dgrad_v2(dy, CNX_BF16, wt, dres, dx, stream_dtype, N, H, W, C, dz_up, dp_up, (cudaStream_t)stream);
}

int cnx_dwconv7_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W,
                      int64_t C, float* partial, int P, void* stream) {
  CNX_REQUIRE(dy && x && partial && P > 0, CNX_E_BADARG, "dwconv7_wgrad: bad argument");
  CNX_REQUIRE(dtype_ok(dy_dtype) && dtype_ok(x_dtype), CNX_E_BADARG, "dwconv7_wgrad: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_wgrad: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_wgrad: C=%lld must be a multiple of 32", (long long)C);
  return dwconv7_wgrad_v2(dy, dy_dtype, x, x_dtype, N, H, W, C, partial, P, (cudaStream_t)stream);
}

int cnx_dwconv7_wgrad_finalize(const float* partial, int P, int64_t C, int accumulate, float* dw, float* db,
                               void* stream) {
  CNX_REQUIRE(partial && dw && db && P > 0 && C > 0, CNX_E_BADARG, "dwconv7_wgrad_finalize: bad argument");
  launch_pdl(dwconv7_wgrad_finalize_kernel, dim3((unsigned)((50 * C + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, partial, P,
             C, accumulate, dw, db);
  return check_launch("dwconv7_wgrad_finalize");
}

}  // extern "C"

/*
 * cnx.h — C-ABI of libcnx.so: the B200 (sm_100a) kernels behind the ConvNeXt training-step hot path.
 *
 * The reference (abelxiaoxing/ImageClassification) has no FFI of its own: the seam is the duck-typed
 * Python protocol in engine.py / train.py.  These entry points are what the Python
 * host mirror (imageclassification_b200/) binds with ctypes; every function names the reference call
 * it replaces.  Conventions:
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*); no function
 *     synchronises, allocates device memory, or keeps a reference to a caller buffer after it returns
 *     (TMA tensor-map descriptors are encoded on the host at every call — cuTensorMapEncodeTiled, no device work — and passed
 *     to the kernel by value as __grid_constant__ parameters; nothing is cached between calls);
 *   - kernels are launched with programmatic stream serialization (their prologue may overlap the previous kernel's tail;
 *     every kernel waits for its predecessors' memory before its first global access), CNX_PDL=0 in the environment disables it;
 *   - return 0 on success; a negative value for argument errors (CNX_E_*), a positive cudaError_t for a
 *     failed launch.  cnx_last_error_string() describes the last failure on the calling thread;
 *   - no exceptions and no torch types cross this boundary.
 *
 * Layouts: activations are channels-last.  "rows" M = N*H*W pixels, row-major [M, C].
 * dtype enum: CNX_F32 / CNX_BF16.  "stream dtype" = residual stream (block in/out), "act dtype" =
 * saved activations and GEMM operands (bf16 under autocast, fp32 otherwise).
 */
#ifndef CNX_H_
#define CNX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNX_VERSION 100
/* exported symbols (the library is built with -fvisibility=hidden) */
#define CNX_API __attribute__((visibility("default")))

enum { CNX_F32 = 0, CNX_BF16 = 1 };

enum {
  CNX_OK = 0,
  CNX_E_BADARG = -1,   /* null pointer, non-positive size, unsupported dtype combination */
  CNX_E_SHAPE = -2,    /* shape not supported by the kernel (alignment / divisibility) */
  CNX_E_DRIVER = -3,   /* driver entry point / tensor-map encode failure */
  CNX_E_WORKSPACE = -4 /* caller workspace too small */
};

CNX_API int cnx_version(void);
CNX_API const char* cnx_last_error_string(void);
/* Number of SMs of the current device (grid sizing for callers that allocate per-CTA partials). */
CNX_API int cnx_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * a10  ModelEmaV3.update  (engine.py:68,77 -> timm ModelEmaV3.apply_update_ -> torch._foreach_lerp_)
 *   ema <- fmaf(w, p - ema, ema) for every tensor of a pointer table, ONE launch.
 *   table_dev: device array of n_tensors cnx_ema_entry, chunk_start ascending; one CTA per chunk of
 *   CNX_EMA_CHUNK elements.  Bit-exact with ATen lerp on both of its branches (|w| < 0.5: fmaf(w, p - ema, ema);
 *   otherwise p - (p - ema) * (1 - w), e.g. w = 1 while ModelEmaV3.get_decay(step) returns 0).
 * ---------------------------------------------------------------------------------------------- */
#define CNX_EMA_CHUNK 8192
typedef struct {
  void* ema;            /* fp32, updated in place */
  const void* param;    /* fp32 */
  int64_t numel;
  int64_t chunk_start;  /* index of this tensor's first chunk */
} cnx_ema_entry;
CNX_API int cnx_ema_lerp_multi(const void* table_dev, int n_tensors, int64_t total_chunks, float w, void* stream);

/* Fused AdamW + EMA over the same kind of pointer table (optim_factory.py:74-75 +
 * engine.py:73-77).  torch.optim.AdamW semantics (decoupled weight decay, no amsgrad) with the rounding of its default CUDA
 * (foreach) path: hyper-parameters cross the ABI as doubles, every derived scalar (1 - lr*wd, 1 - beta1, 1 - beta2,
 * -lr / bias_correction1, ...) is formed in double and rounded once to fp32 as Python + ATen do. */
typedef struct {
  void* param;        /* fp32, updated */
  const void* grad;   /* fp32 */
  void* exp_avg;      /* fp32, updated */
  void* exp_avg_sq;   /* fp32, updated */
  void* ema;          /* fp32, updated; may be NULL */
  int64_t numel;
  int64_t chunk_start;
} cnx_adamw_entry;
CNX_API int cnx_adamw_ema_multi(const void* table_dev, int n_tensors, int64_t total_chunks, double lr, double beta1,
                        double beta2, double eps, double weight_decay, double bias_correction1,
                        double bias_correction2_sqrt, float ema_w, void* stream);

/* Gradient 2-norm and clipping over a pointer table (utils.py:427-468: NativeScalerWithGradNormCount -> get_grad_norm_ /
 * torch.nn.utils.clip_grad_norm_).  Entries are cnx_ema_entry with `ema` = the gradient tensor (`param` unused).
 *   cnx_grad_sumsq_multi: partial[total_chunks] scratch; out[0] = ||g||_2 (partials summed in a fixed order, in double),
 *                         out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (1 when max_norm <= 0).  Two launches.
 *   cnx_scale_multi:      g *= *coef for every tensor of the table (coef is a DEVICE scalar: no host round trip). */
CNX_API int cnx_grad_sumsq_multi(const void* table_dev, int n_tensors, int64_t total_chunks, float* partial, float max_norm,
                         float* out, void* stream);
CNX_API int cnx_scale_multi(const void* table_dev, int n_tensors, int64_t total_chunks, const float* coef, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a9  SoftTargetCrossEntropy (train.py:257, engine.py:49,52):
 *     loss = mean_n sum_k -t[n,k] * log_softmax(x)[n,k]
 *   fwd: x [B,K] (fp32|bf16), t [B,K] fp32 -> loss[1] fp32, lse[B] fp32 (saved for bwd).
 *        row_loss [B] fp32 and counter [1] u32 (zeroed by the callee at exit) are caller scratch; the
 *        final mean is summed in a fixed order (deterministic).
 *   bwd: dx[n,k] = (softmax(x)[n,k] * sum_k t[n,:] - t[n,k]) * dloss / B
 * ---------------------------------------------------------------------------------------------- */
CNX_API int cnx_soft_target_ce_fwd(const void* x, int x_dtype, const float* t, int64_t B, int64_t K, float* loss,
                           float* lse, float* row_loss, unsigned int* counter, void* stream);
CNX_API int cnx_soft_target_ce_bwd(const void* x, int x_dtype, const float* t, const float* lse, const float* dloss,
                           int64_t B, int64_t K, void* dx, int dx_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a11  timm mixup_target (engine.py:44 -> Mixup.__call__):
 *     off = s/K; on = 1 - s + off (double, then rounded to fp32)
 *     out[b,k] = fl(fl(y(t[b])[k] * lam) + fl(y(t[B-1-b])[k] * (1-lam)))   — three separate roundings.
 * ---------------------------------------------------------------------------------------------- */
CNX_API int cnx_mixup_target(const int64_t* target, int64_t B, int64_t K, double lam, double smoothing, float* out,
                     void* stream);

/* a11 (image half)  timm Mixup._mix_batch (engine.py:44) on a contiguous fp32 [B,C,H,W] batch, IN PLACE, pairs (i, B-1-i):
 *   use_cutmix == 0:  x_i <- fl(fl(lam*x_i) + fl((1-lam)*x_{B-1-i}))   (three separate roundings, as mul_/mul_/add_ give)
 *   use_cutmix != 0:  x[i,:,yl:yh,xl:xh] <- x[B-1-i,:,yl:yh,xl:xh]    (lam is not used; the caller area-corrects it)
 * orig (may be NULL) receives the un-mixed batch in the same pass (engine.py:40's second device copy). */
CNX_API int cnx_mixup_batch(float* x, float* orig, int64_t B, int64_t C, int64_t H, int64_t W, double lam, int use_cutmix,
                    int yl, int yh, int xl, int xh, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a1+a2  Block.dwconv + Block.norm (semantic_segmentation/backbone/convnext.py:34-35,45-47)
 *   y  = dwconv7x7(x) + bias              (rounded to act dtype, as autocast's conv output is)
 *   xn = LayerNorm_C(y) * ln_w + ln_b     (eps, biased variance, fp32 math; rounded to act dtype)
 *   x [N,H,W,C] stream dtype; wt [49,C] fp32 = the canonical [C,1,7,7] parameter in TAP-MAJOR order (made by
 *   cnx_dwconv7_weight_prep once per weight update; the kernels fetch 32-channel slices of it by TMA);
 *   outputs y, xn [M,C] act dtype; mean, rstd [M] fp32.
 * ---------------------------------------------------------------------------------------------- */
/* w [C,49] -> wt [49,C] */
CNX_API int cnx_dwconv7_weight_prep(const float* w, int64_t C, float* wt, void* stream);
CNX_API int cnx_dwconv7_ln_fwd(const void* x, int x_dtype, const float* wt, const float* bias, const float* ln_w,
                       const float* ln_b, float eps, int64_t N, int64_t H, int64_t W, int64_t C, void* y,
                       void* xn, int act_dtype, float* mean, float* rstd, void* stream);

/* Stand-alone channels-last LayerNorm forward (used by LayerNorm2d in stem/downsample/head; convnext.py:175-176).
 * x [M,C] (x_dtype) -> out [M,C] (out_dtype), mean/rstd [M] fp32. */
CNX_API int cnx_ln_fwd(const void* x, int x_dtype, const float* ln_w, const float* ln_b, float eps, int64_t M, int64_t C,
               void* out, int out_dtype, float* mean, float* rstd, void* stream);

/* LayerNorm backward.  dy = rstd * (g - mean_C(g) - xhat * mean_C(g*xhat)), g = dxn*ln_w, xhat=(y-mean)*rstd.
 * partial [P, 2, C] fp32 receives per-CTA column sums of (dxn*xhat, dxn); P = number of CTAs the caller
 * wants launched (>=1); reduce with cnx_reduce_partials. */
CNX_API int cnx_ln_bwd(const void* dxn, int dxn_dtype, const void* y, int y_dtype, const float* mean, const float* rstd,
               const float* ln_w, int64_t M, int64_t C, void* dy, int dy_dtype, float* partial, int P,
               void* stream);

/* Downsample LayerNorm (convnext.py:84-89) fused with the 2x2 stride-2 patch gather: x [N,H,W,C] is normalised per pixel and
 * written in patch-major order out [N,H/2,W/2,(ky,kx,C)] = the A operand of the patchify GEMM; the backward reads the GEMM's
 * data gradient dxn in that same order.  mean/rstd/y/dy stay in pixel order. */
CNX_API int cnx_ln_fwd_patch2(const void* x, int x_dtype, const float* ln_w, const float* ln_b, float eps, int64_t N, int64_t H,
                      int64_t W, int64_t C, void* out, int out_dtype, float* mean, float* rstd, void* stream);
CNX_API int cnx_ln_bwd_patch2(const void* dxn, int dxn_dtype, const void* y, int y_dtype, const float* mean, const float* rstd,
                      const float* ln_w, int64_t N, int64_t H, int64_t W, int64_t C, void* dy, int dy_dtype, float* partial,
                      int P, void* stream);

/* out[j] = (accumulate ? out[j] : 0) + scale * sum_p partial[p, j], fixed order (deterministic). */
CNX_API int cnx_reduce_partials(const float* partial, int P, int64_t L, float scale, int accumulate, float* out,
                        void* stream);
/* The same over a [P, La + Lb] partial whose column halves belong to two tensors (LayerNorm: d ln_w | d ln_b), so that each
 * half can be written — or accumulated — straight into its own gradient buffer (e.g. a slot of the data-parallel arena). */
CNX_API int cnx_reduce_partials_split(const float* partial, int P, int64_t La, int64_t Lb, int accumulate, float* out_a,
                              float* out_b, void* stream);

/* dwconv backward-data (+ residual-gradient add): dx = dres + dwconv7x7_flipped(dy).  dres may be NULL.  wt tap-major.
 * dy [M,C] act dtype; dres, dx [M,C] stream dtype. */
CNX_API int cnx_dwconv7_dgrad(const void* dy, int dy_dtype, const float* wt, const void* dres, void* dx, int stream_dtype,
                      int64_t N, int64_t H, int64_t W, int64_t C, void* stream);
/* The same with bf16 activations (dy bf16), additionally writing the operand copy that the UPSTREAM Block's backward would
 * otherwise make of dx with cnx_grad_prep: dz_up[m,c] = bf16(dp_up[n] * dx[m,c]) (dp_up: that Block's per-sample drop-path
 * scale [N], NULL = 1), bit-identical to cnx_grad_prep(dx, ...).  Saves a read of dx per Block (engine.py:69 backward). */
CNX_API int cnx_dwconv7_dgrad_dz(const void* dy, const float* wt, const void* dres, void* dx, int stream_dtype, int64_t N,
                         int64_t H, int64_t W, int64_t C, void* dz_up, const float* dp_up, void* stream);

/* dwconv backward-weights + bias grad: partial [P, 50, C] fp32 (taps 0..48 then bias), P CTAs (persistent). */
CNX_API int cnx_dwconv7_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W,
                      int64_t C, float* partial, int P, void* stream);
/* dw[c, t] (+)= sum_p partial[p, t, c]; db[c] (+)= sum_p partial[p, 49, c]  (transposes to the canonical layout) */
CNX_API int cnx_dwconv7_wgrad_finalize(const float* partial, int P, int64_t C, int accumulate, float* dw, float* db,
                               void* stream);

/* ------------------------------------------------------------------------------------------------
 * a3..a8  pwconv1 -> GELU -> pwconv2 -> gamma -> drop_path -> residual (convnext.py:36-40,48-55)
 * All GEMMs compute  acc[M,N] = A[M,K] . B[N,K]^T  (both operands K-major = nn.Linear layout), fp32
 * accumulate.  dtype = operand/activation dtype: CNX_BF16 -> tcgen05/TMEM/TMA kernels, CNX_F32 ->
 * fp32 CUDA-core kernels (parity mode).  flags: CNX_GEMM_FORCE_SIMT routes bf16 through the CUDA-core
 * kernel (test-only cross-check of the tensor-core path).
 * ---------------------------------------------------------------------------------------------- */
#define CNX_GEMM_FORCE_SIMT 1
/* cnx_gemm_bias_scale_residual_fwd / cnx_gemm_plain: A is a two-segment split operand [hi | mid] with 2K/3 columns (written by
 * cnx_gemm_bias_gelu_fwd_x3, or by cnx_split3 / cnx_dwconv7_ln_fwd_x3 with segments = 2); the K loop covers three segments and
 * the third re-reads the first ([hi | mid | hi]) — the operand crosses HBM as 2 pieces instead of 3. */
#define CNX_GEMM_A_SPLIT2 2
/* cnx_gemm_plain, bf16 operands, fp32 output: store fp32(bf16(acc + bias)) — the values a bf16 output widened to fp32 by its
 * consumer would hold (autocast's Conv2d output entering the fp32 residual stream, convnext.py:84-89 -> :55), without writing
 * the bf16 tensor and re-reading it in a cast pass. */
#define CNX_GEMM_OUT_ROUND_BF16 4

/* fc1: h = round_dtype(A.W1^T + b1) ; g_out = GELU_erf(h) ; gprime_out = GELU_erf'(h) (what backward needs of h:
 * saved instead of h so the dgrad epilogue is one multiply).  gprime_out may be NULL (no-grad forward). */
CNX_API int cnx_gemm_bias_gelu_fwd(const void* A, const void* W1, const float* b1, int64_t M, int64_t N, int64_t K,
                           void* gprime_out, void* g_out, int dtype, int flags, void* stream);

/* fc2: out[m,n] = shortcut[m,n] + dp[m / rows_per_sample] * gamma[n] * (acc[m,n] + b2[n]).
 * dp NULL -> 1 (eval / no drop-path); gamma NULL -> 1; shortcut NULL -> 0.  out/shortcut stream dtype. */
CNX_API int cnx_gemm_bias_scale_residual_fwd(const void* A, const void* W2, const float* b2, const float* gamma,
                                     const float* dp, int64_t rows_per_sample, const void* shortcut, void* out,
                                     int stream_dtype, int64_t M, int64_t N, int64_t K, int dtype, int flags,
                                     void* stream);

/* fp32-ACCURATE forward on the tensor cores ("x3" operands).  A value x is carried as two bf16 pieces, hi = bf16(x) and
 * mid = bf16(x - hi) (x = hi + mid to 2^-17 relative).  With the A operand laid out [hi | mid | hi] and the B operand
 * [hi | hi | mid] along K, ONE ordinary bf16 GEMM over K3 = 3K accumulates a_hi.b_hi + a_mid.b_hi + a_hi.b_mid in fp32 —
 * everything of the fp32 product except the 2^-16 a_mid.b_mid term.  Used for the no-grad forward outside autocast (the
 * accuracy forward of engine.py:89-97 and evaluate(), engine.py:145-225, run in fp32 in the reference).
 *   cnx_split3                  x fp32 [M,C] -> out bf16 [M,3C] = [hi | mid | hi]                    (A side)
 *   cnx_weight_prep mode 3      W fp32 [R,C] -> out bf16 [R,3C] = [hi | hi | mid]                    (B side)
 *   cnx_gemm_bias_gelu_fwd_x3   g2 bf16 [M,2N] = [hi | mid] of GELU_erf(A.W^T + b1) (fp32 epilogue), A3 [M,K3], W3 [N,K3]
 *   fc2: cnx_gemm_bias_scale_residual_fwd(A = g2, W2 = mode-3 weight, K = 3*4C, dtype = CNX_BF16, stream_dtype = CNX_F32,
 *        flags = CNX_GEMM_A_SPLIT2: the third K segment re-reads g2's hi columns, so g leaves and re-enters HBM as 2 pieces)
 *   plain: cnx_gemm_plain(A3, B3, ..., out fp32, K = 3K, dtype = CNX_BF16) */
/* `segments` = 3: out [M,3C] = [hi | mid | hi];  `segments` = 2: out [M,2C] = [hi | mid] for a consumer that wraps its K loop
 * (CNX_GEMM_A_SPLIT2 / a_segments = 2). */
CNX_API int cnx_split3(const float* x, int64_t M, int64_t C, void* out, int segments, void* stream);
/* cnx_dwconv7_ln_fwd on an fp32 stream whose LayerNorm rows leave directly as the A-side split operand xn3 bf16 [M,3C]
 * (no fp32 xn, no separate cnx_split3 pass).  y_scratch fp32 [M,C] holds the conv output between the two halves of the kernel. */
CNX_API int cnx_dwconv7_ln_fwd_x3(const float* x, const float* wt, const float* bias, const float* ln_w, const float* ln_b,
                          float eps, int64_t N, int64_t H, int64_t W, int64_t C, float* y_scratch, void* xn3, float* mean,
                          float* rstd, int segments, void* stream);
/* a_segments = 3: A3 is [M, K3]; a_segments = 2: A3 is [M, 2*K3/3] = [hi | mid] and the K loop wraps (K3/3 a multiple of 32). */
CNX_API int cnx_gemm_bias_gelu_fwd_x3(const void* A3, const void* W3, const float* b1, int64_t M, int64_t N, int64_t K3,
                              void* g3, int a_segments, void* stream);
/* The same for fp32 TRAINING on the tensor cores: additionally writes gprime [M,N] fp32 = GELU_erf'(acc + b1) (what backward needs
 * of the pre-activation) from the same epilogue — replaces cnx_gemm_plain (fp32 h) + cnx_gelu_split, i.e. one write and one read
 * of an fp32 [M,4C] tensor per Block (convnext.py:48-49 under engine.py:52 without autocast). */
CNX_API int cnx_gemm_bias_gelu_fwd_x3_train(const void* A3, const void* W3, const float* b1, int64_t M, int64_t N, int64_t K3,
                                    void* g3, float* gprime, int a_segments, void* stream);
/* fp32 training, data gradient of fc2 with the GELU derivative, leaving as a split operand:
 *   dh2 [M,2N] bf16 = [hi | mid] of (dz . Bt^T)[m,n] * gprime[m,n]      (dz2 [M,2K] or [M,3K] split, Bt3 [N,3K] = [hi|hi|mid],
 * gprime fp32 [M,N] as written by cnx_gemm_bias_gelu_fwd_x3_train) — replaces cnx_gemm_plain (fp32 product) + cnx_mul_split. */
CNX_API int cnx_gemm_dgrad_gelu_bwd_x3(const void* dz2, const void* Bt3, const float* gprime, void* dh2, int64_t M, int64_t N,
                               int64_t K3, int a_segments, void* stream);

/* fp32 TRAINING on the tensor cores (the reference's default --use_amp false): the same split-operand scheme for the training
 * forward and the four backward GEMMs.  Elementwise pieces between the GEMMs:
 *   cnx_gelu_split   h fp32 [M,N] (fc1 pre-activation incl. bias, from cnx_gemm_plain) -> g2 bf16 [M,2N] = [hi | mid] of GELU_erf(h)
 *                    (A operand of fc2); h is overwritten with GELU_erf'(h) (saved for backward)
 *   cnx_mul_split    out2 bf16 [M,2N] = [hi | mid] of t * u   (dh = (dz.W2s) * GELU'(h))
 *   cnx_gemm_wgrad_x3  out fp32 [N1,N2] (+)= X^T.Y from X2 [M,2*N1] = [hi | mid], Y2 [M,2*N2] = [hi | mid]: three bf16 wgrad GEMMs
 *                    (hi.hi + mid.hi + hi.mid) accumulated in fp32; colsum_x [N1] (+)= column sums of X (may be NULL).
 *                    workspace as cnx_gemm_wgrad_workspace_bytes(M, N1, N2, CNX_BF16, 0). */
CNX_API int cnx_gelu_split(float* h, int64_t M, int64_t N, void* g2, void* stream);
CNX_API int cnx_mul_split(const float* t, const float* u, int64_t M, int64_t N, void* out2, void* stream);
CNX_API int cnx_gemm_wgrad_x3(const void* X2, const void* Y2, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                      float* colsum_x, void* workspace, int64_t workspace_bytes, void* stream);
/* The same in ONE launch: the K loop of the wgrad kernel walks hi.hi, mid.hi, hi.mid over the same rows (one set of split-K
 * partials, one reduction: 18 % less time).  Its fp32 accumulation chains in tensor memory are three times as long, and the
 * tensor core's fp32 accumulate is not round-to-nearest: measured against float64 on N(0,1) operands at the ConvNeXt-T / batch-256
 * shapes the error is 3.2e-5 .. 6.2e-5 (max-abs-normalised) where cnx_gemm_wgrad_x3 has 0.9e-5 .. 2.1e-5 and a cuBLAS fp32 GEMM
 * 0.07e-5 .. 0.5e-5 (profiles/wgrad_x3_accuracy.py).  Inside the 1e-4 bar, but the product path keeps the three-launch form;
 * ops.WGRAD_X3_ONE_LOOP / CNX_WGRAD_X3_ONE_LOOP=1 selects this one. */
CNX_API int cnx_gemm_wgrad_x3_one_loop(const void* X2, const void* Y2, int64_t M, int64_t N1, int64_t N2, int accumulate,
                               float* out, float* colsum_x, void* workspace, int64_t workspace_bytes, void* stream);

/* Fused no-grad MLP forward for the HBM-bound stages (C in {96, 128, 192}; bf16 operands, fp32 residual stream):
 *   out[m,:] = shortcut[m,:] + dp[m / rows_per_sample] * gamma * (GELU_erf(xn[m,:].W1^T + b1).W2^T + b2)
 * in one kernel — the [M,4C] hidden activation stays in shared / tensor memory (convnext.py:48-55 in one pass).
 * xn [M,C] bf16, W1 [4C,C] bf16, W2 [C,4C] bf16, shortcut/out [M,C] fp32; dp, gamma may be NULL. */
CNX_API int cnx_mlp_fused_fwd(const void* xn, const void* W1, const float* b1, const void* W2, const float* b2,
                      const float* gamma, const float* dp, int64_t rows_per_sample, const void* shortcut, void* out,
                      int64_t M, int64_t C, void* stream);

/* The same on split operands (fp32-accurate "x3" forward, C = 96; no-grad pass of an fp32 model or the fp32 accuracy forward of
 * engine.py:89-97): xn2 [M,2C] bf16 = [hi | mid] (cnx_dwconv7_ln_fwd_x3 with segments = 2), W1x3 [4C,3C] and W2x3 [C,12C] bf16 =
 * [hi | hi | mid] (cnx_weight_prep mode 3), shortcut/out [M,C] fp32.  Every product is hi.hi + mid.hi + hi.mid with fp32
 * accumulation, GELU is the fp32 erf form of cnx_gemm_bias_gelu_fwd_x3; the hidden activation (2 x 8 bytes per element through
 * HBM in the unfused pair) stays on chip. */
CNX_API int cnx_mlp_fused_fwd_x3(const void* xn2, const void* W1x3, const float* b1, const void* W2x3, const float* b2,
                         const float* gamma, const float* dp, int64_t rows_per_sample, const float* shortcut, float* out,
                         int64_t M, int64_t C, void* stream);

/* dgrad of fc2 with GELU': dh[m,n] = acc[m,n] * gprime[m,n],  acc = dz.W2s (B given as [N=4C, K=C]),
 * gprime = GELU'(h) as saved by cnx_gemm_bias_gelu_fwd. */
CNX_API int cnx_gemm_dgrad_gelu_bwd(const void* dz, const void* Bt, const void* gprime, void* dh, int64_t M, int64_t N,
                            int64_t K, int dtype, int flags, void* stream);
/* The same WITHOUT a saved GELU'(h): the fc1 pre-activation is recomputed in the kernel, dh = acc * GELU'(round(xn.W1^T + b1))
 * (second accumulator in TMEM; xn [M,K], W1 [N,K] in the activation dtype, b1 fp32).  For the HBM-bound stages (C <= 192),
 * where the tensor pipe idles: the forward then stores g only (cnx_gemm_bias_gelu_fwd with gprime_out = NULL), i.e. ONE
 * [M,4C] tensor per Block instead of two.  bf16 only, N % 128 == 0, M >= 256.  Bit-identical to the saved-GELU' path. */
CNX_API int cnx_gemm_dgrad_gelu_recompute_bwd(const void* dz, const void* Bt, const void* xn, const void* W1, const float* b1,
                                      void* dh, int64_t M, int64_t N, int64_t K, int dtype, void* stream);

/* plain GEMM with cast epilogue: out = A.B^T (dgrad of fc1; also patchify convs).  bias may be NULL. */
CNX_API int cnx_gemm_plain(const void* A, const void* B, const float* bias, void* out, int out_dtype, int64_t M,
                   int64_t N, int64_t K, int dtype, int flags, void* stream);

/* wgrad: out[N1,N2] (+)= sum_m X[m,i] * Y[m,j]   (X [M,N1], Y [M,N2], act dtype; fp32 out).
 * colsum_x (optional, [N1] fp32) (+)= sum_m X[m,i]  (the bias gradient; on the bf16 path the tensor core
 * computes it with an all-ones B tile).  workspace: scratch of >= cnx_gemm_wgrad_workspace_bytes(...)
 * bytes.  Deterministic split-K over M (fp32 partials reduced in a fixed order). */
CNX_API int64_t cnx_gemm_wgrad_workspace_bytes(int64_t M, int64_t N1, int64_t N2, int dtype, int flags);
CNX_API int cnx_gemm_wgrad(const void* X, const void* Y, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                   float* colsum_x, void* workspace, int64_t workspace_bytes, int dtype, int flags, void* stream);

/* Gradient prep at block-backward entry: dz[m,c] = act(dp[n] * dout[m,c]) (dp NULL -> 1).
 * dout stream dtype -> dz act dtype. */
CNX_API int cnx_grad_prep(const void* dout, int stream_dtype, const float* dp, int64_t rows_per_sample, int64_t M,
                  int64_t C, void* dz, int act_dtype, void* stream);

/* Weight prep (tiny, [R,Ccols] fp32 -> act dtype):
 *   mode 0: out[r,c]   = W[r,c]                       (cast)
 *   mode 1: out[c,r]   = W[r,c]                       (transpose + cast)      -> W1^T for dgrad fc1
 *   mode 2: out[c,r]   = row_scale[r] * W[r,c]        (scale rows, transpose) -> (gamma.W2)^T for dgrad fc2
 *   mode 3: out[r, 0:C | C:2C | 2C:3C] = hi, hi, mid of W[r,c]   (bf16 only; B-side split operand, see cnx_split3) */
CNX_API int cnx_weight_prep(const float* W, int64_t R, int64_t Ccols, const float* row_scale, int mode, void* out,
                    int out_dtype, void* stream);

/* The same for MANY weights in one launch.  table_dev: device array of n_entries
 *   struct { const float* W; const float* row_scale; void* out; int64_t R, Cc; int32_t mode, out_dtype;
 *            int64_t tile_start, tiles_x; }          (tiles_x = ceil(Cc/32))
 * Entries that derive from the same source W (same R, Cc) are adjacent and carry the SAME tile_start (a group): one CTA reads
 * a 32x32 source tile once and writes it into every layout of the group.  tile_start = prefix sum over GROUPS of
 * ceil(R/32)*ceil(Cc/32); total_tiles = that sum over all groups. */
CNX_API int cnx_weight_prep_multi(const void* table_dev, int n_entries, int64_t total_tiles, void* stream);

/* Layer-scale gradient from the UNSCALED fc2 wgrad G2[c,k] = sum_m dz[m,c] g[m,k], s[c] = sum_m dz[m,c]:
 *   dgamma[c] (+)= sum_k W2[c,k]*G2[c,k] + b2[c]*s[c];  dW2[c,k] (+)= gamma[c]*G2[c,k];  db2[c] (+)= gamma[c]*s[c]
 * (identity used instead of saving z = fc2 output).  gamma NULL -> 1 and dgamma skipped. */
CNX_API int cnx_layerscale_finalize(const float* G2, const float* s, const float* W2, const float* b2, const float* gamma,
                            int64_t C, int64_t K4, int accumulate, float* dW2, float* db2, float* dgamma,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Patchify convolutions (stem 4x4 stride 4, downsample 2x2 stride 2; convnext.py:79-89): a k x k
 * stride-k conv is a GEMM over non-overlapping patches.  These entry points only re-order activations into / out
 * of the [patches, k*k*C] operand; the arithmetic is cnx_gemm_plain / cnx_gemm_wgrad.
 * ---------------------------------------------------------------------------------------------- */
/* x [N,Cin,H,W] fp32 NCHW -> out [N*(H/4)*(W/4), Cin*16], column k = (ci*4 + ky)*4 + kx (the flattening of the
 * canonical [Cout,Cin,4,4] weight). */
CNX_API int cnx_patchify4_nchw(const float* x, int64_t N, int64_t Cin, int64_t H, int64_t W, void* out, int out_dtype,
                               void* stream);
/* gather != 0: in [N,H,W,C] -> out [N,H/2,W/2,(ky,kx,C)];  gather == 0: the inverse permutation. */
CNX_API int cnx_patch2(const void* in, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, void* out, int gather,
                       void* stream);

/* Elementwise fp32 -> bf16 cast of a flat buffer (parameter shadow copies). */
CNX_API int cnx_cast_f32_to_bf16(const float* in, int64_t n, void* out, void* stream);

/* head  timm NormMlpClassifierHead.global_pool (SelectAdaptivePool2d 'avg') on a channels-last [N, HW, C] map and its backward:
 *   out[n,c] = mean_hw x[n,hw,c] (fp32);   dx[n,hw,c] = dout[n,c] / HW.   The head's LayerNorm2d and fc then run on the
 *   cnx_ln_fwd / cnx_ln_bwd and cnx_gemm_plain / cnx_gemm_wgrad kernels ([N,C] rows). */
CNX_API int cnx_avgpool_nhwc_fwd(const void* x, int x_dtype, int64_t N, int64_t HW, int64_t C, float* out, void* stream);
CNX_API int cnx_avgpool_nhwc_bwd(const float* dout, int64_t N, int64_t HW, int64_t C, void* dx, int dx_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNX_H_ */

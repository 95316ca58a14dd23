// core.cu — error plumbing + the small bandwidth/latency-bound kernels of the path:
//   a10 EMA multi-tensor lerp, fused AdamW+EMA, a9 SoftTargetCrossEntropy fwd/bwd, a11 mixup_target,
//   partial-sum reduction, fp32->bf16 cast.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <mutex>

namespace cnx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 148; }
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ================================================================================================
// a10: EMA.  One CTA per CNX_EMA_CHUNK elements; the CTA finds its tensor by binary search over
// chunk_start.  ema <- fmaf(w, p - ema, ema): the |w|<0.5 branch of ATen's lerp (bit-exact).
// 12 B/element of HBM traffic; 128-bit loads, 8 independent loads in flight per thread.
// ================================================================================================
template <typename Entry>
__device__ __forceinline__ int find_tensor(const Entry* __restrict__ table, int n, int64_t chunk) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (table[mid].chunk_start <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ATen lerp (aten/src/ATen/native/Lerp.h): self + w*diff for |w| < 0.5, end - diff*(1-w) otherwise; nvcc contracts both
// into one FMA, which is what the explicit fmaf calls below state.
__device__ __forceinline__ float lerp_aten(float self, float end, float w, float one_minus_w, bool small_w) {
  float diff = end - self;
  return small_w ? fmaf(w, diff, self) : fmaf(-diff, one_minus_w, end);
}

__global__ void __launch_bounds__(256) ema_lerp_multi_kernel(const cnx_ema_entry* __restrict__ table, int n_tensors,
                                                              float w) {
  pdl_wait();
  __shared__ cnx_ema_entry ent;
  int64_t chunk = blockIdx.x;
  if (threadIdx.x == 0) ent = table[find_tensor(table, n_tensors, chunk)];
  __syncthreads();
  float* __restrict__ e = (float*)ent.ema;
  const float* __restrict__ p = (const float*)ent.param;
  int64_t base = (chunk - ent.chunk_start) * (int64_t)CNX_EMA_CHUNK;
  int64_t rem = ent.numel - base;
  int n = rem < CNX_EMA_CHUNK ? (int)rem : CNX_EMA_CHUNK;
  e += base;
  p += base;
  bool aligned = ((((uintptr_t)e) | ((uintptr_t)p)) & 15) == 0;
  const bool small_w = fabsf(w) < 0.5f;
  const float omw = 1.0f - w;
  if (aligned && n == CNX_EMA_CHUNK) {
    // 8192 elements / 256 threads = 8 float4 per thread, all loads issued before any use
    float4 ev[8], pv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int idx = (i * 256 + threadIdx.x);
      ev[i] = reinterpret_cast<const float4*>(e)[idx];
      pv[i] = __ldg(reinterpret_cast<const float4*>(p) + idx);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int idx = (i * 256 + threadIdx.x);
      float4 r;
      r.x = lerp_aten(ev[i].x, pv[i].x, w, omw, small_w);
      r.y = lerp_aten(ev[i].y, pv[i].y, w, omw, small_w);
      r.z = lerp_aten(ev[i].z, pv[i].z, w, omw, small_w);
      r.w = lerp_aten(ev[i].w, pv[i].w, w, omw, small_w);
      reinterpret_cast<float4*>(e)[idx] = r;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += 256) {
      e[i] = lerp_aten(e[i], p[i], w, omw, small_w);
    }
  }
}

// Fused AdamW (+EMA).  torch.optim.AdamW as it runs on CUDA by default (torch/optim/adam.py _multi_tensor_adam: every
// scalar is formed in double on the host and rounded once to fp32, each foreach op rounds its result):
//   p *= 1 - lr*wd ; m = lerp(m, g, 1-b1) ; v *= b2 ; v += (1-b2) * (g*g) ;
//   d = sqrt(v) / sqrt(1-b2^t) + eps ; p += -(lr/(1-b1^t)) * (m/d) ; then ema = lerp(ema, p, w).
struct AdamScalars {
  float decay, omb1, b1, beta2, omb2, bc2_sqrt, eps, neg_step, ema_w, ema_omw;
  int small_omb1, small_ema_w;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamScalars& c) {
  p = __fmul_rn(p, c.decay);
  m = lerp_aten(m, g, c.omb1, c.b1, c.small_omb1);
  v = fmaf(c.omb2, __fmul_rn(g, g), __fmul_rn(v, c.beta2));
  float denom = __fadd_rn(__fdiv_rn(sqrtf(v), c.bc2_sqrt), c.eps);
  p = fmaf(c.neg_step, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(256) adamw_ema_multi_kernel(const cnx_adamw_entry* __restrict__ table, int n_tensors,
                                                               const AdamScalars c) {
  pdl_wait();
  __shared__ cnx_adamw_entry ent;
  int64_t chunk = blockIdx.x;
  if (threadIdx.x == 0) ent = table[find_tensor(table, n_tensors, chunk)];
  __syncthreads();
  int64_t base = (chunk - ent.chunk_start) * (int64_t)CNX_EMA_CHUNK;
  int64_t rem = ent.numel - base;
  int n = rem < CNX_EMA_CHUNK ? (int)rem : CNX_EMA_CHUNK;
  float* __restrict__ p = (float*)ent.param + base;
  const float* __restrict__ g = (const float*)ent.grad + base;
  float* __restrict__ m = (float*)ent.exp_avg + base;
  float* __restrict__ v = (float*)ent.exp_avg_sq + base;
  float* __restrict__ e = ent.ema ? (float*)ent.ema + base : nullptr;
  bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v) | ((uintptr_t)e)) & 15) == 0;
  if (aligned && n == CNX_EMA_CHUNK) {
    // 8192 elements / 256 threads = 8 float4 per thread per array, in two halves of 4 (all loads of a half in flight at once)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 pv[4], gv[4], mv[4], vv[4], ev[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int idx = (h * 4 + i) * 256 + threadIdx.x;
        pv[i] = reinterpret_cast<const float4*>(p)[idx];
        gv[i] = __ldg(reinterpret_cast<const float4*>(g) + idx);
        mv[i] = reinterpret_cast<const float4*>(m)[idx];
        vv[i] = reinterpret_cast<const float4*>(v)[idx];
        if (e) ev[i] = reinterpret_cast<const float4*>(e)[idx];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int idx = (h * 4 + i) * 256 + threadIdx.x;
        adamw_one(pv[i].x, gv[i].x, mv[i].x, vv[i].x, c);
        adamw_one(pv[i].y, gv[i].y, mv[i].y, vv[i].y, c);
        adamw_one(pv[i].z, gv[i].z, mv[i].z, vv[i].z, c);
        adamw_one(pv[i].w, gv[i].w, mv[i].w, vv[i].w, c);
        reinterpret_cast<float4*>(p)[idx] = pv[i];
        reinterpret_cast<float4*>(m)[idx] = mv[i];
        reinterpret_cast<float4*>(v)[idx] = vv[i];
        if (e) {
          float4 r;
          r.x = lerp_aten(ev[i].x, pv[i].x, c.ema_w, c.ema_omw, c.small_ema_w);
          r.y = lerp_aten(ev[i].y, pv[i].y, c.ema_w, c.ema_omw, c.small_ema_w);
          r.z = lerp_aten(ev[i].z, pv[i].z, c.ema_w, c.ema_omw, c.small_ema_w);
          r.w = lerp_aten(ev[i].w, pv[i].w, c.ema_w, c.ema_omw, c.small_ema_w);
          reinterpret_cast<float4*>(e)[idx] = r;
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += 256) {
      float pi = p[i], mi = m[i], vi = v[i];
      adamw_one(pi, g[i], mi, vi, c);
      p[i] = pi; m[i] = mi; v[i] = vi;
      if (e) e[i] = lerp_aten(e[i], pi, c.ema_w, c.ema_omw, c.small_ema_w);
    }
  }
}

// Gradient 2-norm over a pointer table (utils.py:456-468, torch.nn.utils.clip_grad_norm_ at utils.py:440): per-chunk sums of
// squares, then ONE CTA sums the partials in a fixed order (deterministic) and forms norm and the clip coefficient.
__global__ void __launch_bounds__(256) grad_sumsq_multi_kernel(const cnx_ema_entry* __restrict__ table, int n_tensors,
                                                                float* __restrict__ partial) {
  pdl_wait();
  __shared__ cnx_ema_entry ent;
  __shared__ float red[8];
  int64_t chunk = blockIdx.x;
  if (threadIdx.x == 0) ent = table[find_tensor(table, n_tensors, chunk)];
  __syncthreads();
  const float* __restrict__ g = (const float*)ent.ema;
  int64_t base = (chunk - ent.chunk_start) * (int64_t)CNX_EMA_CHUNK;
  int64_t rem = ent.numel - base;
  int n = rem < CNX_EMA_CHUNK ? (int)rem : CNX_EMA_CHUNK;
  g += base;
  float acc = 0.f;
  if ((((uintptr_t)g) & 15) == 0 && n == CNX_EMA_CHUNK) {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(g) + i * 256 + threadIdx.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  } else {
    for (int i = threadIdx.x; i < n; i += 256) acc = fmaf(g[i], g[i], acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[chunk] = s;
  }
}

__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* __restrict__ partial, int64_t n, float max_norm,
                                                                 float* __restrict__ out) {
  pdl_wait();
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)partial[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 32; ++i) s += red[i];
    float norm = (float)sqrt(s);
    out[0] = norm;
    out[1] = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
  }
}

__global__ void __launch_bounds__(256) scale_multi_kernel(const cnx_ema_entry* __restrict__ table, int n_tensors,
                                                           const float* __restrict__ coef_p) {
  pdl_wait();
  __shared__ cnx_ema_entry ent;
  const float coef = __ldg(coef_p);
  if (coef == 1.0f) return;                    // torch multiplies by 1.0 here: same values, no traffic
  int64_t chunk = blockIdx.x;
  if (threadIdx.x == 0) ent = table[find_tensor(table, n_tensors, chunk)];
  __syncthreads();
  float* __restrict__ g = (float*)ent.ema;
  int64_t base = (chunk - ent.chunk_start) * (int64_t)CNX_EMA_CHUNK;
  int64_t rem = ent.numel - base;
  int n = rem < CNX_EMA_CHUNK ? (int)rem : CNX_EMA_CHUNK;
  g += base;
  if ((((uintptr_t)g) & 15) == 0 && n == CNX_EMA_CHUNK) {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = reinterpret_cast<const float4*>(g)[i * 256 + threadIdx.x];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i].x *= coef; v[i].y *= coef; v[i].z *= coef; v[i].w *= coef;
      reinterpret_cast<float4*>(g)[i * 256 + threadIdx.x] = v[i];
    }
  } else {
    for (int i = threadIdx.x; i < n; i += 256) g[i] *= coef;
  }
}

// ================================================================================================
// a9: SoftTargetCrossEntropy.  One CTA per row (K up to a few thousand): two passes over the row held
// in registers/L1 (max, then sum-exp and sum t*x), deterministic final mean by the last CTA to finish.
// ================================================================================================
template <typename TX, int THREADS>
__global__ void __launch_bounds__(THREADS) soft_ce_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ t,
                                                              int64_t B, int64_t K, float* __restrict__ loss,
                                                              float* __restrict__ lse_out, float* __restrict__ row_loss,
                                                              unsigned int* __restrict__ counter) {
  pdl_wait();
  __shared__ float red[THREADS / 32];
  __shared__ float bcast;
  __shared__ bool is_last;
  const int64_t row = blockIdx.x;
  const TX* xr = x + row * K;
  const float* tr = t + row * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  float m = -INFINITY;
  for (int64_t k = threadIdx.x; k < K; k += THREADS) m = fmaxf(m, to_f32(xr[k]));
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red[0];
    for (int i = 1; i < THREADS / 32; ++i) mm = fmaxf(mm, red[i]);
    bcast = mm;
  }
  __syncthreads();
  m = bcast;
  __syncthreads();

  float se = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += THREADS) se += expf(to_f32(xr[k]) - m);
  se = warp_sum(se);
  if (lane == 0) red[warp] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < THREADS / 32; ++i) s += red[i];
    bcast = m + logf(s);
  }
  __syncthreads();
  const float lse = bcast;
  __syncthreads();

  // sum_k -t * (x - lse): same per-element form as -target * log_softmax(x)
  float acc = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += THREADS) acc += -tr[k] * (to_f32(xr[k]) - lse);
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < THREADS / 32; ++i) s += red[i];
    row_loss[row] = s;
    lse_out[row] = lse;
    __threadfence();
    unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == (unsigned int)(B - 1));
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed-order tree: each thread strides, then warp/CTA reduce (deterministic for a given B)
    float s = 0.f;
    for (int64_t r = threadIdx.x; r < B; r += THREADS) s += __ldcg(row_loss + r);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < THREADS / 32; ++i) tot += red[i];
      loss[0] = tot / (float)B;
      *counter = 0u;   // leave the scratch counter ready for the next call
    }
  }
}

template <typename TX, typename TD>
__global__ void __launch_bounds__(256) soft_ce_bwd_kernel(const TX* __restrict__ x, const float* __restrict__ t,
                                                          const float* __restrict__ lse, const float* __restrict__ dloss,
                                                          int64_t B, int64_t K, TD* __restrict__ dx) {
  pdl_wait();
  __shared__ float red[8];
  __shared__ float bcast;
  const int64_t row = blockIdx.x;
  const TX* xr = x + row * K;
  const float* tr = t + row * K;
  TD* dr = dx + row * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float st = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += 256) st += tr[k];
  st = warp_sum(st);
  if (lane == 0) red[warp] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    bcast = s;
  }
  __syncthreads();
  st = bcast;
  const float l = lse[row];
  const float scale = dloss[0] / (float)B;
  for (int64_t k = threadIdx.x; k < K; k += 256) {
    float p = expf(to_f32(xr[k]) - l);
    dr[k] = from_f32<TD>((p * st - tr[k]) * scale);
  }
}

// ================================================================================================
// a11: mixup_target.  y(t)[k] = (k == t) ? on : off, out = fl(fl(y1*lam) + fl(y2*oml)).
// __fmul_rn/__fadd_rn keep the three roundings separate (no FMA contraction), as three ATen kernels do.
// ================================================================================================
__global__ void __launch_bounds__(256) mixup_target_kernel(const int64_t* __restrict__ target, int64_t B, int64_t K,
                                                           float on, float off, float lam, float oml,
                                                           float* __restrict__ out) {
  pdl_wait();
  int64_t total = B * K;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t b = i / K, k = i - b * K;
    float y1 = (target[b] == k) ? on : off;
    float y2 = (target[B - 1 - b] == k) ? on : off;
    out[i] = __fadd_rn(__fmul_rn(y1, lam), __fmul_rn(y2, oml));
  }
}

// ================================================================================================
// a11 (image half): timm Mixup._mix_batch on a contiguous fp32 [B, L] batch, in place, sample pairs (i, B-1-i).
//   mixup : x_i <- fl(fl(lam * x_i) + fl((1-lam) * x_j)), both members of the pair from the OLD values — what
//           `x_flipped = x.flip(0).mul_(1-lam); x.mul_(lam).add_(x_flipped)` computes in five ATen passes.
//   cutmix: x[i, :, yl:yh, xl:xh] <-> x[B-1-i, :, yl:yh, xl:xh]  (`x[..box] = x.flip(0)[..box]`).
// One pass: 16-byte loads of both members, both results stored; `orig` (optional) receives the un-mixed pair in the same
// pass (engine.py:40 keeps a second device copy of the batch for the accuracy forward).
// ================================================================================================
__global__ void __launch_bounds__(256) mixup_batch_kernel(float* __restrict__ x, float* __restrict__ orig, int64_t half_b,
                                                          int64_t B, int64_t L4, float lam, float oml) {
  const int64_t total = half_b * L4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t b = i / L4, e = i - b * L4;
    float4* pa = reinterpret_cast<float4*>(x) + b * L4 + e;
    float4* pb = reinterpret_cast<float4*>(x) + (B - 1 - b) * L4 + e;
    const float4 a = *pa, c = *pb;
    if (orig) {
      reinterpret_cast<float4*>(orig)[b * L4 + e] = a;
      reinterpret_cast<float4*>(orig)[(B - 1 - b) * L4 + e] = c;
    }
    float4 ra, rc;
    ra.x = __fadd_rn(__fmul_rn(a.x, lam), __fmul_rn(c.x, oml)); rc.x = __fadd_rn(__fmul_rn(c.x, lam), __fmul_rn(a.x, oml));
    ra.y = __fadd_rn(__fmul_rn(a.y, lam), __fmul_rn(c.y, oml)); rc.y = __fadd_rn(__fmul_rn(c.y, lam), __fmul_rn(a.y, oml));
    ra.z = __fadd_rn(__fmul_rn(a.z, lam), __fmul_rn(c.z, oml)); rc.z = __fadd_rn(__fmul_rn(c.z, lam), __fmul_rn(a.z, oml));
    ra.w = __fadd_rn(__fmul_rn(a.w, lam), __fmul_rn(c.w, oml)); rc.w = __fadd_rn(__fmul_rn(c.w, lam), __fmul_rn(a.w, oml));
    *pa = ra;
    *pb = rc;
  }
}
// odd batch middle sample (timm requires an even batch in Mixup.__call__, but _mix_batch itself is defined for any B):
// x_m <- fl(fl(lam*x_m) + fl((1-lam)*x_m)); also the scalar path when L is not a multiple of 4
__global__ void __launch_bounds__(256) mixup_batch_scalar_kernel(float* __restrict__ x, float* __restrict__ orig, int64_t B,
                                                                 int64_t L, float lam, float oml) {
  const int64_t hb = (B + 1) / 2, total = hb * L;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t b = i / L, e = i - b * L, j = B - 1 - b;
    const float a = x[b * L + e], c = x[j * L + e];
    if (orig) { orig[b * L + e] = a; orig[j * L + e] = c; }
    x[b * L + e] = __fadd_rn(__fmul_rn(a, lam), __fmul_rn(c, oml));
    if (j != b) x[j * L + e] = __fadd_rn(__fmul_rn(c, lam), __fmul_rn(a, oml));
  }
}
__global__ void __launch_bounds__(256) cutmix_swap_kernel(float* __restrict__ x, int64_t B, int64_t C, int64_t H, int64_t W,
                                                          int yl, int yh, int xl, int xh) {
  const int64_t bw = xh - xl, bh = yh - yl;
  const int64_t total = (B / 2) * C * bh * bw;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t r = i;
    const int64_t xx = r % bw; r /= bw;
    const int64_t yy = r % bh; r /= bh;
    const int64_t c = r % C; const int64_t b = r / C;
    const int64_t off = (c * H + yl + yy) * W + xl + xx;
    float* pa = x + b * C * H * W + off;
    float* pb = x + (B - 1 - b) * C * H * W + off;
    const float a = *pa;
    *pa = *pb;
    *pb = a;
  }
}

// ================================================================================================
// deterministic reduction of per-CTA partial sums: out[j] = scale * sum_p partial[p, j]
// ================================================================================================
// 32 columns per CTA; the 8 warps split the P rows (fixed assignment), then meet in shared memory in a fixed order:
// deterministic, and a tall narrow partial buffer (P = several hundred CTAs x 2C columns) no longer runs on one warp's latency
// (columns [0, La) go to out, columns [La, L) to out_b: the two halves of a [P, 2C] LayerNorm partial land in two tensors)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int P, int64_t L,
                                                              float scale, int accumulate, float* __restrict__ out,
                                                              int64_t La, float* __restrict__ out_b) {
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + tx;
  // eight independent loads in flight per thread (the buffer is a few hundred rows of a few hundred columns, read from L2 by a
  // handful of CTAs: with two in flight a launch took 12 us of pure latency); fixed association, so still deterministic
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (j < L) {
    int p = ty;
    for (; p + 56 < P; p += 64) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += partial[(int64_t)(p + 8 * u) * L + j];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (p + 8 * u < P) s[u] += partial[(int64_t)(p + 8 * u) * L + j];
  }
  red[ty][tx] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (ty == 0 && j < L) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][tx];
    s *= scale;
    float* o = (j < La) ? out + j : out_b + (j - La);
    *o = accumulate ? *o + s : s;
  }
}

// the same reduction for the two outputs of one split-K wgrad (weight gradient + bias gradient) in ONE launch
__global__ void __launch_bounds__(256) reduce_partials2_kernel(const float* __restrict__ pa, int64_t La, float* __restrict__ oa,
                                                               const float* __restrict__ pb, int64_t Lb, float* __restrict__ ob,
                                                               int P, int accumulate) {
  pdl_wait();
  int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= La + Lb) return;
  const float* partial = pa;
  float* out = oa;
  int64_t L = La;
  if (j >= La) { partial = pb; out = ob; L = Lb; j -= La; }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int p = 0;
  for (; p + 4 <= P; p += 4) {
    s0 += partial[(int64_t)(p + 0) * L + j];
    s1 += partial[(int64_t)(p + 1) * L + j];
    s2 += partial[(int64_t)(p + 2) * L + j];
    s3 += partial[(int64_t)(p + 3) * L + j];
  }
  for (; p < P; ++p) s0 += partial[(int64_t)p * L + j];
  float s = (s0 + s1) + (s2 + s3);
  out[j] = accumulate ? out[j] + s : s;
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, int64_t n, bf16* __restrict__ out) {
  pdl_wait();
  int64_t i4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  int64_t stride = (int64_t)gridDim.x * 256 * 4;
  bool aligned = ((((uintptr_t)in) & 15) == 0) && ((((uintptr_t)out) & 7) == 0);
  for (; i4 < n; i4 += stride) {
    if (aligned && i4 + 4 <= n) {
      float v[4];
      load4(in + i4, v);
      store4(out + i4, v);
    } else {
      for (int64_t i = i4; i < n && i < i4 + 4; ++i) out[i] = __float2bfloat16_rn(in[i]);
    }
  }
}

}  // namespace cnx

using namespace cnx;

extern "C" {

int cnx_version(void) { return CNX_VERSION; }
const char* cnx_last_error_string(void) { return cnx::g_err; }
int cnx_sm_count(void) { return cnx::sm_count(); }

int cnx_ema_lerp_multi(const void* table_dev, int n_tensors, int64_t total_chunks, float w, void* stream) {
  CNX_REQUIRE(table_dev && n_tensors > 0 && total_chunks > 0, CNX_E_BADARG, "ema_lerp_multi: empty table");
  CNX_REQUIRE(total_chunks < (1ll << 31), CNX_E_SHAPE, "ema_lerp_multi: too many chunks");
  launch_pdl(ema_lerp_multi_kernel, dim3((unsigned)total_chunks), dim3(256), 0, (cudaStream_t)stream, (const cnx_ema_entry*)table_dev,
                                                                                 n_tensors, w);
  return check_launch("ema_lerp_multi");
}

int cnx_adamw_ema_multi(const void* table_dev, int n_tensors, int64_t total_chunks, double lr, double beta1,
                        double beta2, double eps, double weight_decay, double bias_correction1,
                        double bias_correction2_sqrt, float ema_w, void* stream) {
  CNX_REQUIRE(table_dev && n_tensors > 0 && total_chunks > 0, CNX_E_BADARG, "adamw_ema_multi: empty table");
  CNX_REQUIRE(total_chunks < (1ll << 31), CNX_E_SHAPE, "adamw_ema_multi: too many chunks");
  CNX_REQUIRE(bias_correction1 != 0.0 && bias_correction2_sqrt != 0.0, CNX_E_BADARG, "adamw_ema_multi: zero bias correction");
  AdamScalars c;
  c.decay = (float)(1.0 - lr * weight_decay);
  c.omb1 = (float)(1.0 - beta1);
  c.b1 = 1.0f - c.omb1;                        // ATen lerp's (1 - weight) in fp32
  c.small_omb1 = fabsf(c.omb1) < 0.5f;
  c.beta2 = (float)beta2;
  c.omb2 = (float)(1.0 - beta2);
  c.bc2_sqrt = (float)bias_correction2_sqrt;
  c.eps = (float)eps;
  c.neg_step = (float)(-(lr / bias_correction1));
  c.ema_w = ema_w;
  c.ema_omw = 1.0f - ema_w;
  c.small_ema_w = fabsf(ema_w) < 0.5f;
  launch_pdl(adamw_ema_multi_kernel, dim3((unsigned)total_chunks), dim3(256), 0, (cudaStream_t)stream, (const cnx_adamw_entry*)table_dev, n_tensors, c);
  return check_launch("adamw_ema_multi");
}

int cnx_grad_sumsq_multi(const void* table_dev, int n_tensors, int64_t total_chunks, float* partial, float max_norm,
                         float* out, void* stream) {
  CNX_REQUIRE(table_dev && partial && out && n_tensors > 0 && total_chunks > 0, CNX_E_BADARG, "grad_sumsq_multi: bad argument");
  CNX_REQUIRE(total_chunks < (1ll << 31), CNX_E_SHAPE, "grad_sumsq_multi: too many chunks");
  cudaStream_t s = (cudaStream_t)stream;
  launch_pdl(grad_sumsq_multi_kernel, dim3((unsigned)total_chunks), dim3(256), 0, s, (const cnx_ema_entry*)table_dev, n_tensors, partial);
  launch_pdl(grad_norm_finish_kernel, dim3(1), dim3(1024), 0, s, partial, total_chunks, max_norm, out);
  return check_launch("grad_sumsq_multi");
}

int cnx_scale_multi(const void* table_dev, int n_tensors, int64_t total_chunks, const float* coef, void* stream) {
  CNX_REQUIRE(table_dev && coef && n_tensors > 0 && total_chunks > 0, CNX_E_BADARG, "scale_multi: bad argument");
  CNX_REQUIRE(total_chunks < (1ll << 31), CNX_E_SHAPE, "scale_multi: too many chunks");
  launch_pdl(scale_multi_kernel, dim3((unsigned)total_chunks), dim3(256), 0, (cudaStream_t)stream, (const cnx_ema_entry*)table_dev, n_tensors, coef);
  return check_launch("scale_multi");
}

int cnx_soft_target_ce_fwd(const void* x, int x_dtype, const float* t, int64_t B, int64_t K, float* loss,
                           float* lse, float* row_loss, unsigned int* counter, void* stream) {
  CNX_REQUIRE(x && t && loss && lse && row_loss && counter, CNX_E_BADARG, "soft_target_ce_fwd: null pointer");
  CNX_REQUIRE(B > 0 && K > 0 && dtype_ok(x_dtype), CNX_E_BADARG, "soft_target_ce_fwd: bad shape/dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (K <= 64) {
    if (x_dtype == CNX_F32) launch_pdl(soft_ce_fwd_kernel<float, 32>, dim3((unsigned)B), dim3(32), 0, s, (const float*)x, t, B, K, loss, lse, row_loss, counter);
    else launch_pdl(soft_ce_fwd_kernel<bf16, 32>, dim3((unsigned)B), dim3(32), 0, s, (const bf16*)x, t, B, K, loss, lse, row_loss, counter);
  } else {
    if (x_dtype == CNX_F32) launch_pdl(soft_ce_fwd_kernel<float, 256>, dim3((unsigned)B), dim3(256), 0, s, (const float*)x, t, B, K, loss, lse, row_loss, counter);
    else launch_pdl(soft_ce_fwd_kernel<bf16, 256>, dim3((unsigned)B), dim3(256), 0, s, (const bf16*)x, t, B, K, loss, lse, row_loss, counter);
  }
  return check_launch("soft_target_ce_fwd");
}

int cnx_soft_target_ce_bwd(const void* x, int x_dtype, const float* t, const float* lse, const float* dloss,
                           int64_t B, int64_t K, void* dx, int dx_dtype, void* stream) {
  CNX_REQUIRE(x && t && lse && dloss && dx, CNX_E_BADARG, "soft_target_ce_bwd: null pointer");
  CNX_REQUIRE(B > 0 && K > 0 && dtype_ok(x_dtype) && dtype_ok(dx_dtype), CNX_E_BADARG, "soft_target_ce_bwd: bad shape/dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == CNX_F32 && dx_dtype == CNX_F32) launch_pdl(soft_ce_bwd_kernel<float, float>, dim3((unsigned)B), dim3(256), 0, s, (const float*)x, t, lse, dloss, B, K, (float*)dx);
  else if (x_dtype == CNX_BF16 && dx_dtype == CNX_BF16) launch_pdl(soft_ce_bwd_kernel<bf16, bf16>, dim3((unsigned)B), dim3(256), 0, s, (const bf16*)x, t, lse, dloss, B, K, (bf16*)dx);
  else if (x_dtype == CNX_BF16 && dx_dtype == CNX_F32) launch_pdl(soft_ce_bwd_kernel<bf16, float>, dim3((unsigned)B), dim3(256), 0, s, (const bf16*)x, t, lse, dloss, B, K, (float*)dx);
  else launch_pdl(soft_ce_bwd_kernel<float, bf16>, dim3((unsigned)B), dim3(256), 0, s, (const float*)x, t, lse, dloss, B, K, (bf16*)dx);
  return check_launch("soft_target_ce_bwd");
}

int cnx_mixup_target(const int64_t* target, int64_t B, int64_t K, double lam, double smoothing, float* out,
                     void* stream) {
  CNX_REQUIRE(target && out && B > 0 && K > 0, CNX_E_BADARG, "mixup_target: bad argument");
  // timm mixup_target: python-double arithmetic on the host, rounded to fp32 when it meets the tensor
  double off = smoothing / (double)K;
  double on = 1.0 - smoothing + off;
  double oml = 1.0 - lam;
  int64_t total = B * K;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  launch_pdl(mixup_target_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, target, B, K, (float)on, (float)off, (float)lam,
                                                               (float)oml, out);
  return check_launch("mixup_target");
}

int cnx_mixup_batch(float* x, float* orig, int64_t B, int64_t C, int64_t H, int64_t W, double lam, int use_cutmix, int yl,
                    int yh, int xl, int xh, void* stream) {
  CNX_REQUIRE(x && B > 0 && C > 0 && H > 0 && W > 0, CNX_E_BADARG, "mixup_batch: bad argument");
  CNX_REQUIRE(((uintptr_t)x & 15) == 0 && (!orig || ((uintptr_t)orig & 15) == 0), CNX_E_BADARG,
              "mixup_batch: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t L = C * H * W;
  const unsigned cap = 148u * 16u;
  if (use_cutmix) {
    CNX_REQUIRE(0 <= yl && yl <= yh && yh <= H && 0 <= xl && xl <= xh && xh <= W, CNX_E_BADARG, "mixup_batch: bad cutmix box");
    if (orig) {
      cudaError_t e = cudaMemcpyAsync(orig, x, (size_t)(B * L) * sizeof(float), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) { set_error("mixup_batch: copy failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const int64_t total = (B / 2) * C * (int64_t)(yh - yl) * (int64_t)(xh - xl);
    if (total == 0) return 0;
    unsigned grid = (unsigned)((total + 255) / 256 > cap ? cap : (total + 255) / 256);
    cutmix_swap_kernel<<<grid, 256, 0, s>>>(x, B, C, H, W, yl, yh, xl, xh);
    return check_launch("mixup_batch(cutmix)");
  }
  // timm: python doubles lam and (1. - lam) meet the fp32 tensor as fp32 scalars
  const float flam = (float)lam, foml = (float)(1.0 - lam);
  if ((L & 3) == 0 && (B & 1) == 0) {
    const int64_t total = (B / 2) * (L / 4);
    unsigned grid = (unsigned)((total + 255) / 256 > cap ? cap : (total + 255) / 256);
    mixup_batch_kernel<<<grid, 256, 0, s>>>(x, orig, B / 2, B, L / 4, flam, foml);
  } else {
    const int64_t total = ((B + 1) / 2) * L;
    unsigned grid = (unsigned)((total + 255) / 256 > cap ? cap : (total + 255) / 256);
    mixup_batch_scalar_kernel<<<grid, 256, 0, s>>>(x, orig, B, L, flam, foml);
  }
  return check_launch("mixup_batch");
}

int cnx_reduce_partials(const float* partial, int P, int64_t L, float scale, int accumulate, float* out,
                        void* stream) {
  CNX_REQUIRE(partial && out && P > 0 && L > 0, CNX_E_BADARG, "reduce_partials: bad argument");
  launch_pdl(reduce_partials_kernel, dim3((unsigned)((L + 31) / 32)), dim3(256), 0, (cudaStream_t)stream, partial, P, L, scale,
             accumulate, out, L, (float*)nullptr);
  return check_launch("reduce_partials");
}

int cnx_reduce_partials_split(const float* partial, int P, int64_t La, int64_t Lb, int accumulate, float* out_a, float* out_b,
                              void* stream) {
  CNX_REQUIRE(partial && out_a && out_b && P > 0 && La > 0 && Lb > 0, CNX_E_BADARG, "reduce_partials_split: bad argument");
  launch_pdl(reduce_partials_kernel, dim3((unsigned)((La + Lb + 31) / 32)), dim3(256), 0, (cudaStream_t)stream, partial, P, La + Lb,
             1.0f, accumulate, out_a, La, out_b);
  return check_launch("reduce_partials_split");
}

}  // extern "C"

namespace cnx {
int reduce_partials2(const float* pa, int64_t La, float* oa, const float* pb, int64_t Lb, float* ob, int P, int accumulate,
                     cudaStream_t s) {
  launch_pdl(reduce_partials2_kernel, dim3((unsigned)((La + Lb + 255) / 256)), dim3(256), 0, s, pa, La, oa, pb, Lb, ob, P, accumulate);
  return check_launch("reduce_partials2");
}
}  // namespace cnx

extern "C" {

int cnx_cast_f32_to_bf16(const float* in, int64_t n, void* out, void* stream) {
  CNX_REQUIRE(in && out && n > 0, CNX_E_BADARG, "cast_f32_to_bf16: bad argument");
  int64_t blocks = (n / 4 + 255) / 256 + 1;
  if (blocks > 148 * 32) blocks = 148 * 32;
  launch_pdl(cast_f32_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, in, n, (bf16*)out);
  return check_launch("cast_f32_to_bf16");
}

}  // extern "C"

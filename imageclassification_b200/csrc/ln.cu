// ln.cu — a2: channels-last LayerNorm forward (stand-alone) and backward.
// One warp per row; a row's C values sit in registers as NJ groups of 4 (lane + 32*j -th 4-vector), so
// loads/stores are 8-byte (bf16) or 16-byte (fp32) and fully coalesced.  Statistics are two-pass fp32
// (mean, then biased variance), reductions are warp shuffles.  Backward keeps the per-channel sums
// (d ln_w, d ln_b) in warp-private shared-memory columns across all the rows a persistent CTA visits and
// writes one partial row per CTA (deterministic; reduced by cnx_reduce_partials).
#include "common.cuh"
#include <stdlib.h>

namespace cnx {

constexpr int LN_WARPS = 8;

static bool ln_v3_off();
static bool ln_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_LN_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}


// pixel row m of an [N,H,W] grid -> its row in the 2x2-patch-major order [N,H/2,W/2,(ky,kx)] (pH == 0: identity).  Lets the
// downsample LayerNorm write the patchify-GEMM operand directly and its backward read the GEMM's data gradient in place.
__device__ __forceinline__ int64_t patch2_row(int64_t m, int pH, int pW) {
  if (pH == 0) return m;
  const int xx = (int)(m % pW);
  const int64_t t = m / pW;
  const int yy = (int)(t % pH);
  const int64_t n = t / pH;
  return (((n * (pH >> 1) + (yy >> 1)) * (pW >> 1) + (xx >> 1)) << 2) + ((yy & 1) << 1) + (xx & 1);
}

template <typename TX, typename TO, int NJ>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ ln_w,
                                                               const float* __restrict__ ln_b, float eps, int64_t M,
                                                               int C, TO* __restrict__ out, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out, int pH, int pW) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 2;
  float lw[NJ][4], lb[NJ][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    int v = lane + 32 * j;
    if (v < nvec) { load4(ln_w + v * 4, lw[j]); load4(ln_b + v * 4, lb[j]); }
  }
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < M; row += (int64_t)gridDim.x * LN_WARPS) {
    const TX* xr = x + row * C;
    float v[NJ][4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int vi = lane + 32 * j;
      if (vi < nvec) {
        load4(xr + vi * 4, v[j]);
        s += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
      }
    }
    s = warp_sum(s);
    const float mu = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int vi = lane + 32 * j;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { float d = v[j][k] - mu; q = fmaf(d, d, q); }
      }
    }
    q = warp_sum(q);
    const float rs = rsqrtf(q / (float)C + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rs;
    }
    TO* orow = out + patch2_row(row, pH, pW) * C;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int vi = lane + 32 * j;
      if (vi < nvec) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = fmaf((v[j][k] - mu) * rs, lw[j][k], lb[j][k]);
        store4(orow + vi * 4, o);
      }
    }
  }
}


// ---- LayerNorm forward, second generation: LPP lanes per row, 8-element vectors, U rows in flight per warp pass ----
template <typename TX, typename TO, int NJ, int LPP, int U>
__global__ void __launch_bounds__(LN_WARPS * 32, (NJ >= 3 ? 2 : 3)) ln_fwd_v2_kernel(const TX* __restrict__ x, const float* __restrict__ ln_w,
                                                                     const float* __restrict__ ln_b, float eps, int64_t M, int C,
                                                                     TO* __restrict__ out, float* __restrict__ mean_out,
                                                                     float* __restrict__ rstd_out, int pH, int pW) {
  pdl_wait();
  constexpr int PPW = 32 / LPP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPP, l = lane % LPP;
  const int VPR = C >> 3;
  const float invC = 1.0f / (float)C;
  float lw[NJ][8], lb[NJ][8];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int v = l + LPP * j;
#pragma unroll
    for (int e = 0; e < 8; ++e) { lw[j][e] = v < VPR ? __ldg(ln_w + v * 8 + e) : 0.f; lb[j][e] = v < VPR ? __ldg(ln_b + v * 8 + e) : 0.f; }
  }
  const int64_t gw = (int64_t)blockIdx.x * LN_WARPS + warp, tw = (int64_t)gridDim.x * LN_WARPS;
  for (int64_t base = gw * (PPW * U); base < M; base += tw * (PPW * U)) {
    float v[U][NJ][8];
    int64_t row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      row[u] = base + u * PPW + sub;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int vi = l + LPP * j;
        if (row[u] < M && vi < VPR) load8(x + row[u] * C + vi * 8, v[u][j]);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[u][j][e] = 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[u][j][e];
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mu = s * invC;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (l + LPP * j < VPR) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float d = v[u][j][e] - mu; q = fmaf(d, d, q); }
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rs = rsqrtf(q * invC + eps);
      if (row[u] < M) {
        if (l == 0) {
          if (mean_out) mean_out[row[u]] = mu;
          if (rstd_out) rstd_out[row[u]] = rs;
        }
        TO* orow = out + patch2_row(row[u], pH, pW) * C;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int vi = l + LPP * j;
          if (vi < VPR) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = fmaf((v[u][j][e] - mu) * rs, lw[j][e], lb[j][e]);
            store8(orow + vi * 8, o8);
          }
        }
      }
    }
  }
}

// dy = rstd * (g - mean_C(g) - xhat * mean_C(g * xhat)),  g = dxn * ln_w
template <typename TG, typename TY, typename TD, int NJ>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const TG* __restrict__ dxn, const TY* __restrict__ y,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ ln_w, int64_t M, int C,
                                                               TD* __restrict__ dy, float* __restrict__ partial, int pH,
                                                               int pW) {
  pdl_wait();
  extern __shared__ __align__(16) float colacc[];        // [LN_WARPS][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 2;
  float* my_dw = colacc + (size_t)warp * 2 * C;
  float* my_db = my_dw + C;
  for (int i = lane; i < 2 * C; i += 32) my_dw[i] = 0.f;
  float lw[NJ][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    int v = lane + 32 * j;
    if (v < nvec) load4(ln_w + v * 4, lw[j]);
  }
  __syncwarp();
  const float invC = 1.0f / (float)C;
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < M; row += (int64_t)gridDim.x * LN_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    const int64_t grow = patch2_row(row, pH, pW);
    float g[NJ][4], xh[NJ][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int vi = lane + 32 * j;
      if (vi < nvec) {
        float d[4], yv[4];
        load4(dxn + grow * C + vi * 4, d);
        load4(y + row * C + vi * 4, yv);
        float4 aw = *reinterpret_cast<float4*>(my_dw + vi * 4);
        float4 ab = *reinterpret_cast<float4*>(my_db + vi * 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xh[j][k] = (yv[k] - mu) * rs;
          g[j][k] = d[k] * lw[j][k];
          s1 += g[j][k];
          s2 = fmaf(g[j][k], xh[j][k], s2);
        }
        aw.x = fmaf(d[0], xh[j][0], aw.x); aw.y = fmaf(d[1], xh[j][1], aw.y);
        aw.z = fmaf(d[2], xh[j][2], aw.z); aw.w = fmaf(d[3], xh[j][3], aw.w);
        ab.x += d[0]; ab.y += d[1]; ab.z += d[2]; ab.w += d[3];
        *reinterpret_cast<float4*>(my_dw + vi * 4) = aw;
        *reinterpret_cast<float4*>(my_db + vi * 4) = ab;
      }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int vi = lane + 32 * j;
      if (vi < nvec) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = rs * (g[j][k] - s1 - xh[j][k] * s2);
        store4(dy + row * C + vi * 4, o);
      }
    }
  }
  __syncthreads();
  // sum the 8 warp-private columns in a fixed order -> partial[blockIdx.x][2][C]
  for (int i = threadIdx.x; i < 2 * C; i += LN_WARPS * 32) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < LN_WARPS; ++wv) s += colacc[(size_t)wv * 2 * C + i];
    partial[(int64_t)blockIdx.x * 2 * C + i] = s;
  }
}


// ---- LayerNorm backward, second generation -------------------------------------------------------
// LPP lanes per row (32/LPP rows per warp pass), 8-channel (16-byte bf16 / 2x16-byte fp32) vectors, U passes in flight so one
// DRAM round trip covers U rows per lane; a lane always owns the same channels, so d ln_w / d ln_b accumulate in REGISTERS
// over all the rows the warp visits (the first generation did a shared-memory read-modify-write per vector per row) and
// meet in shared memory once at the end.  Output format unchanged: one partial row [2][C] per CTA.
template <typename TG, typename TY, typename TD, int NJ, int LPP, int U>
#ifndef CNX_LNB_MINB
#define CNX_LNB_MINB 3
#endif
#ifndef CNX_LNB_U1
#define CNX_LNB_U1 2
#endif
#ifndef CNX_LNB_U2
#define CNX_LNB_U2 2
#endif
__global__ void __launch_bounds__(LN_WARPS * 32, CNX_LNB_MINB) ln_bwd_v2_kernel(const TG* __restrict__ dxn, const TY* __restrict__ y,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ ln_w, int64_t M, int C,
                                                                  TD* __restrict__ dy, float* __restrict__ partial, int pH,
                                                                  int pW) {
  pdl_wait();
  extern __shared__ __align__(16) float colacc[];        // [LN_WARPS][2][C]
  constexpr int PPW = 32 / LPP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPP, l = lane % LPP;
  const int VPR = C >> 3;
  const float invC = 1.0f / (float)C;
  float lw[NJ][8], aw[NJ][8], ab[NJ][8];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int v = l + LPP * j;
#pragma unroll
    for (int e = 0; e < 8; ++e) { lw[j][e] = v < VPR ? __ldg(ln_w + v * 8 + e) : 0.f; aw[j][e] = 0.f; ab[j][e] = 0.f; }
  }
  const int64_t gw = (int64_t)blockIdx.x * LN_WARPS + warp, tw = (int64_t)gridDim.x * LN_WARPS;
  for (int64_t base = gw * (PPW * U); base < M; base += tw * (PPW * U)) {
    float d[U][NJ][8], yv[U][NJ][8], mu[U], rs[U];
    int64_t row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      row[u] = base + u * PPW + sub;
      const bool ok = row[u] < M;
      mu[u] = ok ? __ldg(mean + row[u]) : 0.f;
      rs[u] = ok ? __ldg(rstd + row[u]) : 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int v = l + LPP * j;
        if (ok && v < VPR) {
          load8(dxn + patch2_row(row[u], pH, pW) * C + v * 8, d[u][j]);
          load8(y + row[u] * C + v * 8, yv[u][j]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { d[u][j][e] = 0.f; yv[u][j][e] = mu[u]; }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (yv[u][j][e] - mu[u]) * rs[u];
          const float g = d[u][j][e] * lw[j][e];
          aw[j][e] = fmaf(d[u][j][e], xh, aw[j][e]);
          ab[j][e] += d[u][j][e];
          yv[u][j][e] = xh;
          d[u][j][e] = g;
          s1 += g;
          s2 = fmaf(g, xh, s2);
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 *= invC;
      s2 *= invC;
      if (row[u] < M) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int v = l + LPP * j;
          if (v < VPR) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = rs[u] * (d[u][j][e] - s1 - yv[u][j][e] * s2);
            store8(dy + row[u] * C + v * 8, o8);
          }
        }
      }
    }
  }
  // rows handled by the two halves of a warp (LPP = 16) meet first, then the warps through shared memory in a fixed order
  if (LPP == 16) {
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        aw[j][e] += __shfl_xor_sync(0xffffffffu, aw[j][e], 16);
        ab[j][e] += __shfl_xor_sync(0xffffffffu, ab[j][e], 16);
      }
  }
  float* my = colacc + (size_t)warp * 2 * C;
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int v = l + LPP * j;
      if (v < VPR) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { my[v * 8 + e] = aw[j][e]; my[C + v * 8 + e] = ab[j][e]; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += LN_WARPS * 32) {
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < LN_WARPS; ++wv) sum += colacc[(size_t)wv * 2 * C + i];
    partial[(int64_t)blockIdx.x * 2 * C + i] = sum;
  }
}

// ---- LayerNorm backward, third generation (the all-bf16 hot path) ----------------------------------
// LPP in {4, 8, 16, 32} lanes per row, each lane NJ <= 4 eight-channel vectors, chosen so that NO lane idles at the ConvNeXt
// widths (C = 96: 4 x 3, 192: 8 x 3, 384: 16 x 3, 768: 32 x 3, 128/256/512/1024: x 4) — the second generation left a quarter of
// the lanes idle at C = 96 / 192 and did not cover C > 256 at all.  Inputs stay PACKED (raw 16-byte vectors) in registers
// between the two passes over a row, so a lane keeps ~NJ*U*8 registers of data in flight instead of NJ*U*16 floats, and the
// output is three FMAs per element:  dy = (rs*w_c)*d + (-rs^2*s2)*y + (rs^2*s2*mu - rs*s1).
// d ln_w / d ln_b accumulate in registers over all rows a lane visits; one partial row [2][C] per CTA, as before.
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
  v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
  v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
  v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <int NJ, int LPP, int U>
__global__ void __launch_bounds__(LN_WARPS * 32, 2) ln_bwd_v3_kernel(const bf16* __restrict__ dxn, const bf16* __restrict__ y,
                                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                     const float* __restrict__ ln_w, int64_t M, int C,
                                                                     bf16* __restrict__ dy, float* __restrict__ partial, int pH,
                                                                     int pW) {
  pdl_wait();
  extern __shared__ __align__(16) float colacc[];        // [LN_WARPS][2][C], then ln_w [C] when it does not fit in registers
  constexpr int PPW = 32 / LPP;
  constexpr bool LW_SMEM = NJ >= 3;                      // 16 accumulators per vector already: keep ln_w in shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPP, l = lane % LPP;
  const int VPR = C >> 3;
  const float invC = 1.0f / (float)C;
  float lwr[LW_SMEM ? 1 : NJ][8], aw[NJ][8], ab[NJ][8];
  float* slw = colacc + (size_t)LN_WARPS * 2 * C;
  if (LW_SMEM) {
    for (int i = threadIdx.x; i < C; i += LN_WARPS * 32) slw[i] = __ldg(ln_w + i);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int v = l + LPP * j;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (!LW_SMEM) lwr[j][e] = v < VPR ? __ldg(ln_w + v * 8 + e) : 0.f;
      aw[j][e] = 0.f; ab[j][e] = 0.f;
    }
  }
  auto get_lw = [&](int j, float (&w8)[8]) {
    if (LW_SMEM) {
      const int v = l + LPP * j;
      const float4* p4 = reinterpret_cast<const float4*>(slw + (v < VPR ? v : 0) * 8);
      const float4 a4 = p4[0], b4 = p4[1];
      w8[0] = a4.x; w8[1] = a4.y; w8[2] = a4.z; w8[3] = a4.w; w8[4] = b4.x; w8[5] = b4.y; w8[6] = b4.z; w8[7] = b4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) w8[e] = lwr[j][e];
    }
  };
  const int64_t gw = (int64_t)blockIdx.x * LN_WARPS + warp, tw = (int64_t)gridDim.x * LN_WARPS;
  for (int64_t base = gw * (PPW * U); base < M; base += tw * (PPW * U)) {
    uint4 rd[U][NJ], ry[U][NJ];
    float mu[U], rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = base + u * PPW + sub;
      const bool ok = row < M;
      mu[u] = ok ? __ldg(mean + row) : 0.f;
      rs[u] = ok ? __ldg(rstd + row) : 0.f;
      const uint4* pd = reinterpret_cast<const uint4*>(dxn + patch2_row(ok ? row : 0, pH, pW) * C);
      const uint4* py = reinterpret_cast<const uint4*>(y + (ok ? row : 0) * C);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int v = l + LPP * j;
        if (ok && v < VPR) {
          rd[u][j] = pd[v];
          ry[u][j] = py[v];
        } else {
          rd[u][j] = make_uint4(0u, 0u, 0u, 0u);
          ry[u][j] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = base + u * PPW + sub;
      const float nmr = -mu[u] * rs[u];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float d[8], yv[8], w8[8];
        unpack8(rd[u][j], d);
        unpack8(ry[u][j], yv);
        get_lw(j, w8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = fmaf(yv[e], rs[u], nmr);
          const float g = d[e] * w8[e];
          aw[j][e] = fmaf(d[e], xh, aw[j][e]);       // padded lanes / rows: d = 0 -> no contribution
          ab[j][e] += d[e];
          s1 += g;
          s2 = fmaf(g, xh, s2);
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 *= invC;
      s2 *= invC;
      // dy = rs*(g - s1 - xh*s2) with xh = rs*y - rs*mu   ->   (rs*w)*d + cy*y + c0
      const float cy = -rs[u] * rs[u] * s2;
      const float c0 = -fmaf(cy, mu[u], rs[u] * s1);
      if (row < M) {
        uint4* po = reinterpret_cast<uint4*>(dy + row * C);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int v = l + LPP * j;
          if (v < VPR) {
            float d[8], yv[8], o8[8], w8[8];
            unpack8(rd[u][j], d);
            unpack8(ry[u][j], yv);
            get_lw(j, w8);
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = fmaf(rs[u] * w8[e], d[e], fmaf(cy, yv[e], c0));
            po[v] = make_uint4(pack2_bf16(o8[0], o8[1]), pack2_bf16(o8[2], o8[3]), pack2_bf16(o8[4], o8[5]),
                               pack2_bf16(o8[6], o8[7]));
          }
        }
      }
    }
  }
  // the PPW row groups of a warp meet by shuffles (fixed order), then the warps through shared memory (fixed order)
#pragma unroll
  for (int o = LPP; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        aw[j][e] += __shfl_xor_sync(0xffffffffu, aw[j][e], o);
        ab[j][e] += __shfl_xor_sync(0xffffffffu, ab[j][e], o);
      }
  }
  float* my = colacc + (size_t)warp * 2 * C;
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int v = l + LPP * j;
      if (v < VPR) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { my[v * 8 + e] = aw[j][e]; my[C + v * 8 + e] = ab[j][e]; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += LN_WARPS * 32) {
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < LN_WARPS; ++wv) sum += colacc[(size_t)wv * 2 * C + i];
    partial[(int64_t)blockIdx.x * 2 * C + i] = sum;
  }
}

static bool ln_v3_off() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_LN_V3");
    v = (e && e[0] == '0') ? 1 : 0;
  }
  return v != 0;
}

static int launch_ln_bwd_v3(const bf16* dxn, const bf16* y, const float* mean, const float* rstd, const float* ln_w, int64_t M,
                            int64_t C, bf16* dy, float* partial, int P, cudaStream_t s, int pH, int pW, bool* handled) {
  *handled = false;
  const int vpr = (int)(C / 8);
  if (C % 8 != 0 || vpr > 128 || ln_v1() || ln_v3_off()) return 0;
  if ((((uintptr_t)dxn) | ((uintptr_t)y) | ((uintptr_t)dy)) & 15) return 0;
  int lpp = 4;
  while ((vpr + lpp - 1) / lpp > 4) lpp <<= 1;
  const int nj = (vpr + lpp - 1) / lpp;
  const size_t smem = (size_t)(LN_WARPS * 2 + 1) * C * sizeof(float);
  *handled = true;
#define CNX_LNB3(NJ, LPP, U)                                                                                      \
  do {                                                                                                            \
    auto k = ln_bwd_v3_kernel<NJ, LPP, U>;                                                                        \
    if (smem > 48 * 1024) {                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
      if (e != cudaSuccess) { set_error("ln_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }          \
    }                                                                                                             \
    launch_pdl(k, dim3((unsigned)P), dim3(LN_WARPS * 32), smem, s, dxn, y, mean, rstd, ln_w, M, (int)C, dy, partial, pH, pW); \
    return check_launch("ln_bwd");                                                                                \
  } while (0)
#define CNX_LNB3_NJ(LPP)                     \
  switch (nj) {                              \
    case 1: CNX_LNB3(1, LPP, 4);             \
    case 2: CNX_LNB3(2, LPP, 2);             \
    case 3: CNX_LNB3(3, LPP, 1);             \
    default: CNX_LNB3(4, LPP, 1);            \
  }
  switch (lpp) {
    case 4: CNX_LNB3_NJ(4);
    case 8: CNX_LNB3_NJ(8);
    case 16: CNX_LNB3_NJ(16);
    default: CNX_LNB3_NJ(32);
  }
#undef CNX_LNB3_NJ
#undef CNX_LNB3
  return 0;
}

static inline int nj_for(int64_t C) { return (int)((C / 4 + 31) / 32); }

#define CNX_NJ_SWITCH(nj, ...)                                   \
  switch (nj) {                                                  \
    case 1: { constexpr int NJ = 1; __VA_ARGS__; } break;        \
    case 2: { constexpr int NJ = 2; __VA_ARGS__; } break;        \
    case 3: { constexpr int NJ = 3; __VA_ARGS__; } break;        \
    case 4: { constexpr int NJ = 4; __VA_ARGS__; } break;        \
    case 5: case 6: { constexpr int NJ = 6; __VA_ARGS__; } break;   \
    case 7: case 8: { constexpr int NJ = 8; __VA_ARGS__; } break;   \
    case 9: case 10: case 11: case 12: { constexpr int NJ = 12; __VA_ARGS__; } break; \
    case 13: case 14: case 15: case 16: { constexpr int NJ = 16; __VA_ARGS__; } break; \
    default: set_error("layer_norm: C too large (max 2048)"); return CNX_E_SHAPE; \
  }

template <typename TX, typename TO>
static int launch_ln_fwd(const void* x, const float* ln_w, const float* ln_b, float eps, int64_t M, int64_t C, void* out,
                         float* mean, float* rstd, cudaStream_t s, int pH = 0, int pW = 0) {
  int64_t blocks = (M + LN_WARPS - 1) / LN_WARPS;
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (C % 24 == 0 && M >= 4096 && !ln_v1() && !ln_v3_off()) {
    // rows of 3 x 2^k sixteen-byte vectors (C = 96, 192): 3 vectors per lane, 4 / 8 lanes per row, no idle lanes
    const int lpp = (int)(C / 24);
    int64_t b3 = (int64_t)sm_count() * 2;
#define CNX_LNF3(LPP)                                                                                                              \
    launch_pdl(ln_fwd_v2_kernel<TX, TO, 3, LPP, 2>, dim3((unsigned)b3), dim3(LN_WARPS * 32), 0, s, (const TX*)x, ln_w, ln_b, eps, M, (int)C, (TO*)out, \
                                                                                mean, rstd, pH, pW)
    if (lpp == 4) { CNX_LNF3(4); return check_launch("ln_fwd"); }
    if (lpp == 8) { CNX_LNF3(8); return check_launch("ln_fwd"); }
    // (C = 384 / 768 measured no better than the one-warp-per-row kernel below: profiles/r01i_kbench_ln_fwd.txt)
#undef CNX_LNF3
  }
  if (C % 8 == 0 && C <= 256 && M >= 4096 && !ln_v1()) {
    int64_t b2 = (int64_t)sm_count() * 3;
    if (C <= 128)
      launch_pdl(ln_fwd_v2_kernel<TX, TO, 1, 16, 4>, dim3((unsigned)b2), dim3(LN_WARPS * 32), 0, s, (const TX*)x, ln_w, ln_b, eps, M, (int)C, (TO*)out,
                                                                                mean, rstd, pH, pW);
    else
      launch_pdl(ln_fwd_v2_kernel<TX, TO, 1, 32, 4>, dim3((unsigned)b2), dim3(LN_WARPS * 32), 0, s, (const TX*)x, ln_w, ln_b, eps, M, (int)C, (TO*)out,
                                                                                mean, rstd, pH, pW);
    return check_launch("ln_fwd");
  }
  CNX_NJ_SWITCH(nj_for(C), (launch_pdl(ln_fwd_kernel<TX, TO, NJ>, dim3((unsigned)blocks), dim3(LN_WARPS * 32), 0, s, 
                               (const TX*)x, ln_w, ln_b, eps, M, (int)C, (TO*)out, mean, rstd, pH, pW)));
  return check_launch("ln_fwd");
}

template <typename TG, typename TY, typename TD>
static int launch_ln_bwd(const void* dxn, const void* y, const float* mean, const float* rstd, const float* ln_w,
                         int64_t M, int64_t C, void* dy, float* partial, int P, cudaStream_t s, int pH = 0, int pW = 0) {
  size_t smem = (size_t)LN_WARPS * 2 * C * sizeof(float);
  // measured (profiles/r01d_*): the register-accumulating kernel wins for rows of up to 32 vectors (C <= 256) at 3 CTAs/SM;
  // wider rows keep the first-generation kernel
  if (C % 8 == 0 && C <= 256 && !ln_v1()) {
#define CNX_LNB2(NJ, LPP, U)                                                                                         \
  do {                                                                                                               \
    auto k = ln_bwd_v2_kernel<TG, TY, TD, NJ, LPP, U>;                                                               \
    if (smem > 48 * 1024) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
      if (e != cudaSuccess) { set_error("ln_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }             \
    }                                                                                                                \
    launch_pdl(k, dim3((unsigned)P), dim3(LN_WARPS * 32), smem, s, (const TG*)dxn, (const TY*)y, mean, rstd, ln_w, M, (int)C,  \
               (TD*)dy, partial, pH, pW);                                                                            \
    return check_launch("ln_bwd");                                                                                   \
  } while (0)
    const int64_t vpr = C / 8;
    if (vpr <= 16) CNX_LNB2(1, 16, CNX_LNB_U1);
    if (vpr <= 32) CNX_LNB2(1, 32, CNX_LNB_U1);
#undef CNX_LNB2
  }
  CNX_NJ_SWITCH(nj_for(C), {
    auto k = ln_bwd_kernel<TG, TY, TD, NJ>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("ln_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    }
    launch_pdl(k, dim3((unsigned)P), dim3(LN_WARPS * 32), smem, s, (const TG*)dxn, (const TY*)y, mean, rstd, ln_w, M, (int)C, (TD*)dy,
               partial, pH, pW);
  });
  return check_launch("ln_bwd");
}

}  // namespace cnx

using namespace cnx;

extern "C" {

static int ln_fwd_impl(const void* x, int x_dtype, const float* ln_w, const float* ln_b, float eps, int64_t M, int64_t C,
                       void* out, int out_dtype, float* mean, float* rstd, void* stream, int pH, int pW) {
  CNX_REQUIRE(x && ln_w && ln_b && out, CNX_E_BADARG, "ln_fwd: null pointer");
  CNX_REQUIRE(dtype_ok(x_dtype) && dtype_ok(out_dtype) && M > 0 && C > 0, CNX_E_BADARG, "ln_fwd: bad shape/dtype");
  CNX_REQUIRE(C % 4 == 0, CNX_E_SHAPE, "ln_fwd: C=%lld must be a multiple of 4", (long long)C);
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == CNX_F32 && out_dtype == CNX_F32) return launch_ln_fwd<float, float>(x, ln_w, ln_b, eps, M, C, out, mean, rstd, s, pH, pW);
  if (x_dtype == CNX_F32 && out_dtype == CNX_BF16) return launch_ln_fwd<float, bf16>(x, ln_w, ln_b, eps, M, C, out, mean, rstd, s, pH, pW);
  if (x_dtype == CNX_BF16 && out_dtype == CNX_F32) return launch_ln_fwd<bf16, float>(x, ln_w, ln_b, eps, M, C, out, mean, rstd, s, pH, pW);
  return launch_ln_fwd<bf16, bf16>(x, ln_w, ln_b, eps, M, C, out, mean, rstd, s, pH, pW);
}

int cnx_ln_fwd(const void* x, int x_dtype, const float* ln_w, const float* ln_b, float eps, int64_t M, int64_t C,
               void* out, int out_dtype, float* mean, float* rstd, void* stream) {
  return ln_fwd_impl(x, x_dtype, ln_w, ln_b, eps, M, C, out, out_dtype, mean, rstd, stream, 0, 0);
}

int cnx_ln_fwd_patch2(const void* x, int x_dtype, const float* ln_w, const float* ln_b, float eps, int64_t N, int64_t H,
                      int64_t W, int64_t C, void* out, int out_dtype, float* mean, float* rstd, void* stream) {
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, CNX_E_SHAPE, "ln_fwd_patch2: H=%lld, W=%lld must be even",
              (long long)H, (long long)W);
  return ln_fwd_impl(x, x_dtype, ln_w, ln_b, eps, N * H * W, C, out, out_dtype, mean, rstd, stream, (int)H, (int)W);
}

static int ln_bwd_impl(const void* dxn, int dxn_dtype, const void* y, int y_dtype, const float* mean, const float* rstd,
                       const float* ln_w, int64_t M, int64_t C, void* dy, int dy_dtype, float* partial, int P,
                       void* stream, int pH, int pW) {
  CNX_REQUIRE(dxn && y && mean && rstd && ln_w && dy && partial && P > 0, CNX_E_BADARG, "ln_bwd: bad argument");
  CNX_REQUIRE(dtype_ok(dxn_dtype) && dtype_ok(y_dtype) && dtype_ok(dy_dtype) && M > 0 && C > 0, CNX_E_BADARG,
              "ln_bwd: bad shape/dtype");
  CNX_REQUIRE(C % 4 == 0, CNX_E_SHAPE, "ln_bwd: C=%lld must be a multiple of 4", (long long)C);
  cudaStream_t s = (cudaStream_t)stream;
  const int key = (dxn_dtype << 2) | (y_dtype << 1) | dy_dtype;
  switch (key) {
    case 0: return launch_ln_bwd<float, float, float>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
    case 7: {
      bool handled = false;
      int rc = launch_ln_bwd_v3((const bf16*)dxn, (const bf16*)y, mean, rstd, ln_w, M, C, (bf16*)dy, partial, P, s, pH, pW, &handled);
      if (handled) return rc;
      return launch_ln_bwd<bf16, bf16, bf16>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
    }
    case 1: return launch_ln_bwd<float, float, bf16>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
    case 2: return launch_ln_bwd<float, bf16, float>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
    case 3: return launch_ln_bwd<float, bf16, bf16>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
    case 4: return launch_ln_bwd<bf16, float, float>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
    case 5: return launch_ln_bwd<bf16, float, bf16>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
    default: return launch_ln_bwd<bf16, bf16, float>(dxn, y, mean, rstd, ln_w, M, C, dy, partial, P, s, pH, pW);
  }
}

int cnx_ln_bwd(const void* dxn, int dxn_dtype, const void* y, int y_dtype, const float* mean, const float* rstd,
               const float* ln_w, int64_t M, int64_t C, void* dy, int dy_dtype, float* partial, int P,
               void* stream) {
  return ln_bwd_impl(dxn, dxn_dtype, y, y_dtype, mean, rstd, ln_w, M, C, dy, dy_dtype, partial, P, stream, 0, 0);
}

int cnx_ln_bwd_patch2(const void* dxn, int dxn_dtype, const void* y, int y_dtype, const float* mean, const float* rstd,
                      const float* ln_w, int64_t N, int64_t H, int64_t W, int64_t C, void* dy, int dy_dtype, float* partial,
                      int P, void* stream) {
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, CNX_E_SHAPE, "ln_bwd_patch2: H=%lld, W=%lld must be even",
              (long long)H, (long long)W);
  return ln_bwd_impl(dxn, dxn_dtype, y, y_dtype, mean, rstd, ln_w, N * H * W, C, dy, dy_dtype, partial, P, stream, (int)H,
                     (int)W);
}

}  // extern "C"

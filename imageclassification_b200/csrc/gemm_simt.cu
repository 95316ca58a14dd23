// gemm_simt.cu — fp32-accumulate CUDA-core GEMMs with the fused Block epilogues.
// This is the fp32 parity path (BASELINE config 1: fp32 <= 1e-4 needs true fp32 products, which the
// tensor cores do not offer) and the on-device cross-check of the tcgen05 kernels (CNX_GEMM_FORCE_SIMT).
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles, operands converted to fp32 in shared memory.
#include "common.cuh"
#include "epilogue.cuh"

namespace cnx {

constexpr int SB_M = 128, SB_N = 128, SB_K = 16;

// acc[M,N] = A[M,K] . B[N,K]^T
template <typename TIN, typename TOUT, int KIND>
__global__ void __launch_bounds__(256) gemm_tn_simt_kernel(const TIN* __restrict__ A, const TIN* __restrict__ B,
                                                           int64_t M, int64_t N, int64_t K, EpiParams ep) {
  __shared__ __align__(16) float As[SB_K][SB_M + 4];
  __shared__ __align__(16) float Bs[SB_K][SB_N + 4];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int64_t m0 = (int64_t)blockIdx.y * SB_M, n0 = (int64_t)blockIdx.x * SB_N;
  const int lrow = t >> 1, lk = (t & 1) * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += SB_K) {
    float va[8], vb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { va[i] = 0.f; vb[i] = 0.f; }
    if (m0 + lrow < M) {
      if (k0 + lk + 8 <= K) load8(A + (m0 + lrow) * K + k0 + lk, va);
      else for (int i = 0; i < 8 && k0 + lk + i < K; ++i) va[i] = to_f32(A[(m0 + lrow) * K + k0 + lk + i]);
    }
    if (n0 + lrow < N) {
      if (k0 + lk + 8 <= K) load8(B + (n0 + lrow) * K + k0 + lk, vb);
      else for (int i = 0; i < 8 && k0 + lk + i < K; ++i) vb[i] = to_f32(B[(n0 + lrow) * K + k0 + lk + i]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) { As[lk + i][lrow] = va[i]; Bs[lk + i][lrow] = vb[i]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const int64_t n = n0 + tx * 8;
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t m = m0 + ty * 8 + i;
      if (m < M) epilogue_store8<KIND, TOUT>(ep, m, n, acc[i]);
    }
  }
}

// wgrad: part[z][i][j] = sum_{m in split z} X[m,i] * Y[m,j];  colsum part[z][i] = sum_m X[m,i]
template <typename TIN>
__global__ void __launch_bounds__(256) gemm_wgrad_simt_kernel(const TIN* __restrict__ X, const TIN* __restrict__ Y,
                                                              int64_t M, int64_t N1, int64_t N2, int64_t m_per_split,
                                                              float* __restrict__ part, float* __restrict__ cs_part) {
  __shared__ __align__(16) float Xs[SB_K][SB_M + 4];
  __shared__ __align__(16) float Ys[SB_K][SB_N + 4];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int64_t i0 = (int64_t)blockIdx.y * SB_M, j0 = (int64_t)blockIdx.x * SB_N;
  const int64_t mb = (int64_t)blockIdx.z * m_per_split;
  const int64_t me = (mb + m_per_split < M) ? mb + m_per_split : M;
  const int lr = t >> 4, lc = (t & 15) * 8;     // 16 rows (m) x 16 groups of 8 columns
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;

  for (int64_t k0 = mb; k0 < me; k0 += SB_K) {
    float vx[8], vy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { vx[i] = 0.f; vy[i] = 0.f; }
    if (k0 + lr < me) {
      if (i0 + lc < N1) load8(X + (k0 + lr) * N1 + i0 + lc, vx);
      if (j0 + lc < N2) load8(Y + (k0 + lr) * N2 + j0 + lc, vy);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Xs[lr][lc]) = make_float4(vx[0], vx[1], vx[2], vx[3]);
    *reinterpret_cast<float4*>(&Xs[lr][lc + 4]) = make_float4(vx[4], vx[5], vx[6], vx[7]);
    *reinterpret_cast<float4*>(&Ys[lr][lc]) = make_float4(vy[0], vy[1], vy[2], vy[3]);
    *reinterpret_cast<float4*>(&Ys[lr][lc + 4]) = make_float4(vy[4], vy[5], vy[6], vy[7]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&Xs[k][ty * 8]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&Xs[k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Ys[k][tx * 8]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Ys[k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        cs[i] += a[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
  float* po = part + (int64_t)blockIdx.z * N1 * N2;
  const int64_t j = j0 + tx * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t ii = i0 + ty * 8 + i;
    if (ii < N1 && j < N2) store8(po + ii * N2 + j, acc[i]);
    if (cs_part && blockIdx.x == 0 && tx == 0 && ii < N1) cs_part[(int64_t)blockIdx.z * N1 + ii] = cs[i];
  }
}

static int wgrad_splits_simt(int64_t M, int64_t N1, int64_t N2) {
  int64_t tiles = ((N1 + SB_M - 1) / SB_M) * ((N2 + SB_N - 1) / SB_N);
  int64_t want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
  int64_t maxs = (M + 255) / 256;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  return (int)want;
}

int64_t wgrad_workspace_bytes_simt(int64_t M, int64_t N1, int64_t N2) {
  return (int64_t)wgrad_splits_simt(M, N1, N2) * (N1 * N2 + N1) * 4;
}

template <typename TIN>
int gemm_wgrad_simt(const void* X, const void* Y, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                    float* colsum_x, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
  int splits = wgrad_splits_simt(M, N1, N2);
  CNX_REQUIRE(workspace_bytes >= (int64_t)splits * (N1 * N2 + N1) * 4, CNX_E_WORKSPACE, "gemm_wgrad: workspace too small");
  int64_t mps = ((M + splits - 1) / splits + SB_K - 1) / SB_K * SB_K;
  float* part = (float*)workspace;
  float* cs_part = part + (int64_t)splits * N1 * N2;
  dim3 grid((unsigned)((N2 + SB_N - 1) / SB_N), (unsigned)((N1 + SB_M - 1) / SB_M), (unsigned)splits);
  gemm_wgrad_simt_kernel<TIN><<<grid, 256, 0, s>>>((const TIN*)X, (const TIN*)Y, M, N1, N2, mps, part,
                                                   colsum_x ? cs_part : nullptr);
  if (int rc = check_launch("gemm_wgrad_simt")) return rc;
  if (int rc = cnx_reduce_partials(part, splits, N1 * N2, 1.0f, accumulate, out, s)) return rc;
  if (colsum_x) return cnx_reduce_partials(cs_part, splits, N1, 1.0f, accumulate, colsum_x, s);
  return 0;
}
template int gemm_wgrad_simt<float>(const void*, const void*, int64_t, int64_t, int64_t, int, float*, float*, void*,
                                    int64_t, cudaStream_t);
template int gemm_wgrad_simt<bf16>(const void*, const void*, int64_t, int64_t, int64_t, int, float*, float*, void*,
                                   int64_t, cudaStream_t);

template <typename TIN, typename TOUT, int KIND>
int gemm_tn_simt(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s) {
  dim3 grid((unsigned)((N + SB_N - 1) / SB_N), (unsigned)((M + SB_M - 1) / SB_M));
  CNX_REQUIRE(grid.y < 65536, CNX_E_SHAPE, "gemm_simt: M too large for the CUDA-core path");
  gemm_tn_simt_kernel<TIN, TOUT, KIND><<<grid, 256, 0, s>>>((const TIN*)A, (const TIN*)B, M, N, K, ep);
  return check_launch("gemm_tn_simt");
}

#define CNX_INST(TIN, TOUT, KIND) \
  template int gemm_tn_simt<TIN, TOUT, KIND>(const void*, const void*, int64_t, int64_t, int64_t, const EpiParams&, cudaStream_t);
CNX_INST(float, float, EPI_PLAIN)
CNX_INST(float, float, EPI_BIAS_GELU)
CNX_INST(float, float, EPI_SCALE_RES)
CNX_INST(float, float, EPI_DGELU)
CNX_INST(bf16, bf16, EPI_PLAIN)
CNX_INST(bf16, float, EPI_PLAIN)
CNX_INST(bf16, bf16, EPI_BIAS_GELU)
CNX_INST(bf16, float, EPI_SCALE_RES)
CNX_INST(bf16, bf16, EPI_SCALE_RES)
CNX_INST(bf16, bf16, EPI_DGELU)
#undef CNX_INST

// ---- small prep / finalize kernels -------------------------------------------------------------

template <typename TS, typename TA>
__global__ void __launch_bounds__(256) grad_prep_kernel(const TS* __restrict__ dout, const float* __restrict__ dp,
                                                        int64_t rows_per_sample, int64_t M, int64_t C,
                                                        TA* __restrict__ dz) {
  pdl_wait();
  const int64_t vec_per_row = C >> 3;
  const int64_t total = M * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t m = i / vec_per_row;
    float s = dp ? dp[m / rows_per_sample] : 1.0f;
    float v[8];
    load8(dout + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= s;
    store8(dz + i * 8, v);
  }
}

// fp32 [M, C] -> bf16 [M, 3C] = [hi | mid | hi]: the A-side split operand (x ~ hi + mid to 2^-17 relative); 8 elements per thread
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, int64_t M, int64_t C, bf16* __restrict__ out,
                                                     int seg) {
  pdl_wait();
  const int64_t vpr = C >> 3, total = M * vpr;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t m = i / vpr, v = i - m * vpr;
    float a[8], r[8];
    load8(x + m * C + v * 8, a);
    uint4 hi;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      h2[k] = __floats2bfloat162_rn(a[2 * k], a[2 * k + 1]);
      r[2 * k] = a[2 * k] - __low2float(h2[k]);
      r[2 * k + 1] = a[2 * k + 1] - __high2float(h2[k]);
    }
    bf16* o = out + m * seg * C + v * 8;
    *reinterpret_cast<uint4*>(o) = hi;
    store8(o + C, r);
    if (seg == 3) *reinterpret_cast<uint4*>(o + 2 * C) = hi;
  }
}

__device__ __forceinline__ void split_store8(bf16* o, int64_t N, const float (&a)[8]) {
  uint4 hi;
  float r[8];
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    h2[k] = __floats2bfloat162_rn(a[2 * k], a[2 * k + 1]);
    r[2 * k] = a[2 * k] - __low2float(h2[k]);
    r[2 * k + 1] = a[2 * k + 1] - __high2float(h2[k]);
  }
  *reinterpret_cast<uint4*>(o) = hi;
  store8(o + N, r);
}

// fp32 training on the tensor cores (split operands): h [M,N] fp32 (fc1 pre-activation incl. bias) ->
//   g2 [M,2N] bf16 = [hi | mid] of GELU_erf(h)  (A operand of fc2),  h <- GELU_erf'(h) in place (what backward needs of h)
__global__ void __launch_bounds__(256) gelu_split_kernel(float* __restrict__ h, int64_t M, int64_t N, bf16* __restrict__ g2) {
  pdl_wait();
  const int64_t vpr = N >> 3, total = M * vpr;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t m = i / vpr, v = i - m * vpr;
    float a[8], g[8];
    load8(h + m * N + v * 8, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) { g[k] = gelu_erf(a[k]); a[k] = gelu_erf_grad(a[k]); }
    split_store8(g2 + m * 2 * N + v * 8, N, g);
    store8(h + m * N + v * 8, a);
  }
}

// out2 [M,2N] bf16 = [hi | mid] of t[m,n] * u[m,n]  (dh = (dz.W2s) * GELU'(h), leaving as the split operand of the next GEMMs)
__global__ void __launch_bounds__(256) mul_split_kernel(const float* __restrict__ t, const float* __restrict__ u, int64_t M, int64_t N,
                                                        bf16* __restrict__ out2) {
  pdl_wait();
  const int64_t vpr = N >> 3, total = M * vpr;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t m = i / vpr, v = i - m * vpr;
    float a[8], b[8];
    load8(t + m * N + v * 8, a);
    load8(u + m * N + v * 8, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= b[k];
    split_store8(out2 + m * 2 * N + v * 8, N, a);
  }
}

template <typename TO>
__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ W, int64_t R, int64_t Cc,
                                                          const float* __restrict__ row_scale, int mode,
                                                          TO* __restrict__ out) {
  pdl_wait();
  // 32x32 shared-memory transpose tiles; mode 0 is a straight cast
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
  for (int i = ty; i < 32; i += 8) {
    int64_t r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < Cc) {
      v = W[r * Cc + c];
      if (mode == 2 && row_scale) v *= row_scale[r];
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (mode == 0) {
    for (int i = ty; i < 32; i += 8) {
      int64_t r = r0 + i, c = c0 + tx;
      if (r < R && c < Cc) out[r * Cc + c] = from_f32<TO>(tile[i][tx]);
    }
  } else if (mode == 3) {
    // split operand, B side: out[r, 0:Cc] = hi, [Cc:2Cc] = hi, [2Cc:3Cc] = mid   (hi = bf16(w), mid = bf16(w - hi))
    for (int i = ty; i < 32; i += 8) {
      int64_t r = r0 + i, c = c0 + tx;
      if (r < R && c < Cc) {
        const float w = tile[i][tx];
        const TO hi = from_f32<TO>(w);
        const TO mid = from_f32<TO>(w - to_f32(hi));
        out[r * 3 * Cc + c] = hi;
        out[r * 3 * Cc + Cc + c] = hi;
        out[r * 3 * Cc + 2 * Cc + c] = mid;
      }
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      int64_t c = c0 + i, r = r0 + tx;
      if (r < R && c < Cc) out[c * R + r] = from_f32<TO>(tile[tx][i]);
    }
  }
}

// All derived weight layouts of a model in ONE launch: a device table of (source, scale, destination, shape, mode) entries,
// one 32x32 SOURCE tile per CTA found by binary search over the entries' first-tile prefix.  Entries of one source weight are
// adjacent and share `tile_start` (a group): the CTA reads the fp32 tile once and writes every layout derived from it (bf16
// copy, transposes, gamma-scaled transpose, split operand) — a Block's fc weights have five.  Replaces 4 launches per Block.
struct WeightPrepEntry {
  const float* W;
  const float* row_scale;
  void* out;
  int64_t R, Cc;
  int32_t mode, out_dtype;
  int64_t tile_start;      // first CTA of this entry's group
  int64_t tiles_x;         // ceil(Cc / 32)
};

__global__ void __launch_bounds__(256) weight_prep_multi_kernel(const WeightPrepEntry* __restrict__ table, int n) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.x;
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].tile_start <= b) lo = mid; else hi = mid - 1;
  }
  const int last = lo;                                   // last entry of the group; walk back to its first
  const int64_t start = table[last].tile_start;
  int first = last;
  while (first > 0 && table[first - 1].tile_start == start) --first;
  const WeightPrepEntry e0 = table[first];
  const int64_t t = b - start;
  const int64_t r0 = (t / e0.tiles_x) * 32, c0 = (t % e0.tiles_x) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < e0.R && c < e0.Cc) ? e0.W[r * e0.Cc + c] : 0.f;
  }
  __syncthreads();
  for (int g = first; g <= last; ++g) {
    const WeightPrepEntry e = table[g];
    for (int i = ty; i < 32; i += 8) {
      int64_t r, c, o;
      float v;
      if (e.mode == 0 || e.mode == 3) { r = r0 + i; c = c0 + tx; o = r * e.Cc + c; v = tile[i][tx]; }
      else { c = c0 + i; r = r0 + tx; o = c * e.R + r; v = tile[tx][i]; }
      if (r < e.R && c < e.Cc) {
        if (e.mode == 2 && e.row_scale) v *= e.row_scale[r];
        if (e.mode == 3) {                                   // [hi | hi | mid], bf16 only
          bf16* o3 = reinterpret_cast<bf16*>(e.out) + r * 3 * e.Cc + c;
          const bf16 hi = from_f32<bf16>(v);
          o3[0] = hi;
          o3[e.Cc] = hi;
          o3[2 * e.Cc] = from_f32<bf16>(v - to_f32(hi));
        } else if (e.out_dtype == CNX_F32) reinterpret_cast<float*>(e.out)[o] = v;
        else reinterpret_cast<bf16*>(e.out)[o] = from_f32<bf16>(v);
      }
    }
  }
}

// one CTA per channel c: dgamma[c] = sum_k W2[c,k]*G2[c,k] + b2[c]*s[c]; dW2[c,:] = gamma[c]*G2[c,:]
__global__ void __launch_bounds__(256) layerscale_finalize_kernel(const float* __restrict__ G2, const float* __restrict__ s,
                                                                  const float* __restrict__ W2, const float* __restrict__ b2,
                                                                  const float* __restrict__ gamma, int64_t C, int64_t K4,
                                                                  int accumulate, float* __restrict__ dW2,
                                                                  float* __restrict__ db2, float* __restrict__ dgamma) {
  pdl_wait();
  __shared__ float red[8];
  const int64_t c = blockIdx.x;
  const float gm = gamma ? gamma[c] : 1.0f;
  float dot = 0.f;
  for (int64_t k = threadIdx.x; k < K4; k += 256) {
    float g = G2[c * K4 + k];
    dot = fmaf(W2[c * K4 + k], g, dot);
    float v = gm * g;
    float* o = dW2 + c * K4 + k;
    *o = accumulate ? *o + v : v;
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float sc = s[c];
    float v = gm * sc;
    db2[c] = accumulate ? db2[c] + v : v;
    if (dgamma && gamma) {
      float dg = tot + b2[c] * sc;
      dgamma[c] = accumulate ? dgamma[c] + dg : dg;
    }
  }
}

}  // namespace cnx

using namespace cnx;

extern "C" {

int cnx_grad_prep(const void* dout, int stream_dtype, const float* dp, int64_t rows_per_sample, int64_t M,
                  int64_t C, void* dz, int act_dtype, void* stream) {
  CNX_REQUIRE(dout && dz && M > 0 && C > 0 && rows_per_sample > 0, CNX_E_BADARG, "grad_prep: bad argument");
  CNX_REQUIRE(dtype_ok(stream_dtype) && dtype_ok(act_dtype), CNX_E_BADARG, "grad_prep: bad dtype");
  CNX_REQUIRE(C % 8 == 0, CNX_E_SHAPE, "grad_prep: C must be a multiple of 8");
  int64_t total = M * (C / 8);
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (stream_dtype == CNX_F32 && act_dtype == CNX_F32) launch_pdl(grad_prep_kernel<float, float>, dim3((unsigned)blocks), dim3(256), 0, s, (const float*)dout, dp, rows_per_sample, M, C, (float*)dz);
  else if (stream_dtype == CNX_F32 && act_dtype == CNX_BF16) launch_pdl(grad_prep_kernel<float, bf16>, dim3((unsigned)blocks), dim3(256), 0, s, (const float*)dout, dp, rows_per_sample, M, C, (bf16*)dz);
  else if (stream_dtype == CNX_BF16 && act_dtype == CNX_BF16) launch_pdl(grad_prep_kernel<bf16, bf16>, dim3((unsigned)blocks), dim3(256), 0, s, (const bf16*)dout, dp, rows_per_sample, M, C, (bf16*)dz);
  else { set_error("grad_prep: unsupported dtype combination"); return CNX_E_BADARG; }
  return check_launch("grad_prep");
}

int cnx_weight_prep(const float* W, int64_t R, int64_t Ccols, const float* row_scale, int mode, void* out,
                    int out_dtype, void* stream) {
  CNX_REQUIRE(W && out && R > 0 && Ccols > 0 && mode >= 0 && mode <= 3 && dtype_ok(out_dtype), CNX_E_BADARG,
              "weight_prep: bad argument");
  CNX_REQUIRE(mode != 3 || out_dtype == CNX_BF16, CNX_E_BADARG, "weight_prep: mode 3 (split operand) writes bf16");
  dim3 grid((unsigned)((Ccols + 31) / 32), (unsigned)((R + 31) / 32));
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == CNX_F32) launch_pdl(weight_prep_kernel<float>, dim3(grid), dim3(256), 0, s, W, R, Ccols, row_scale, mode, (float*)out);
  else launch_pdl(weight_prep_kernel<bf16>, dim3(grid), dim3(256), 0, s, W, R, Ccols, row_scale, mode, (bf16*)out);
  return check_launch("weight_prep");
}

int cnx_split3(const float* x, int64_t M, int64_t C, void* out, int segments, void* stream) {
  CNX_REQUIRE(x && out && M > 0 && C > 0 && (segments == 2 || segments == 3), CNX_E_BADARG, "split3: bad argument");
  CNX_REQUIRE(C % 8 == 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)out) & 15) == 0, CNX_E_SHAPE,
              "split3: C=%lld must be a multiple of 8 and the pointers 16-byte aligned", (long long)C);
  const int64_t total = M * (C / 8);
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  launch_pdl(split3_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, x, M, C, (bf16*)out, segments);
  return check_launch("split3");
}

int cnx_weight_prep_multi(const void* table_dev, int n_entries, int64_t total_tiles, void* stream) {
  CNX_REQUIRE(table_dev && n_entries > 0 && total_tiles > 0 && total_tiles < (1ll << 31), CNX_E_BADARG,
              "weight_prep_multi: bad argument");
  launch_pdl(weight_prep_multi_kernel, dim3((unsigned)total_tiles), dim3(256), 0, (cudaStream_t)stream, (const WeightPrepEntry*)table_dev, n_entries);
  return check_launch("weight_prep_multi");
}

int cnx_gelu_split(float* h, int64_t M, int64_t N, void* g2, void* stream) {
  CNX_REQUIRE(h && g2 && M > 0 && N > 0, CNX_E_BADARG, "gelu_split: bad argument");
  CNX_REQUIRE(N % 8 == 0 && (((uintptr_t)h) & 15) == 0 && (((uintptr_t)g2) & 15) == 0, CNX_E_SHAPE,
              "gelu_split: N=%lld must be a multiple of 8 and the pointers 16-byte aligned", (long long)N);
  int64_t grid = (M * (N / 8) + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  launch_pdl(gelu_split_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, h, M, N, (bf16*)g2);
  return check_launch("gelu_split");
}

int cnx_mul_split(const float* t, const float* u, int64_t M, int64_t N, void* out2, void* stream) {
  CNX_REQUIRE(t && u && out2 && M > 0 && N > 0, CNX_E_BADARG, "mul_split: bad argument");
  CNX_REQUIRE(N % 8 == 0 && ((((uintptr_t)t) | ((uintptr_t)u) | ((uintptr_t)out2)) & 15) == 0, CNX_E_SHAPE,
              "mul_split: N=%lld must be a multiple of 8 and the pointers 16-byte aligned", (long long)N);
  int64_t grid = (M * (N / 8) + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  launch_pdl(mul_split_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, t, u, M, N, (bf16*)out2);
  return check_launch("mul_split");
}

int cnx_layerscale_finalize(const float* G2, const float* s, const float* W2, const float* b2, const float* gamma,
                            int64_t C, int64_t K4, int accumulate, float* dW2, float* db2, float* dgamma,
                            void* stream) {
  CNX_REQUIRE(G2 && s && W2 && b2 && dW2 && db2 && C > 0 && K4 > 0, CNX_E_BADARG, "layerscale_finalize: bad argument");
  launch_pdl(layerscale_finalize_kernel, dim3((unsigned)C), dim3(256), 0, (cudaStream_t)stream, G2, s, W2, b2, gamma, C, K4, accumulate,
                                                                           dW2, db2, dgamma);
  return check_launch("layerscale_finalize");
}

}  // extern "C"

// dwconv.cu — a1/a2: depthwise 7x7 conv (+fused channels-last LayerNorm) forward, backward-data
// (+residual-gradient add) and backward-weights.  Channels-last, shared-memory halo tiles.
//
// Tiling (all three kernels): a CTA owns TH x TW output pixels of one image and walks the channels in
// chunks of 32 (lane == channel, so every shared-memory access is a conflict-free 128-byte row).  Each
// warp owns CPW adjacent output columns and a vertical run of TH rows: an input value loaded from
// shared memory feeds up to 7*CPW FMAs from registers (the 49 taps of the lane's channel live in
// registers).  Halo pixels outside the image are zero-filled, which is the conv's padding=3.
//
// Forward additionally keeps the conv output tile [TH*TW, C] in shared memory (rounded to the
// activation dtype, exactly what autocast's conv hands to layer_norm), then normalises it per pixel with
// warp-shuffle reductions (two-pass mean / biased variance in fp32) and writes y and xn with 128-bit
// stores.  Algorithmic bytes: 2*MC*e + 8M.
#include "common.cuh"

namespace cnx {

constexpr int CH = 32;        // channels per chunk == warp width
constexpr int NWARP = 8;      // warps per CTA

enum { MODE_FWD_LN = 0, MODE_DGRAD = 1 };

template <int TH, int CPW>
struct DwTile {
  static constexpr int TW = NWARP * CPW;
  static constexpr int HH = TH + 6;
  static constexpr int HW = TW + 6;
  static constexpr int HALO_FLOATS = HH * HW * CH;
  static constexpr int W_FLOATS = CH * 49;
};

// load one 32-channel chunk of the halo tile into shared memory as fp32 (zero outside the image)
template <typename TIN, int HH, int HW>
__device__ __forceinline__ void load_halo(float* __restrict__ sm, const TIN* __restrict__ img, int64_t H, int64_t W,
                                          int64_t C, int y0, int x0, int c0) {
  // 8 lanes x 4 channels cover the 32-channel chunk of one pixel; a warp covers 4 pixels per pass
  const int sub = threadIdx.x & 7;          // which 4-channel group
  const int pl = threadIdx.x >> 3;          // pixel slot 0..31
  for (int p = pl; p < HH * HW; p += 32) {
    int iy = p / HW, ix = p - iy * HW;
    int gy = y0 + iy - 3, gx = x0 + ix - 3;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) load4(img + ((int64_t)gy * W + gx) * C + c0 + sub * 4, v);
    *reinterpret_cast<float4*>(sm + p * CH + sub * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// The register-tiled 7x7 correlation for one chunk: acc[q][r] for output column (CPW*warp+q), row r.
template <int TH, int CPW, bool FLIP>
__device__ __forceinline__ void conv_chunk(const float* __restrict__ halo, const float* __restrict__ wsm, int warp,
                                           int lane, float (&acc)[CPW][TH]) {
  constexpr int HW = NWARP * CPW + 6;
  float wr[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wr[t] = wsm[lane * 49 + (FLIP ? 48 - t : t)];
#pragma unroll
  for (int q = 0; q < CPW; ++q)
#pragma unroll
    for (int r = 0; r < TH; ++r) acc[q][r] = 0.f;
#pragma unroll
  for (int j = 0; j < 6 + CPW; ++j) {           // input column CPW*warp + j of the halo tile
#pragma unroll
    for (int iy = 0; iy < TH + 6; ++iy) {
      float v = halo[(iy * HW + CPW * warp + j) * CH + lane];
#pragma unroll
      for (int q = 0; q < CPW; ++q) {
        const int kx = j - q;                   // tap column for output column q
        if (kx >= 0 && kx <= 6) {
#pragma unroll
          for (int ky = 0; ky < 7; ++ky) {
            const int r = iy - ky;
            if (r >= 0 && r < TH) acc[q][r] = fmaf(v, wr[ky * 7 + kx], acc[q][r]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward: dwconv + bias + LayerNorm
// grid = (tiles_x * tiles_y, N); dynamic smem = halo + weights + ytile[TH*TW][C] (TACT) + stats
// ------------------------------------------------------------------------------------------------
template <typename TIN, typename TACT, int TH, int CPW>
__global__ void __launch_bounds__(NWARP * 32) dwconv7_ln_fwd_kernel(
    const TIN* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int64_t H, int64_t W, int64_t C,
    int tiles_x, TACT* __restrict__ y, TACT* __restrict__ xn, float* __restrict__ mean_out,
    float* __restrict__ rstd_out) {
  typedef DwTile<TH, CPW> T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* halo = reinterpret_cast<float*>(smem_raw);
  float* wsm = halo + T::HALO_FLOATS;
  float* stat = wsm + T::W_FLOATS;                       // [TH*TW][2] mean, rstd
  TACT* ytile = reinterpret_cast<TACT*>(stat + TH * T::TW * 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * T::TW;
  const int64_t n = blockIdx.y;
  const TIN* img = x + n * H * W * C;

  for (int c0 = 0; c0 < C; c0 += CH) {
    __syncthreads();                                     // previous chunk's readers are done
    load_halo<TIN, T::HH, T::HW>(halo, img, H, W, C, y0, x0, c0);
    for (int i = threadIdx.x; i < T::W_FLOATS; i += NWARP * 32) wsm[i] = w[(int64_t)c0 * 49 + i];
    __syncthreads();
    float acc[CPW][TH];
    conv_chunk<TH, CPW, false>(halo, wsm, warp, lane, acc);
    const float b = bias[c0 + lane];
#pragma unroll
    for (int q = 0; q < CPW; ++q)
#pragma unroll
      for (int r = 0; r < TH; ++r) {
        int p = r * T::TW + CPW * warp + q;
        ytile[(int64_t)p * C + c0 + lane] = from_f32<TACT>(acc[q][r] + b);
      }
  }
  __syncthreads();

  // per-pixel statistics: one warp per pixel, two-pass (mean, then biased variance) in fp32
  for (int p = warp; p < TH * T::TW; p += NWARP) {
    const TACT* row = ytile + (int64_t)p * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += to_f32(row[c]);
    s = warp_sum(s);
    const float mu = s / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { float d = to_f32(row[c]) - mu; q = fmaf(d, d, q); }
    q = warp_sum(q);
    const float rs = rsqrtf(q / (float)C + eps);
    if (lane == 0) {
      stat[2 * p] = mu;
      stat[2 * p + 1] = rs;
      int r = p / T::TW, cx = p - r * T::TW;
      int gy = y0 + r, gx = x0 + cx;
      if (gy < H && gx < W) {
        int64_t m = (n * H + gy) * W + gx;
        mean_out[m] = mu;
        rstd_out[m] = rs;
      }
    }
  }
  __syncthreads();

  // write y and xn: flattened (pixel, 8-channel vector) so every lane stores 16 B (bf16) / 32 B (fp32)
  const int vec_per_px = (int)(C >> 3);
  const int total = TH * T::TW * vec_per_px;
  for (int i = threadIdx.x; i < total; i += NWARP * 32) {
    int p = i / vec_per_px, cv = (i - p * vec_per_px) << 3;
    int r = p / T::TW, cx = p - r * T::TW;
    int gy = y0 + r, gx = x0 + cx;
    if (gy >= H || gx >= W) continue;
    float v[8], o[8];
    load8(ytile + (int64_t)p * C + cv, v);
    const float mu = stat[2 * p], rs = stat[2 * p + 1];
    float lw[8], lb[8];
    load8(ln_w + cv, lw);
    load8(ln_b + cv, lb);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf((v[k] - mu) * rs, lw[k], lb[k]);
    int64_t off = ((n * H + gy) * W + gx) * C + cv;
    store8(y + off, v);
    store8(xn + off, o);
  }
}

// ------------------------------------------------------------------------------------------------
// backward-data: dx = dres + corr(dy, flipped w).  Output written straight from registers
// (32 lanes x 4 B = one 128-byte line per pixel-chunk for an fp32 residual stream).
// ------------------------------------------------------------------------------------------------
template <typename TIN, typename TOUT, int TH, int CPW>
__global__ void __launch_bounds__(NWARP * 32) dwconv7_dgrad_kernel(const TIN* __restrict__ dy,
                                                                    const float* __restrict__ w,
                                                                    const TOUT* __restrict__ dres, int64_t H, int64_t W,
                                                                    int64_t C, int tiles_x, TOUT* __restrict__ dx) {
  typedef DwTile<TH, CPW> T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* halo = reinterpret_cast<float*>(smem_raw);
  float* wsm = halo + T::HALO_FLOATS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * T::TW;
  const int64_t n = blockIdx.y;
  const TIN* img = dy + n * H * W * C;
  for (int c0 = 0; c0 < C; c0 += CH) {
    __syncthreads();
    load_halo<TIN, T::HH, T::HW>(halo, img, H, W, C, y0, x0, c0);
    for (int i = threadIdx.x; i < T::W_FLOATS; i += NWARP * 32) wsm[i] = w[(int64_t)c0 * 49 + i];
    __syncthreads();
    float acc[CPW][TH];
    conv_chunk<TH, CPW, true>(halo, wsm, warp, lane, acc);
#pragma unroll
    for (int q = 0; q < CPW; ++q) {
      int gx = x0 + CPW * warp + q;
      if (gx >= W) continue;
#pragma unroll
      for (int r = 0; r < TH; ++r) {
        int gy = y0 + r;
        if (gy >= H) continue;
        int64_t off = ((n * H + gy) * W + gx) * C + c0 + lane;
        float v = acc[q][r];
        if (dres) v += to_f32(dres[off]);
        dx[off] = from_f32<TOUT>(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward-weights (+bias): persistent CTAs, each bound to one 32-channel chunk; 49 tap accumulators
// + 1 bias accumulator per thread live in registers across all the tiles the CTA visits, then the 8
// warps are summed through shared memory and one partial row [50, 32] is written per CTA.
// grid = rows * nchunks;  CTA b: chunk = b % nchunks, row = b / nchunks.
// ------------------------------------------------------------------------------------------------
template <typename TDY, typename TX, int TH, int CPW>
__global__ void __launch_bounds__(NWARP * 32) dwconv7_wgrad_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x,
                                                                    int64_t N, int64_t H, int64_t W, int64_t C,
                                                                    int tiles_x, int tiles_y, int rows,
                                                                    float* __restrict__ partial) {
  typedef DwTile<TH, CPW> T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* halo = reinterpret_cast<float*>(smem_raw);                 // x halo tile
  float* dsm = halo + T::HALO_FLOATS;                                // dy tile [TH][TW][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (int)(C / CH);
  const int chunk = blockIdx.x % nchunks, row = blockIdx.x / nchunks;
  const int c0 = chunk * CH;
  const int64_t tiles_per_img = (int64_t)tiles_x * tiles_y;
  const int64_t ntiles = N * tiles_per_img;

  float accw[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) accw[t] = 0.f;
  float accb = 0.f;

  for (int64_t tile = row; tile < ntiles; tile += rows) {
    const int64_t n = tile / tiles_per_img;
    const int tt = (int)(tile - n * tiles_per_img);
    const int ty = tt / tiles_x, tx = tt - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * T::TW;
    __syncthreads();
    load_halo<TX, T::HH, T::HW>(halo, x + n * H * W * C, H, W, C, y0, x0, c0);
    {
      const TDY* img = dy + n * H * W * C;
      const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
      for (int p = pl; p < TH * T::TW; p += 32) {
        int r = p / T::TW, cx = p - r * T::TW;
        int gy = y0 + r, gx = x0 + cx;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gy < H && gx < W) load4(img + ((int64_t)gy * W + gx) * C + c0 + sub * 4, v);
        *reinterpret_cast<float4*>(dsm + p * CH + sub * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    __syncthreads();
    float d[CPW][TH];
#pragma unroll
    for (int q = 0; q < CPW; ++q)
#pragma unroll
      for (int r = 0; r < TH; ++r) {
        d[q][r] = dsm[(r * T::TW + CPW * warp + q) * CH + lane];
        accb += d[q][r];
      }
#pragma unroll
    for (int j = 0; j < 6 + CPW; ++j) {
      float xc[TH + 6];
#pragma unroll
      for (int iy = 0; iy < TH + 6; ++iy) xc[iy] = halo[(iy * T::HW + CPW * warp + j) * CH + lane];
#pragma unroll
      for (int q = 0; q < CPW; ++q) {
        const int kx = j - q;
        if (kx >= 0 && kx <= 6) {
#pragma unroll
          for (int ky = 0; ky < 7; ++ky) {
            float s = accw[ky * 7 + kx];
#pragma unroll
            for (int r = 0; r < TH; ++r) s = fmaf(d[q][r], xc[r + ky], s);
            accw[ky * 7 + kx] = s;
          }
        }
      }
    }
  }
  // cross-warp reduction through shared memory (fixed order), then one partial row per CTA
  __syncthreads();
  float* red = halo;                       // [NWARP][50][32]
#pragma unroll
  for (int t = 0; t < 49; ++t) red[(warp * 50 + t) * CH + lane] = accw[t];
  red[(warp * 50 + 49) * CH + lane] = accb;
  __syncthreads();
  for (int i = threadIdx.x; i < 50 * CH; i += NWARP * 32) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < NWARP; ++wv) s += red[wv * 50 * CH + i];
    int t = i / CH, l = i - t * CH;
    partial[((int64_t)row * 50 + t) * C + c0 + l] = s;
  }
}

__global__ void __launch_bounds__(256) dwconv7_wgrad_finalize_kernel(const float* __restrict__ partial, int P, int64_t C,
                                                                     int accumulate, float* __restrict__ dw,
                                                                     float* __restrict__ db) {
  // one thread per (tap t, channel c), reading partial[p][t][c] coalesced over c
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
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

// ---- host-side configuration ------------------------------------------------------------------
struct DwCfg { int th, cpw; };

static inline size_t fwd_smem(int th, int cpw, int64_t C, int act_size) {
  int tw = NWARP * cpw;
  return (size_t)((th + 6) * (tw + 6) * CH + CH * 49 + th * tw * 2) * 4 + (size_t)th * tw * C * act_size;
}
static inline size_t dgrad_smem(int th, int cpw) {
  int tw = NWARP * cpw;
  return (size_t)((th + 6) * (tw + 6) * CH + CH * 49) * 4;
}
static inline size_t wgrad_smem(int th, int cpw) {
  int tw = NWARP * cpw;
  size_t a = (size_t)((th + 6) * (tw + 6) * CH + th * tw * CH) * 4;
  size_t b = (size_t)NWARP * 50 * CH * 4;
  return a > b ? a : b;
}
constexpr size_t kMaxSmem = 227 * 1024;

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

template <typename TIN, typename TACT, int TH, int CPW>
static int launch_fwd(const void* x, const float* w, const float* bias, const float* ln_w, const float* ln_b, float eps,
                      int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn, float* mean, float* rstd,
                      cudaStream_t s) {
  constexpr int TW = NWARP * CPW;
  int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
  size_t smem = fwd_smem(TH, CPW, C, sizeof(TACT));
  auto k = dwconv7_ln_fwd_kernel<TIN, TACT, TH, CPW>;
  if (int rc = set_smem(k, smem)) return rc;
  dim3 grid(tiles_x * tiles_y, (unsigned)N);
  k<<<grid, NWARP * 32, smem, s>>>((const TIN*)x, w, bias, ln_w, ln_b, eps, H, W, C, tiles_x, (TACT*)y, (TACT*)xn,
                                   mean, rstd);
  return check_launch("dwconv7_ln_fwd");
}

template <typename TIN, typename TACT>
static int pick_fwd(const void* x, const float* w, const float* bias, const float* ln_w, const float* ln_b, float eps,
                    int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn, float* mean, float* rstd,
                    cudaStream_t s) {
  const int es = sizeof(TACT);
  // widest tile that fits in shared memory and is not mostly padding for this feature map
  if (W > 8 && fwd_smem(8, 2, C, es) <= kMaxSmem)
    return launch_fwd<TIN, TACT, 8, 2>(x, w, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, mean, rstd, s);
  if (H > 4 && fwd_smem(8, 1, C, es) <= kMaxSmem)
    return launch_fwd<TIN, TACT, 8, 1>(x, w, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, mean, rstd, s);
  if (fwd_smem(4, 1, C, es) <= kMaxSmem)
    return launch_fwd<TIN, TACT, 4, 1>(x, w, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, mean, rstd, s);
  set_error("dwconv7_ln_fwd: C=%lld does not fit in shared memory", (long long)C);
  return CNX_E_SHAPE;
}

template <typename TIN, typename TOUT, int TH, int CPW>
static int launch_dgrad(const void* dy, const float* w, const void* dres, void* dx, int64_t N, int64_t H, int64_t W,
                        int64_t C, cudaStream_t s) {
  constexpr int TW = NWARP * CPW;
  int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
  size_t smem = dgrad_smem(TH, CPW);
  auto k = dwconv7_dgrad_kernel<TIN, TOUT, TH, CPW>;
  if (int rc = set_smem(k, smem)) return rc;
  dim3 grid(tiles_x * tiles_y, (unsigned)N);
  k<<<grid, NWARP * 32, smem, s>>>((const TIN*)dy, w, (const TOUT*)dres, H, W, C, tiles_x, (TOUT*)dx);
  return check_launch("dwconv7_dgrad");
}

template <typename TDY, typename TX, int TH, int CPW>
static int launch_wgrad(const void* dy, const void* x, int64_t N, int64_t H, int64_t W, int64_t C, float* partial,
                        int P, cudaStream_t s) {
  constexpr int TW = NWARP * CPW;
  int tiles_x = (int)((W + TW - 1) / TW), tiles_y = (int)((H + TH - 1) / TH);
  size_t smem = wgrad_smem(TH, CPW);
  auto k = dwconv7_wgrad_kernel<TDY, TX, TH, CPW>;
  if (int rc = set_smem(k, smem)) return rc;
  int nchunks = (int)(C / CH);
  k<<<(unsigned)(P * nchunks), NWARP * 32, smem, s>>>((const TDY*)dy, (const TX*)x, N, H, W, C, tiles_x, tiles_y, P,
                                                       partial);
  return check_launch("dwconv7_wgrad");
}

}  // namespace cnx

using namespace cnx;

extern "C" {

int cnx_dwconv7_ln_fwd(const void* x, int x_dtype, const float* w, const float* bias, const float* ln_w,
                       const float* ln_b, float eps, int64_t N, int64_t H, int64_t W, int64_t C, void* y,
                       void* xn, int act_dtype, float* mean, float* rstd, void* stream) {
  CNX_REQUIRE(x && w && bias && ln_w && ln_b && y && xn && mean && rstd, CNX_E_BADARG, "dwconv7_ln_fwd: null pointer");
  CNX_REQUIRE(dtype_ok(x_dtype) && dtype_ok(act_dtype), CNX_E_BADARG, "dwconv7_ln_fwd: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && N < 65536, CNX_E_BADARG, "dwconv7_ln_fwd: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_ln_fwd: C=%lld must be a multiple of 32", (long long)C);
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == CNX_F32 && act_dtype == CNX_F32) return pick_fwd<float, float>(x, w, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, mean, rstd, s);
  if (x_dtype == CNX_F32 && act_dtype == CNX_BF16) return pick_fwd<float, bf16>(x, w, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, mean, rstd, s);
  if (x_dtype == CNX_BF16 && act_dtype == CNX_BF16) return pick_fwd<bf16, bf16>(x, w, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, mean, rstd, s);
  set_error("dwconv7_ln_fwd: bf16 stream with fp32 activations is not a supported combination");
  return CNX_E_BADARG;
}

int cnx_dwconv7_dgrad(const void* dy, int dy_dtype, const float* w, const void* dres, void* dx, int stream_dtype,
                      int64_t N, int64_t H, int64_t W, int64_t C, void* stream) {
  CNX_REQUIRE(dy && w && dx, CNX_E_BADARG, "dwconv7_dgrad: null pointer");
  CNX_REQUIRE(dtype_ok(dy_dtype) && dtype_ok(stream_dtype), CNX_E_BADARG, "dwconv7_dgrad: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && N < 65536, CNX_E_BADARG, "dwconv7_dgrad: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_dgrad: C=%lld must be a multiple of 32", (long long)C);
  cudaStream_t s = (cudaStream_t)stream;
  const bool wide = W > 8;
  const bool tall = H > 4;
#define CNX_DG(TI, TO)                                                                              \
  do {                                                                                              \
    if (wide) return launch_dgrad<TI, TO, 8, 2>(dy, w, dres, dx, N, H, W, C, s);                    \
    if (tall) return launch_dgrad<TI, TO, 8, 1>(dy, w, dres, dx, N, H, W, C, s);                    \
    return launch_dgrad<TI, TO, 4, 1>(dy, w, dres, dx, N, H, W, C, s);                              \
  } while (0)
  if (dy_dtype == CNX_F32 && stream_dtype == CNX_F32) CNX_DG(float, float);
  if (dy_dtype == CNX_BF16 && stream_dtype == CNX_F32) CNX_DG(bf16, float);
  if (dy_dtype == CNX_BF16 && stream_dtype == CNX_BF16) CNX_DG(bf16, bf16);
#undef CNX_DG
  set_error("dwconv7_dgrad: fp32 activations with a bf16 stream is not a supported combination");
  return CNX_E_BADARG;
}

int cnx_dwconv7_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W,
                      int64_t C, float* partial, int P, void* stream) {
  CNX_REQUIRE(dy && x && partial && P > 0, CNX_E_BADARG, "dwconv7_wgrad: bad argument");
  CNX_REQUIRE(dtype_ok(dy_dtype) && dtype_ok(x_dtype), CNX_E_BADARG, "dwconv7_wgrad: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_wgrad: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_wgrad: C=%lld must be a multiple of 32", (long long)C);
  cudaStream_t s = (cudaStream_t)stream;
  const bool wide = W > 8;
#define CNX_WG(TD, TX)                                                                 \
  do {                                                                                 \
    if (wide) return launch_wgrad<TD, TX, 8, 2>(dy, x, N, H, W, C, partial, P, s);     \
    return launch_wgrad<TD, TX, 8, 1>(dy, x, N, H, W, C, partial, P, s);               \
  } while (0)
  if (dy_dtype == CNX_F32 && x_dtype == CNX_F32) CNX_WG(float, float);
  if (dy_dtype == CNX_BF16 && x_dtype == CNX_F32) CNX_WG(bf16, float);
  if (dy_dtype == CNX_BF16 && x_dtype == CNX_BF16) CNX_WG(bf16, bf16);
#undef CNX_WG
  set_error("dwconv7_wgrad: unsupported dtype combination");
  return CNX_E_BADARG;
}

int cnx_dwconv7_wgrad_finalize(const float* partial, int P, int64_t C, int accumulate, float* dw, float* db,
                               void* stream) {
  CNX_REQUIRE(partial && dw && db && P > 0 && C > 0, CNX_E_BADARG, "dwconv7_wgrad_finalize: bad argument");
  dwconv7_wgrad_finalize_kernel<<<(unsigned)((50 * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, P, C,
                                                                                                   accumulate, dw, db);
  return check_launch("dwconv7_wgrad_finalize");
}

}  // extern "C"

// dwconv.cu — a1/a2: depthwise 7x7 conv (+fused channels-last LayerNorm) forward, backward-data
// (+residual-gradient add) and backward-weights, channels-last.
//
// The 7x7 depthwise conv is 98 flop per 8-12 bytes: right at the fp32 ridge of a B200, so the kernels are built
// to keep BOTH the FMA pipe and HBM busy:
//   * halo tiles arrive by TMA (cp.async.bulk.tensor.4d over a {C, W, H, N} tensor map, box {32 ch, TW+6, TH+6,
//     NB}); out-of-bounds box elements are zero-filled by the TMA unit, which IS the conv's padding=3 — no
//     per-element address math, no bounds branches;
//   * 2-stage mbarrier pipeline, persistent CTAs: the load of (tile, chunk) g+1 overlaps the FMAs of g, across
//     tile boundaries;
//   * a half-warp owns 32 channels as 16 channel PAIRS, so every shared-memory access is a conflict-free 8-byte
//     (fp32) / 4-byte (bf16) row and every FMA is the packed fma.rn.f32x2 (2 channels per instruction);
//   * each thread keeps the 49 tap pairs of its channels and a CPW x TH output strip in registers: an input
//     value read once from shared memory feeds up to 7*CPW packed FMAs.
// Forward writes y (conv + bias, rounded to the activation dtype — what autocast hands to layer_norm) straight
// from registers, then the CTA normalises its own tile from L2 (one warp per pixel, two-pass fp32 statistics)
// and writes xn, mean, rstd.  Weights are consumed tap-major ([49][C], cnx_dwconv7_weight_prep).
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace cnx {
namespace dw {

constexpr int CH = 32;          // channels per chunk (16 pairs per half-warp)
constexpr int NTHREADS = 256;   // 8 warps = 16 half-warp workers
constexpr int W_BYTES = 49 * CH * 4;          // one chunk of tap-major weights
constexpr int W_STRIDE = 6400;                // padded to a multiple of 128 B

enum { MODE_FWD = 0, MODE_DGRAD = 1 };

template <int TH_, int CPW_, int WX_, bool IMG_>
struct Geo {
  static constexpr int TH = TH_, CPW = CPW_, WX = WX_;
  static constexpr bool IMG = IMG_;
  static constexpr int WY = 16 / WX;
  static constexpr int TW = WX * CPW;                 // output columns per tile
  static constexpr int ROWS = IMG ? TH : WY * TH;     // output rows per tile (per image)
  static constexpr int NB = IMG ? WY : 1;             // images per tile
  static constexpr int HH = ROWS + 6, HW = TW + 6;
  static constexpr int HALO_ELEMS = NB * HH * HW * CH;
  static constexpr int TILE_ELEMS = NB * ROWS * TW * CH;
};

// ---- PTX wrappers (same idioms as gemm_tc.cu) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// channel-pair loads from a [pixel][32 ch] shared-memory tile
__device__ __forceinline__ float2 ld_pair(const float* sm, int idx) { return *reinterpret_cast<const float2*>(sm + idx); }
__device__ __forceinline__ float2 ld_pair(const bf16* sm, int idx) {
  uint32_t u = *reinterpret_cast<const uint32_t*>(sm + idx);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void st_pair(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st_pair(bf16* p, float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162*>(p) = h;
}
__device__ __forceinline__ float2 ldg_pair(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ldg_pair(const bf16* p) {
  uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

struct TileCoord { int n0, y0, x0; };

template <class G>
__device__ __forceinline__ TileCoord decode_tile(int tile, int tiles_x, int tiles_y) {
  TileCoord t;
  int tx = tile % tiles_x;
  int r = tile / tiles_x;
  int ty = r % tiles_y;
  t.n0 = (r / tiles_y) * G::NB;
  t.y0 = ty * G::ROWS;
  t.x0 = tx * G::TW;
  return t;
}

// The register-tiled 7x7 correlation for one 32-channel chunk: acc[q][r] for output column (col0+q), row (row0+r).
template <class G, typename TS, bool FLIP>
__device__ __forceinline__ void conv_chunk(const TS* __restrict__ halo, const float* __restrict__ wsm, int base, int cp,
                                           float2 (&acc)[G::CPW][G::TH]) {
  float2 wr[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wr[t] = *reinterpret_cast<const float2*>(wsm + (FLIP ? 48 - t : t) * CH + 2 * cp);
#pragma unroll
  for (int q = 0; q < G::CPW; ++q)
#pragma unroll
    for (int r = 0; r < G::TH; ++r) acc[q][r] = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 6 + G::CPW; ++j) {
#pragma unroll
    for (int iy = 0; iy < G::TH + 6; ++iy) {
      const float2 v = ld_pair(halo, base + (iy * G::HW + j) * CH);
#pragma unroll
      for (int q = 0; q < G::CPW; ++q) {
        const int kx = j - q;
        if (kx >= 0 && kx <= 6) {
#pragma unroll
          for (int ky = 0; ky < 7; ++ky) {
            const int r = iy - ky;
            if (r >= 0 && r < G::TH) acc[q][r] = __ffma2_rn(v, wr[ky * 7 + kx], acc[q][r]);
          }
        }
      }
    }
  }
}


// LayerNorm over C for the pixels of one tile, rows held in registers: one warp per pixel, U pixels in flight per
// warp so that a single L2 round trip covers U rows.  NV = 128-channel steps per row (C <= 128*NV).
template <class G, int NV, int U, typename TOUT>
__device__ __forceinline__ void ln_tile_regs(const TileCoord& t, int N, int H, int W, int C, const TOUT* __restrict__ y,
                                             const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                                             TOUT* __restrict__ xn, float* __restrict__ mean_out,
                                             float* __restrict__ rstd_out, int warp, int lane) {
  constexpr int P = G::NB * G::ROWS * G::TW;
  const int nvec = C >> 2;
  float lw[NV][4], lb[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = lane + 32 * j;
    if (v < nvec) { load4(ln_w + v * 4, lw[j]); load4(ln_b + v * 4, lb[j]); }
  }
  const float invC = 1.0f / (float)C;
  for (int p0 = warp * U; p0 < P; p0 += 8 * U) {
    float a[U][NV][4];
    int64_t moff[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u;
      const int b = p / (G::ROWS * G::TW), rr = (p / G::TW) % G::ROWS, cx = p % G::TW;
      const int nn = t.n0 + b, gy = t.y0 + rr, gx = t.x0 + cx;
      ok[u] = (p < P) && nn < N && gy < H && gx < W;
      moff[u] = ok[u] ? (((int64_t)nn * H + gy) * W + gx) : 0;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int v = lane + 32 * j;
        if (ok[u] && v < nvec) load4(y + moff[u] * C + v * 4, a[u][j]);
        else { a[u][j][0] = a[u][j][1] = a[u][j][2] = a[u][j][3] = 0.f; }
      }
    }
    float mu[U], rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s1 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) s1 += (a[u][j][0] + a[u][j][1]) + (a[u][j][2] + a[u][j][3]);
      mu[u] = warp_sum(s1) * invC;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (lane + 32 * j < nvec) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float d = a[u][j][e] - mu[u]; s2 = fmaf(d, d, s2); }
        }
      }
      rs[u] = rsqrtf(warp_sum(s2) * invC + eps);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int v = lane + 32 * j;
        if (v < nvec) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = fmaf((a[u][j][e] - mu[u]) * rs[u], lw[j][e], lb[j][e]);
          store4(xn + moff[u] * C + v * 4, o);
        }
      }
      if (lane == 0) { mean_out[moff[u]] = mu[u]; rstd_out[moff[u]] = rs[u]; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward (MODE_FWD):  y = conv(x) + bias -> TOUT ; then per-tile LayerNorm -> xn, mean, rstd
// dgrad   (MODE_DGRAD): dx = dres + conv_flipped(dy) -> TOUT
// ------------------------------------------------------------------------------------------------
template <class G, int MODE, typename TIN, typename TOUT>
#ifndef CNX_DW_MINB
#define CNX_DW_MINB 1
#endif
__global__ void __launch_bounds__(NTHREADS, (MODE == MODE_DGRAD ? CNX_DW_MINB : 1))
dwconv7_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int N, int H, int W, int C,
               int tiles_x, int tiles_y, int ntiles, const float* __restrict__ bias, const TOUT* __restrict__ dres,
               TOUT* __restrict__ out, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
               TOUT* __restrict__ xn, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int HALO_BYTES = G::HALO_ELEMS * (int)sizeof(TIN);
  constexpr int STAGE_BYTES = ((HALO_BYTES + 127) / 128) * 128 + W_STRIDE;
  const uint32_t base_u = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* base_p = smem_raw + (base_u - smem_u32(smem_raw));
  const uint32_t bars = base_u + 2 * STAGE_BYTES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int worker = warp * 2 + (lane >> 4), cp = lane & 15;
  const int wx = worker % G::WX, wy = worker / G::WX;
  const int row0 = G::IMG ? 0 : wy * G::TH;
  const int img = G::IMG ? wy : 0;
  const int col0 = wx * G::CPW;
  const int hbase = ((img * G::HH + row0) * G::HW + col0) * CH + 2 * cp;

  const int nchunks = C / CH;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total = my_tiles * nchunks;

  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int g) {
    const int i = g / nchunks, k = g - i * nchunks;
    const TileCoord t = decode_tile<G>((int)blockIdx.x + i * (int)gridDim.x, tiles_x, tiles_y);
    const int s = g & 1;
    const uint32_t dst = base_u + s * STAGE_BYTES;
    mbar_expect_tx(bars + 8 * s, HALO_BYTES + W_BYTES);
    tma_load_4d(dst, &tmX, bars + 8 * s, k * CH, t.x0 - 3, t.y0 - 3, t.n0);
    tma_load_2d(dst + STAGE_BYTES - W_STRIDE, &tmW, bars + 8 * s, k * CH, 0);
  };
  if (tid == 0) {
    if (total > 0) issue(0);
    if (total > 1) issue(1);
  }

  for (int g = 0; g < total; ++g) {
    const int i = g / nchunks, k = g - i * nchunks;
    const TileCoord t = decode_tile<G>((int)blockIdx.x + i * (int)gridDim.x, tiles_x, tiles_y);
    const int s = g & 1;
    mbar_wait(bars + 8 * s, (uint32_t)((g >> 1) & 1));
    const TIN* halo = reinterpret_cast<const TIN*>(base_p + s * STAGE_BYTES);
    const float* wsm = reinterpret_cast<const float*>(base_p + s * STAGE_BYTES + STAGE_BYTES - W_STRIDE);
    const int c = k * CH + 2 * cp;
    const int n = t.n0 + img;
    float2 b2 = make_float2(0.f, 0.f);
    if (MODE == MODE_FWD) b2 = __ldg(reinterpret_cast<const float2*>(bias + c));
    // dgrad: the residual-gradient values are fetched BEFORE the FMAs (clamped addresses, no branches), so their
    // DRAM latency hides under the 784 packed FMAs instead of serialising 16 load->add->store chains after them
    float2 dr[G::CPW][G::TH];
    if (MODE == MODE_DGRAD && dres != nullptr) {
      const int nc = n < N ? n : N - 1;
#pragma unroll
      for (int q = 0; q < G::CPW; ++q) {
        int gx = t.x0 + col0 + q;
        gx = gx < W ? gx : W - 1;
#pragma unroll
        for (int r = 0; r < G::TH; ++r) {
          int gy = t.y0 + row0 + r;
          gy = gy < H ? gy : H - 1;
          dr[q][r] = ldg_pair(dres + (((int64_t)nc * H + gy) * W + gx) * C + c);
        }
      }
    }
    float2 acc[G::CPW][G::TH];
    conv_chunk<G, TIN, MODE == MODE_DGRAD>(halo, wsm, hbase, cp, acc);
    if (n < N) {
#pragma unroll
      for (int q = 0; q < G::CPW; ++q) {
        const int gx = t.x0 + col0 + q;
        if (gx < W) {
#pragma unroll
          for (int r = 0; r < G::TH; ++r) {
            const int gy = t.y0 + row0 + r;
            if (gy < H) {
              const int64_t off = (((int64_t)n * H + gy) * W + gx) * C + c;
              float2 v = acc[q][r];
              if (MODE == MODE_FWD) { v.x += b2.x; v.y += b2.y; }
              else if (dres != nullptr) { v.x += dr[q][r].x; v.y += dr[q][r].y; }
              st_pair(out + off, v);
            }
          }
        }
      }
    }
    __syncthreads();                     // every warp is done with stage s (and y of this chunk is written)
    if (tid == 0 && g + 2 < total) issue(g + 2);

#ifndef CNX_DW_NOLN
#define CNX_DW_NOLN 0
#endif
    if (MODE == MODE_FWD && k == nchunks - 1 && !CNX_DW_NOLN) {
      // ---- LayerNorm over C for the pixels of this tile (rows are L2-resident: this CTA just wrote them) ----
      if (C <= 128) ln_tile_regs<G, 1, 8, TOUT>(t, N, H, W, C, out, ln_w, ln_b, eps, xn, mean_out, rstd_out, warp, lane);
      else if (C <= 256) ln_tile_regs<G, 2, 8, TOUT>(t, N, H, W, C, out, ln_w, ln_b, eps, xn, mean_out, rstd_out, warp, lane);
      else if (C <= 512) ln_tile_regs<G, 4, 4, TOUT>(t, N, H, W, C, out, ln_w, ln_b, eps, xn, mean_out, rstd_out, warp, lane);
      else if (C <= 1024) ln_tile_regs<G, 8, 2, TOUT>(t, N, H, W, C, out, ln_w, ln_b, eps, xn, mean_out, rstd_out, warp, lane);
      else ln_tile_regs<G, 16, 1, TOUT>(t, N, H, W, C, out, ln_w, ln_b, eps, xn, mean_out, rstd_out, warp, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward-weights (+bias): persistent CTAs, each bound to one 32-channel chunk; 49 tap-pair accumulators
// + 1 bias pair per thread live in registers across all the tiles the CTA visits (x halo tile and dy tile
// arrive by TMA, double-buffered), then the 16 workers are summed through shared memory in a fixed order
// and one partial row [50, 32] is written per CTA.   grid = rows * nchunks: chunk = b % nchunks, row = b / nchunks.
// ------------------------------------------------------------------------------------------------
template <class G, typename TDY, typename TX>
__global__ void __launch_bounds__(NTHREADS, 1)
dwconv7_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, int N, int H, int W,
                     int C, int tiles_x, int tiles_y, int ntiles, int rows, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int HALO_BYTES = G::HALO_ELEMS * (int)sizeof(TX);
  constexpr int DY_BYTES = G::TILE_ELEMS * (int)sizeof(TDY);
  constexpr int HALO_PAD = ((HALO_BYTES + 127) / 128) * 128;
  constexpr int STAGE_BYTES = HALO_PAD + ((DY_BYTES + 127) / 128) * 128;
  const uint32_t base_u = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* base_p = smem_raw + (base_u - smem_u32(smem_raw));
  const uint32_t bars = base_u + 2 * STAGE_BYTES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int worker = warp * 2 + (lane >> 4), cp = lane & 15;
  const int wx = worker % G::WX, wy = worker / G::WX;
  const int row0 = G::IMG ? 0 : wy * G::TH;
  const int img = G::IMG ? wy : 0;
  const int col0 = wx * G::CPW;
  const int hbase = ((img * G::HH + row0) * G::HW + col0) * CH + 2 * cp;
  const int dbase = ((img * G::ROWS + row0) * G::TW + col0) * CH + 2 * cp;

  const int nchunks = C / CH;
  const int chunk = (int)blockIdx.x % nchunks, row = (int)blockIdx.x / nchunks;
  const int my_tiles = (row < ntiles) ? (ntiles - row + rows - 1) / rows : 0;

  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int g) {
    const TileCoord t = decode_tile<G>(row + g * rows, tiles_x, tiles_y);
    const int s = g & 1;
    const uint32_t dst = base_u + s * STAGE_BYTES;
    mbar_expect_tx(bars + 8 * s, HALO_BYTES + DY_BYTES);
    tma_load_4d(dst, &tmX, bars + 8 * s, chunk * CH, t.x0 - 3, t.y0 - 3, t.n0);
    tma_load_4d(dst + HALO_PAD, &tmDY, bars + 8 * s, chunk * CH, t.x0, t.y0, t.n0);
  };
  if (tid == 0) {
    if (my_tiles > 0) issue(0);
    if (my_tiles > 1) issue(1);
  }

  float2 accw[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) accw[t] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);

  for (int g = 0; g < my_tiles; ++g) {
    const int s = g & 1;
    mbar_wait(bars + 8 * s, (uint32_t)((g >> 1) & 1));
    const TX* halo = reinterpret_cast<const TX*>(base_p + s * STAGE_BYTES);
    const TDY* dsm = reinterpret_cast<const TDY*>(base_p + s * STAGE_BYTES + HALO_PAD);
    // dy outside the image is zero-filled by TMA, so no masking is needed
    float2 d[G::CPW][G::TH];
#pragma unroll
    for (int q = 0; q < G::CPW; ++q)
#pragma unroll
      for (int r = 0; r < G::TH; ++r) {
        d[q][r] = ld_pair(dsm, dbase + (r * G::TW + q) * CH);
        accb.x += d[q][r].x;
        accb.y += d[q][r].y;
      }
#pragma unroll
    for (int j = 0; j < 6 + G::CPW; ++j) {
#pragma unroll
      for (int iy = 0; iy < G::TH + 6; ++iy) {
        const float2 v = ld_pair(halo, hbase + (iy * G::HW + j) * CH);
#pragma unroll
        for (int q = 0; q < G::CPW; ++q) {
          const int kx = j - q;
          if (kx >= 0 && kx <= 6) {
#pragma unroll
            for (int ky = 0; ky < 7; ++ky) {
              const int r = iy - ky;
              if (r >= 0 && r < G::TH) accw[ky * 7 + kx] = __ffma2_rn(d[q][r], v, accw[ky * 7 + kx]);
            }
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0 && g + 2 < my_tiles) issue(g + 2);
  }
  // cross-worker reduction through shared memory (fixed order), then one partial row per CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(base_p);          // [16 workers][50][32]  = 102400 B (fits in the two stages)
#pragma unroll
  for (int t = 0; t < 49; ++t) *reinterpret_cast<float2*>(red + (worker * 50 + t) * CH + 2 * cp) = accw[t];
  *reinterpret_cast<float2*>(red + (worker * 50 + 49) * CH + 2 * cp) = accb;
  __syncthreads();
  for (int i = tid; i < 50 * CH; i += NTHREADS) {
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < 16; ++wv) sum += red[wv * 50 * CH + i];
    const int t = i / CH, l = i - t * CH;
    partial[((int64_t)row * 50 + t) * C + chunk * CH + l] = sum;
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

// w [C][49] -> wt [49][C]
__global__ void __launch_bounds__(256) weight_transpose_kernel(const float* __restrict__ w, int64_t C, float* __restrict__ wt) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= 49 * C) return;
  int64_t t = i / C, c = i - t * C;
  wt[i] = w[c * 49 + t];
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      (void)cudaGetLastError();
  });
  return fn;
}

// channels-last activation [N][H][W][C]: 4-D map {C, W, H, N}, box {32, bw, bh, bn}, no swizzle, OOB -> zeros
static int make_map_nhwc(CUtensorMap* map, const void* ptr, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, int bw,
                         int bh, int bn) {
  EncodeTiledFn enc = get_encode();
  CNX_REQUIRE(enc != nullptr, CNX_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  CNX_REQUIRE((((uintptr_t)ptr) & 15) == 0, CNX_E_SHAPE, "dwconv: activation pointer must be 16-byte aligned");
  const cuuint64_t e = (cuuint64_t)dtype_size(dtype);
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * e, (cuuint64_t)W * C * e, (cuuint64_t)H * W * C * e};
  cuuint32_t box[4] = {CH, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, dtype == CNX_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CNX_REQUIRE(r == CUDA_SUCCESS, CNX_E_DRIVER, "cuTensorMapEncodeTiled(nhwc) failed (%d) N=%lld H=%lld W=%lld C=%lld box=%dx%dx%d",
              (int)r, (long long)N, (long long)H, (long long)W, (long long)C, bw, bh, bn);
  return 0;
}
// tap-major weights [49][C] fp32: 2-D map {C, 49}, box {32, 49}
static int make_map_wt(CUtensorMap* map, const float* wt, int64_t C) {
  EncodeTiledFn enc = get_encode();
  CNX_REQUIRE(enc != nullptr, CNX_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  CNX_REQUIRE((((uintptr_t)wt) & 15) == 0, CNX_E_SHAPE, "dwconv: weight pointer must be 16-byte aligned");
  cuuint64_t gdim[2] = {(cuuint64_t)C, 49};
  cuuint64_t gstr[1] = {(cuuint64_t)C * 4};
  cuuint32_t box[2] = {CH, 49};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wt), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CNX_REQUIRE(r == CUDA_SUCCESS, CNX_E_DRIVER, "cuTensorMapEncodeTiled(wt) failed (%d) C=%lld", (int)r, (long long)C);
  return 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// tile geometries: A = wide feature maps, B = 9..16 wide, I = whole small images (<= 8x8), 2 per tile
typedef Geo<8, 2, 16, false> GeoA;   // 32 x 8
typedef Geo<8, 2, 8, false> GeoB;    // 16 x 16
typedef Geo<8, 1, 8, true> GeoI;     // 8 x 8 x 2 images

template <class G, int MODE, typename TIN, typename TOUT>
static int launch_conv(const void* x, int x_dtype, const float* wt, const float* bias, const void* dres, void* out,
                       const float* ln_w, const float* ln_b, float eps, void* xn, float* mean, float* rstd, int64_t N,
                       int64_t H, int64_t W, int64_t C, cudaStream_t s) {
  CUtensorMap tmX, tmW;
  if (int rc = make_map_nhwc(&tmX, x, x_dtype, N, H, W, C, G::HW, G::HH, G::NB)) return rc;
  if (int rc = make_map_wt(&tmW, wt, C)) return rc;
  constexpr size_t HALO_BYTES = (size_t)G::HALO_ELEMS * sizeof(TIN);
  constexpr size_t STAGE = ((HALO_BYTES + 127) / 128) * 128 + W_STRIDE;
  constexpr size_t SMEM = 2 * STAGE + 16 + 128;
  static_assert(SMEM <= 227 * 1024, "dwconv tile does not fit in shared memory");
  auto k = dwconv7_kernel<G, MODE, TIN, TOUT>;
  if (int rc = set_smem(k, SMEM)) return rc;
  const int tiles_x = (int)((W + G::TW - 1) / G::TW), tiles_y = (int)((H + G::ROWS - 1) / G::ROWS);
  const int64_t nt = (int64_t)tiles_x * tiles_y * ((N + G::NB - 1) / G::NB);
  CNX_REQUIRE(nt < (1ll << 30), CNX_E_SHAPE, "dwconv: too many tiles");
  int grid = sm_count() * (MODE == MODE_DGRAD ? CNX_DW_MINB : 1);
  if (grid > nt) grid = (int)nt;
  k<<<grid, NTHREADS, SMEM, s>>>(tmX, tmW, (int)N, (int)H, (int)W, (int)C, tiles_x, tiles_y, (int)nt, bias, (const TOUT*)dres,
                                 (TOUT*)out, ln_w, ln_b, eps, (TOUT*)xn, mean, rstd);
  return check_launch(MODE == MODE_FWD ? "dwconv7_ln_fwd" : "dwconv7_dgrad");
}

template <int MODE, typename TIN, typename TOUT>
static int pick_conv(const void* x, int x_dtype, const float* wt, const float* bias, const void* dres, void* out,
                     const float* ln_w, const float* ln_b, float eps, void* xn, float* mean, float* rstd, int64_t N,
                     int64_t H, int64_t W, int64_t C, cudaStream_t s) {
  if (W <= 8 && H <= 8)
    return launch_conv<GeoI, MODE, TIN, TOUT>(x, x_dtype, wt, bias, dres, out, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  if (W <= 16)
    return launch_conv<GeoB, MODE, TIN, TOUT>(x, x_dtype, wt, bias, dres, out, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  return launch_conv<GeoA, MODE, TIN, TOUT>(x, x_dtype, wt, bias, dres, out, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
}

template <class G, typename TDY, typename TX>
static int launch_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W, int64_t C,
                        float* partial, int P, cudaStream_t s) {
  CUtensorMap tmX, tmDY;
  if (int rc = make_map_nhwc(&tmX, x, x_dtype, N, H, W, C, G::HW, G::HH, G::NB)) return rc;
  if (int rc = make_map_nhwc(&tmDY, dy, dy_dtype, N, H, W, C, G::TW, G::ROWS, G::NB)) return rc;
  constexpr size_t HALO_PAD = (((size_t)G::HALO_ELEMS * sizeof(TX) + 127) / 128) * 128;
  constexpr size_t STAGE = HALO_PAD + (((size_t)G::TILE_ELEMS * sizeof(TDY) + 127) / 128) * 128;
  constexpr size_t RED = (size_t)16 * 50 * CH * 4;
  constexpr size_t SMEM = (2 * STAGE > RED ? 2 * STAGE : RED) + 16 + 128;
  static_assert(SMEM <= 227 * 1024, "dwconv wgrad tile does not fit in shared memory");
  auto k = dwconv7_wgrad_kernel<G, TDY, TX>;
  if (int rc = set_smem(k, SMEM)) return rc;
  const int tiles_x = (int)((W + G::TW - 1) / G::TW), tiles_y = (int)((H + G::ROWS - 1) / G::ROWS);
  const int64_t nt = (int64_t)tiles_x * tiles_y * ((N + G::NB - 1) / G::NB);
  CNX_REQUIRE(nt < (1ll << 30), CNX_E_SHAPE, "dwconv wgrad: too many tiles");
  const int nchunks = (int)(C / CH);
  k<<<(unsigned)(P * nchunks), NTHREADS, SMEM, s>>>(tmX, tmDY, (int)N, (int)H, (int)W, (int)C, tiles_x, tiles_y, (int)nt, P,
                                                     partial);
  return check_launch("dwconv7_wgrad");
}

}  // namespace dw
}  // namespace cnx

namespace cnx {
// second-generation kernels (dwconv2.cu); CNX_DW_V1=1 in the environment selects the first-generation ones below
int dwconv7_ln_fwd_v2(const void* x, int x_dtype, const float* wt, const float* bias, const float* ln_w, const float* ln_b,
                      float eps, int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn, int act_dtype, float* mean,
                      float* rstd, cudaStream_t s);
int dwconv7_dgrad_v2(const void* dy, int dy_dtype, const float* wt, const void* dres, void* dx, int stream_dtype, int64_t N,
                     int64_t H, int64_t W, int64_t C, cudaStream_t s);
int dwconv7_ln_fwd_x3_v2(const void* x, const float* wt, const float* bias, const float* ln_w, const float* ln_b, float eps,
                         int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn3, float* mean, float* rstd, cudaStream_t s);
int dwconv7_wgrad_v2(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W, int64_t C,
                     float* partial, int P, cudaStream_t s);
static bool dw_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_DW_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}
}  // namespace cnx

using namespace cnx;
using namespace cnx::dw;

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
  cudaStream_t s = (cudaStream_t)stream;
  if (!dw_v1()) return dwconv7_ln_fwd_v2(x, x_dtype, wt, bias, ln_w, ln_b, eps, N, H, W, C, y, xn, act_dtype, mean, rstd, s);
  if (x_dtype == CNX_F32 && act_dtype == CNX_F32)
    return pick_conv<MODE_FWD, float, float>(x, x_dtype, wt, bias, nullptr, y, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  if (x_dtype == CNX_F32 && act_dtype == CNX_BF16)
    return pick_conv<MODE_FWD, float, bf16>(x, x_dtype, wt, bias, nullptr, y, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  if (x_dtype == CNX_BF16 && act_dtype == CNX_BF16)
    return pick_conv<MODE_FWD, bf16, bf16>(x, x_dtype, wt, bias, nullptr, y, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  set_error("dwconv7_ln_fwd: bf16 stream with fp32 activations is not a supported combination");
  return CNX_E_BADARG;
}

int cnx_dwconv7_ln_fwd_x3(const float* x, const float* wt, const float* bias, const float* ln_w, const float* ln_b, float eps,
                          int64_t N, int64_t H, int64_t W, int64_t C, float* y_scratch, void* xn3, float* mean, float* rstd,
                          void* stream) {
  CNX_REQUIRE(x && wt && bias && ln_w && ln_b && y_scratch && xn3 && mean && rstd, CNX_E_BADARG, "dwconv7_ln_fwd_x3: null pointer");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_ln_fwd_x3: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_ln_fwd_x3: C=%lld must be a multiple of 32", (long long)C);
  return dwconv7_ln_fwd_x3_v2(x, wt, bias, ln_w, ln_b, eps, N, H, W, C, y_scratch, xn3, mean, rstd, (cudaStream_t)stream);
}

int cnx_dwconv7_dgrad(const void* dy, int dy_dtype, const float* wt, const void* dres, void* dx, int stream_dtype,
                      int64_t N, int64_t H, int64_t W, int64_t C, void* stream) {
  CNX_REQUIRE(dy && wt && dx, CNX_E_BADARG, "dwconv7_dgrad: null pointer");
  CNX_REQUIRE(dtype_ok(dy_dtype) && dtype_ok(stream_dtype), CNX_E_BADARG, "dwconv7_dgrad: bad dtype");
  CNX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, CNX_E_BADARG, "dwconv7_dgrad: bad shape");
  CNX_REQUIRE(C % 32 == 0, CNX_E_SHAPE, "dwconv7_dgrad: C=%lld must be a multiple of 32", (long long)C);
  cudaStream_t s = (cudaStream_t)stream;
  if (!dw_v1()) return dwconv7_dgrad_v2(dy, dy_dtype, wt, dres, dx, stream_dtype, N, H, W, C, s);
  if (dy_dtype == CNX_F32 && stream_dtype == CNX_F32)
    return pick_conv<MODE_DGRAD, float, float>(dy, dy_dtype, wt, nullptr, dres, dx, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, N, H, W, C, s);
  if (dy_dtype == CNX_BF16 && stream_dtype == CNX_F32)
    return pick_conv<MODE_DGRAD, bf16, float>(dy, dy_dtype, wt, nullptr, dres, dx, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, N, H, W, C, s);
  if (dy_dtype == CNX_BF16 && stream_dtype == CNX_BF16)
    return pick_conv<MODE_DGRAD, bf16, bf16>(dy, dy_dtype, wt, nullptr, dres, dx, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, N, H, W, C, s);
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
  if (!dw_v1()) return dwconv7_wgrad_v2(dy, dy_dtype, x, x_dtype, N, H, W, C, partial, P, s);
#define CNX_WG(TD, TX)                                                                                  \
  do {                                                                                                  \
    if (W <= 8 && H <= 8) return launch_wgrad<GeoI, TD, TX>(dy, dy_dtype, x, x_dtype, N, H, W, C, partial, P, s); \
    if (W <= 16) return launch_wgrad<GeoB, TD, TX>(dy, dy_dtype, x, x_dtype, N, H, W, C, partial, P, s);          \
    return launch_wgrad<GeoA, TD, TX>(dy, dy_dtype, x, x_dtype, N, H, W, C, partial, P, s);                       \
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

// dwconv2.cu — a1/a2, second generation: warp-specialised depthwise 7x7 conv (+ fused channels-last LayerNorm) forward and
// backward-data (+ residual-gradient add).  See dwconv_common.cuh for the shared design notes.
//
// One persistent CTA per SM, three warp roles:
//   * warp 0            producer: one elected lane issues every TMA load — the halo tile (cp.async.bulk.tensor.4d over a
//                       {C,W,H,N} map, out-of-bounds elements zero-filled == padding 3) and the chunk's tap-major weights —
//                       into a STAGES-deep mbarrier ring; it runs ahead across chunk AND tile boundaries.
//   * warps 1..NWC      compute: a half-warp owns a CPW x TH pixel strip x 32 channels (16 channel pairs), all FMAs are the
//                       packed fma.rn.f32x2.  A warp hands its ring slot back as soon as its FMAs are done (per-warp
//                       mbarrier arrive, no CTA-wide barrier), then adds the bias, rounds to the activation dtype, keeps the
//                       per-pixel sum / sum-of-squares of the ROUNDED values in registers across the chunks of the tile and
//                       writes y.  At the end of a tile a 4-step half-warp butterfly turns the 16 per-thread partials into
//                       one (mean, rstd) per lane.
//   * warps NWC+1..     LayerNorm (forward only): stream the tile's y rows back from L2 (this SM just wrote them), normalise
//                       with the statistics the compute warps left in shared memory and write xn — while the compute warps
//                       are already on the next tile.  The FMA pipe never waits for the normalisation pass.
// Tile geometries are exact for 56/28/14/7-pixel maps (no padded work, no bounds code: template EXACT); other sizes use
// the generic 8x32 / 16x16 / 8x8x2 tiles with bounds checks.
#include "dwconv_common.cuh"

namespace cnx {
namespace dw2 {

enum { MODE_FWD = 0, MODE_DGRAD = 1 };
// rows in flight per LayerNorm warp (narrow rows: several pixels per pass / wide rows: one pixel per pass): the LN half streams
// the tile back from L2 and is latency-bound, so its throughput is the bytes it keeps in flight
#ifndef CNX_DW_LNW32
#define CNX_DW_LNW32 8
#endif
#ifndef CNX_DW_LNU1
#define CNX_DW_LNU1 8
#endif
#ifndef CNX_DW_LNU2
#define CNX_DW_LNU2 4
#endif
// LayerNorm warps: 4 for bf16 activations; 8 for fp32 activations (twice the bytes to stream back and, for the split operand,
// 1.5x the bytes to write: with 4 warps the LN half, not the FMA half, set the kernel time)
// (only where 1 + NWC + 8 warps still leave 128 registers per thread: the exact 56 / 28 / 14 / 7-pixel geometries)
template <typename TOUT, int NWC> struct LnWarps { static constexpr int N = (sizeof(TOUT) == 4 && NWC <= 7) ? CNX_DW_LNW32 : 4; };


template <class G, int MODE, typename TIN, typename TOUT>
struct ConvCfg {
  static constexpr int NWC = G::NWC;
  static constexpr int NLN = (MODE == MODE_FWD) ? LnWarps<TOUT, G::NWC>::N : 0;
  static constexpr int NT = (1 + NWC + NLN) * 32;
  static constexpr int HALO_BYTES = G::HALO_ELEMS * (int)sizeof(TIN);
  static constexpr int HALO_PAD = round128(HALO_BYTES);
  static constexpr int STAGE_BYTES = HALO_PAD + W_STRIDE;
  static constexpr int STATS_BYTES = (MODE == MODE_FWD) ? 2 * round128(G::P * 8) : 0;
  static constexpr int MISC = STATS_BYTES + 256 /*barriers*/ + 128 /*alignment slack*/;
  static constexpr int S_MAX = (SMEM_MAX - MISC) / STAGE_BYTES;
  static constexpr int STAGES = S_MAX > 4 ? 4 : S_MAX;
  static constexpr int SMEM = STAGES * STAGE_BYTES + MISC;
  static_assert(STAGES >= 2, "dwconv tile does not leave room for a double-buffered ring");
};

__device__ __forceinline__ float2 ldg_pair(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ldg_pair(const bf16* p) {
  uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// value after rounding to the activation dtype (what autocast hands to layer_norm)
__device__ __forceinline__ float2 round_pair(float2 v, float*) { return v; }
__device__ __forceinline__ float2 round_pair(float2 v, bf16*) {
  uint32_t u;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v.y), "f"(v.x));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// coherent 16-byte global load (NOT the read-only path: the data was written by other warps of this CTA)
__device__ __forceinline__ uint4 ld_global_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

// 16 lanes x NPIX per-thread partial (sum, sumsq) pairs -> lane L of each half-warp ends with the totals of pixel slot L
template <int NPIX>
__device__ __forceinline__ float2 halfwarp_transpose_reduce(const float2 (&st)[NPIX], int lane) {
  float a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = i < NPIX ? st[i].x : 0.f;
    b[i] = i < NPIX ? st[i].y : 0.f;
  }
#pragma unroll
  for (int d = 8; d >= 1; d >>= 1) {
    const bool up = (lane & d) != 0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
      const float sa = up ? a[j] : a[j + d], ka = up ? a[j + d] : a[j];
      const float sb = up ? b[j] : b[j + d], kb = up ? b[j + d] : b[j];
      a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, d);
      b[j] = kb + __shfl_xor_sync(0xffffffffu, sb, d);
    }
  }
  return make_float2(a[0], b[0]);
}

template <typename TOUT> struct Vec16;
template <> struct Vec16<bf16> { static constexpr int N = 8; };
template <> struct Vec16<float> { static constexpr int N = 4; };

__device__ __forceinline__ void unpack16(const uint4& r, float (&v)[8], bf16*) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}
__device__ __forceinline__ void unpack16(const uint4& r, float (&v)[4], float*) {
  v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
}
__device__ __forceinline__ uint4 pack16(const float (&v)[8], bf16*) {
  uint4 r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(v[1]), "f"(v[0]));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(v[3]), "f"(v[2]));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.z) : "f"(v[5]), "f"(v[4]));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.w) : "f"(v[7]), "f"(v[6]));
  return r;
}
__device__ __forceinline__ uint4 pack16(const float (&v)[4], float*) {
  return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}

// LayerNorm half: one lane item = 8 consecutive channels = one 16-byte vector of bf16 or two of fp32 (fp32 activations used to go
// 4 channels per item: twice the memory instructions per element and 8-byte split stores; the LayerNorm half sets the time of the
// fp32-activation forward)
template <typename TOUT> struct LnIO;
template <> struct LnIO<bf16> { static constexpr int NV = 1; };
template <> struct LnIO<float> { static constexpr int NV = 2; };
__device__ __forceinline__ void ln_load8(const bf16* p, uint4 (&raw)[1]) { raw[0] = ld_global_16(p); }
__device__ __forceinline__ void ln_load8(const float* p, uint4 (&raw)[2]) { raw[0] = ld_global_16(p); raw[1] = ld_global_16(p + 4); }
__device__ __forceinline__ void ln_unpack8(const uint4 (&raw)[1], float (&v)[8], bf16* t) { unpack16(raw[0], v, t); }
__device__ __forceinline__ void ln_unpack8(const uint4 (&raw)[2], float (&v)[8], float* t) {
  unpack16(raw[0], *reinterpret_cast<float(*)[4]>(&v[0]), t);
  unpack16(raw[1], *reinterpret_cast<float(*)[4]>(&v[4]), t);
}
__device__ __forceinline__ void ln_store8(bf16* p, const float (&v)[8]) { *reinterpret_cast<uint4*>(p) = pack16(v, (bf16*)nullptr); }
__device__ __forceinline__ void ln_store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = pack16(*reinterpret_cast<const float(*)[4]>(&v[0]), (float*)nullptr);
  *reinterpret_cast<uint4*>(p + 4) = pack16(*reinterpret_cast<const float(*)[4]>(&v[4]), (float*)nullptr);
}
// fp32-accurate forward ("x3", include/cnx.h): the normalised row leaves as the A-side split operand [hi | mid (| hi)] (bf16,
// row stride 2C or 3C; in the 2-segment form the consuming GEMM's K loop wraps instead) — 8 values per lane, 16-byte stores
__device__ __forceinline__ void store_split8(bf16* row3, int C, int col, const float (&x)[8], bool third) {
  uint4 hi;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.x) : "f"(x[1]), "f"(x[0]));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.y) : "f"(x[3]), "f"(x[2]));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.z) : "f"(x[5]), "f"(x[4]));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.w) : "f"(x[7]), "f"(x[6]));
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
  float r[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r[2 * i] = x[2 * i] - __uint_as_float(h[i] << 16);
    r[2 * i + 1] = x[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
  }
  *reinterpret_cast<uint4*>(row3 + col) = hi;
  *reinterpret_cast<uint4*>(row3 + C + col) = pack16(r, (bf16*)nullptr);
  if (third) *reinterpret_cast<uint4*>(row3 + 2 * C + col) = hi;
}

// tile pixel slot p (row-major over [NB][ROWS][TW]) -> global pixel index, or -1 outside the tensor
template <class G, bool EXACT>
__device__ __forceinline__ int64_t tile_pixel(const TileCoord& t, int p, int N, int H, int W) {
  const int b = p / (G::ROWS * G::TW), rr = (p / G::TW) % G::ROWS, cx = p % G::TW;
  const int nn = t.n0 + b, gy = t.y0 + rr, gx = t.x0 + cx;
  if (!EXACT && (nn >= N || gy >= H || gx >= W)) return -1;
  return ((int64_t)nn * H + gy) * W + gx;
}

// ------------------------------------------------------------------------------------------------
// forward (MODE_FWD):  y = conv(x) + bias -> TOUT ; LayerNorm over C -> xn, mean, rstd
// dgrad   (MODE_DGRAD): dx = dres + conv_flipped(dy) -> TOUT; with bf16 activations optionally also the operand copy the
//                       UPSTREAM Block's backward needs of this gradient, dz_up = bf16(dp_up[n] * dx) (its cnx_grad_prep pass,
//                       a 6-byte-per-element round trip, folded into this epilogue).  In this mode `xn3` carries dz_up and `ln_w`
//                       carries dp_up (per-sample drop-path scale of the upstream Block, or null).
// ------------------------------------------------------------------------------------------------
template <class G, int MODE, typename TIN, typename TOUT, bool EXACT>
__global__ void __launch_bounds__(ConvCfg<G, MODE, TIN, TOUT>::NT, 1)
dwconv7_v2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int N, int H, int W, int C,
                  int tiles_x, int tiles_y, int ntiles, const float* __restrict__ bias, const TOUT* __restrict__ dres,
                  TOUT* out, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, TOUT* __restrict__ xn,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, bf16* __restrict__ xn3, int xseg) {
  typedef ConvCfg<G, MODE, TIN, TOUT> Cfg;
  constexpr int STAGES = Cfg::STAGES, NWC = Cfg::NWC, NLN = Cfg::NLN;
  constexpr int NPIX = G::CPW * G::TH;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* base_p = smem_raw + (base_u - smem_u32(smem_raw));
  float2* stats = reinterpret_cast<float2*>(base_p + STAGES * Cfg::STAGE_BYTES);     // [2][round(P)] (mean, rstd)
  constexpr int STATS_STRIDE = round128(G::P * 8) / 8;
  const uint32_t bars = base_u + STAGES * Cfg::STAGE_BYTES + Cfg::STATS_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto lnfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + b); };
  auto lnempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = C / CH;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), NWC); }
    for (int b = 0; b < 2; ++b) { mbar_init(lnfull_bar(b), NWC); mbar_init(lnempty_bar(b), NLN > 0 ? NLN : 1); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();                                  // everything above overlapped the previous kernel's tail (common.cuh)

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const TileCoord t = decode_tile<G>((int)blockIdx.x + i * (int)gridDim.x, tiles_x, tiles_y);
        for (int k = 0; k < nchunks; ++k) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t dst = base_u + s * Cfg::STAGE_BYTES;
          mbar_expect_tx(full_bar(s), Cfg::HALO_BYTES + W_BYTES);
          tma_load_4d(dst, &tmX, full_bar(s), k * CH, t.x0 - 3, t.y0 - 3, t.n0);
          tma_load_2d(dst + Cfg::HALO_PAD, &tmW, full_bar(s), k * CH, 0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp <= NWC) {
    // ===== compute =====
    const int ct = tid - 32;
    const int worker = ct >> 4, cp = ct & 15;
    const int wx = worker % G::WX, tt = worker / G::WX;
    const int wy = tt % G::WY, img = tt / G::WY;
    const int row0 = wy * G::TH, col0 = wx * G::CPW;
    const int hbase = ((img * G::HH + row0) * G::HW + col0) * CH + 2 * cp;
    // byte offsets of this thread's NPIX pixels relative to its strip origin: tile-independent, computed once
    uint32_t poff[NPIX];
#pragma unroll
    for (int q = 0; q < G::CPW; ++q)
#pragma unroll
      for (int r = 0; r < G::TH; ++r) poff[q * G::TH + r] = (uint32_t)((r * W + q) * C) * (uint32_t)sizeof(TOUT);
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const TileCoord t = decode_tile<G>((int)blockIdx.x + i * (int)gridDim.x, tiles_x, tiles_y);
      const int n = t.n0 + img;
      const int gy0 = t.y0 + row0, gx0 = t.x0 + col0;
      const bool n_ok = EXACT || n < N;
      const int64_t pix0 = ((int64_t)(n_ok ? n : 0) * H + gy0) * W + gx0;
      float2 st[NPIX];
#pragma unroll
      for (int p = 0; p < NPIX; ++p) st[p] = make_float2(0.f, 0.f);
      const bool want_dz = (MODE == MODE_DGRAD) && sizeof(TIN) == 2 && xn3 != nullptr;
      float sc2 = 1.0f;
      if (want_dz && ln_w != nullptr) sc2 = __ldg(ln_w + (n_ok ? n : 0));
      for (int k = 0; k < nchunks; ++k) {
        const int c = k * CH + 2 * cp;
        float2 b2 = make_float2(0.f, 0.f);
        if (MODE == MODE_FWD) b2 = __ldg(reinterpret_cast<const float2*>(bias + c));
        uint8_t* op = reinterpret_cast<uint8_t*>(out + pix0 * C + c);
        // dgrad: the residual-gradient values are fetched BEFORE the FMAs so their latency hides under them
        float2 dr[G::CPW][G::TH];
        if (MODE == MODE_DGRAD) {
          const uint8_t* dp = reinterpret_cast<const uint8_t*>(dres + pix0 * C + c);
#pragma unroll
          for (int q = 0; q < G::CPW; ++q)
#pragma unroll
            for (int r = 0; r < G::TH; ++r) {
              dr[q][r] = make_float2(0.f, 0.f);
              if (dres != nullptr && (EXACT || (n_ok && gx0 + q < W && gy0 + r < H)))
                dr[q][r] = ldg_pair(reinterpret_cast<const TOUT*>(dp + poff[q * G::TH + r]));
            }
        }
        mbar_wait(full_bar(s), ph);
        const TIN* halo = reinterpret_cast<const TIN*>(base_p + s * Cfg::STAGE_BYTES);
        const float* wsm = reinterpret_cast<const float*>(base_p + s * Cfg::STAGE_BYTES + Cfg::HALO_PAD);
        float2 acc[G::CPW][G::TH];
        conv_chunk<G, TIN, MODE == MODE_DGRAD>(halo, wsm, hbase, cp, acc, b2);     // accumulators start at the bias
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));          // this warp is done with the slot
        if (++s == STAGES) { s = 0; ph ^= 1; }
#pragma unroll
        for (int q = 0; q < G::CPW; ++q) {
#pragma unroll
          for (int r = 0; r < G::TH; ++r) {
            float2 v = acc[q][r];
            if (MODE == MODE_FWD) {
              v = round_pair(v, (TOUT*)nullptr);
              // (sum, sum of squares) of the rounded pair in two packed FMAs
              float2& sp = st[q * G::TH + r];
              sp = __ffma2_rn(make_float2(v.x, v.x), make_float2(1.0f, v.x), sp);
              sp = __ffma2_rn(make_float2(v.y, v.y), make_float2(1.0f, v.y), sp);
            } else {
              v = __fadd2_rn(v, dr[q][r]);
            }
            if (EXACT || (n_ok && gx0 + q < W && gy0 + r < H)) {
              st_pair(reinterpret_cast<TOUT*>(op + poff[q * G::TH + r]), v);
              if (want_dz) {
                const float2 vs = round_pair(v, (TOUT*)nullptr);          // the value as stored in dx, then scaled and rounded
                st_pair(xn3 + pix0 * C + c + poff[q * G::TH + r] / (uint32_t)sizeof(TOUT), make_float2(vs.x * sc2, vs.y * sc2));
              }
            }
          }
        }
      }
      if (MODE == MODE_FWD) {
        const float2 tot = halfwarp_transpose_reduce<NPIX>(st, lane);
        const int L = lane & 15;
        const float invC = 1.0f / (float)C;
        const float mu = tot.x * invC;
        const float var = fmaxf(fmaf(-mu, mu, tot.y * invC), 0.f);
        const float rs = rsqrtf(var + eps);
        const int tb = i & 1;
        mbar_wait(lnempty_bar(tb), (uint32_t)(((i >> 1) & 1) ^ 1));   // the LN warps are done with this stats buffer
        if (L < NPIX) {
          const int q = L / G::TH, r = L - q * G::TH;
          const int ptile = ((img * G::ROWS + row0 + r) * G::TW + col0 + q);
          stats[tb * STATS_STRIDE + ptile] = make_float2(mu, rs);
          if (EXACT || (n_ok && gx0 + q < W && gy0 + r < H)) {
            const int64_t m = pix0 + (int64_t)r * W + q;
            mean_out[m] = mu;
            rstd_out[m] = rs;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(lnfull_bar(tb));        // release: y rows of this tile + statistics are visible
      }
    }
  } else if (MODE == MODE_FWD) {
    // ===== LayerNorm: xn = (y - mean) * rstd * w + b over the tile the compute warps just finished =====
    constexpr int VEC = 8, NV = LnIO<TOUT>::NV;
    const int lw = warp - 1 - NWC;
    const int VPR = C / VEC;                               // 8-channel lane items per pixel row
    if (VPR <= 32) {
      // several pixels per warp pass; this lane's slice of ln_w / ln_b stays in registers for the whole kernel
      const int ppw = 32 / VPR;
      const int psub = lane / VPR, v = lane - psub * VPR;
      const bool active = psub < ppw;
      float w[VEC], b[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) { w[e] = active ? __ldg(ln_w + v * VEC + e) : 0.f; b[e] = active ? __ldg(ln_b + v * VEC + e) : 0.f; }
      constexpr int U = CNX_DW_LNU1 / NV;                  // the same bytes in flight per lane for both dtypes
      for (int i = 0; i < my_tiles; ++i) {
        const TileCoord t = decode_tile<G>((int)blockIdx.x + i * (int)gridDim.x, tiles_x, tiles_y);
        const int tb = i & 1;
        mbar_wait(lnfull_bar(tb), (uint32_t)((i >> 1) & 1));
        const float2* stp = stats + tb * STATS_STRIDE;
        for (int p0 = lw * ppw; p0 < G::P; p0 += NLN * ppw * U) {
          uint4 raw[U][NV];
          int64_t off[U], mm[U];
          float2 ms[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + u * NLN * ppw + psub;
            off[u] = -1;
            if (active && p < G::P) {
              const int64_t m = tile_pixel<G, EXACT>(t, p, N, H, W);
              if (m >= 0) {
                off[u] = m * C + v * VEC;
                mm[u] = m;
                ln_load8(out + off[u], raw[u]);
                ms[u] = stp[p];
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (off[u] >= 0) {
              float x[VEC];
              ln_unpack8(raw[u], x, (TOUT*)nullptr);
#pragma unroll
              for (int e = 0; e < VEC; ++e) x[e] = fmaf((x[e] - ms[u].x) * ms[u].y, w[e], b[e]);
              if (sizeof(TOUT) == 4 && xn3 != nullptr) store_split8(xn3 + mm[u] * xseg * C, C, v * VEC, x, xseg == 3);
              else ln_store8(xn + off[u], x);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(lnempty_bar(tb));
      }
    } else {
      // wide rows: one pixel per warp pass, lanes stride over the row; two pixels in flight
      constexpr int U = (CNX_DW_LNU2 / NV) > 0 ? (CNX_DW_LNU2 / NV) : 1;
      for (int i = 0; i < my_tiles; ++i) {
        const TileCoord t = decode_tile<G>((int)blockIdx.x + i * (int)gridDim.x, tiles_x, tiles_y);
        const int tb = i & 1;
        mbar_wait(lnfull_bar(tb), (uint32_t)((i >> 1) & 1));
        const float2* stp = stats + tb * STATS_STRIDE;
        for (int p0 = lw; p0 < G::P; p0 += NLN * U) {
          int64_t mrow[U];
          float2 ms[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + u * NLN;
            mrow[u] = (p < G::P) ? tile_pixel<G, EXACT>(t, p, N, H, W) : -1;
            ms[u] = (p < G::P) ? stp[p] : make_float2(0.f, 0.f);
          }
          for (int v0 = 0; v0 < VPR; v0 += 32) {
            const int v = v0 + lane;
            uint4 raw[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (v < VPR && mrow[u] >= 0) ln_load8(out + mrow[u] * C + v * VEC, raw[u]);
            if (v < VPR) {
              float w[VEC], b[VEC];
#pragma unroll
              for (int e = 0; e < VEC; e += 4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(ln_w + v * VEC + e));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(ln_b + v * VEC + e));
                w[e] = w4.x; w[e + 1] = w4.y; w[e + 2] = w4.z; w[e + 3] = w4.w;
                b[e] = b4.x; b[e + 1] = b4.y; b[e + 2] = b4.z; b[e + 3] = b4.w;
              }
#pragma unroll
              for (int u = 0; u < U; ++u) {
                if (mrow[u] >= 0) {
                  float x[VEC];
                  ln_unpack8(raw[u], x, (TOUT*)nullptr);
#pragma unroll
                  for (int e = 0; e < VEC; ++e) x[e] = fmaf((x[e] - ms[u].x) * ms[u].y, w[e], b[e]);
                  if (sizeof(TOUT) == 4 && xn3 != nullptr) store_split8(xn3 + mrow[u] * xseg * C, C, v * VEC, x, xseg == 3);
                  else ln_store8(xn + mrow[u] * C + v * VEC, x);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(lnempty_bar(tb));
      }
    }
  }
}

template <class G, int MODE, typename TIN, typename TOUT, bool EXACT>
static int launch_conv(const void* x, int x_dtype, const float* wt, const float* bias, const void* dres, void* out,
                       const float* ln_w, const float* ln_b, float eps, void* xn, float* mean, float* rstd, int64_t N,
                       int64_t H, int64_t W, int64_t C, cudaStream_t s, void* xn3 = nullptr, int xseg = 3) {
  typedef ConvCfg<G, MODE, TIN, TOUT> Cfg;
  CUtensorMap tmX, tmW;
  if (int rc = make_map_nhwc(&tmX, x, x_dtype, N, H, W, C, G::HW, G::HH, G::NB)) return rc;
  if (int rc = make_map_wt(&tmW, wt, C)) return rc;
  static_assert(Cfg::SMEM <= SMEM_MAX, "dwconv tile does not fit in shared memory");
  auto k = dwconv7_v2_kernel<G, MODE, TIN, TOUT, EXACT>;
  if (int rc = set_smem(k, Cfg::SMEM)) return rc;
  int tiles_x, tiles_y;
  const int64_t nt = num_tiles<G>(N, H, W, &tiles_x, &tiles_y);
  CNX_REQUIRE(nt < (1ll << 30), CNX_E_SHAPE, "dwconv: too many tiles");
  int grid = sm_count();
  if (grid > nt) grid = (int)nt;
  launch_pdl(k, dim3(grid), dim3(Cfg::NT), Cfg::SMEM, s, tmX, tmW, (int)N, (int)H, (int)W, (int)C, tiles_x, tiles_y, (int)nt, bias,
             (const TOUT*)dres, (TOUT*)out, ln_w, ln_b, eps, (TOUT*)xn, mean, rstd, (bf16*)xn3, xseg);
  return check_launch(MODE == MODE_FWD ? "dwconv7_ln_fwd" : "dwconv7_dgrad");
}

template <int MODE, typename TIN, typename TOUT>
static int pick_conv(const void* x, int x_dtype, const float* wt, const float* bias, const void* dres, void* out,
                     const float* ln_w, const float* ln_b, float eps, void* xn, float* mean, float* rstd, int64_t N,
                     int64_t H, int64_t W, int64_t C, cudaStream_t s, void* xn3 = nullptr, int xseg = 3) {
  // forward with fp32 activations (the x3 / fp32 paths): the LayerNorm half sets the kernel time and needs its 8 warps, which
  // only fit beside 7 compute warps (measured: 28 x 8 with 4 LayerNorm warps 0.455 ms against 0.337 ms at C96, 56^2)
  const int gid = pick_geo(N, H, W, !(MODE == MODE_FWD && sizeof(TOUT) == 4));
  CNX_GEO_SWITCH(gid, {
    const bool exact = (W % G::TW == 0) && (H % G::ROWS == 0) && (N % G::NB == 0);
    if (exact)
      return launch_conv<G, MODE, TIN, TOUT, true>(x, x_dtype, wt, bias, dres, out, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s, xn3, xseg);
    return launch_conv<G, MODE, TIN, TOUT, false>(x, x_dtype, wt, bias, dres, out, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s, xn3, xseg);
  });
  return CNX_E_BADARG;
}


// ------------------------------------------------------------------------------------------------
// backward-weights (+bias): persistent CTAs, each bound to one 32-channel chunk and one "row" of the partial buffer.
// 49 tap-pair accumulators + 1 bias pair per thread live in registers across ALL the tiles the CTA visits; the x halo
// tile and the dy tile arrive by TMA through the same producer-warp ring (zero fill outside the image, so there is no
// bounds code at all); at the end the workers are summed through shared memory in a fixed order (deterministic) and one
// partial row [50][32] is written.   grid = rows * nchunks: chunk = b % nchunks, row = b / nchunks.
// ------------------------------------------------------------------------------------------------
template <class G, typename TDY, typename TX>
struct WgCfg {
  static constexpr int NWC = G::NWC;
  static constexpr int NT = (1 + NWC) * 32;
  static constexpr int HALO_BYTES = G::HALO_ELEMS * (int)sizeof(TX);
  static constexpr int HALO_PAD = round128(HALO_BYTES);
  static constexpr int DY_BYTES = G::TILE_ELEMS * (int)sizeof(TDY);
  static constexpr int STAGE_BYTES = HALO_PAD + round128(DY_BYTES);
  static constexpr int RED_BYTES = G::NWORK * 50 * CH * 4;
  static constexpr int MISC = 256 + 128;
  static constexpr int S_MAX = (SMEM_MAX - MISC) / STAGE_BYTES;
  static constexpr int STAGES = S_MAX > 4 ? 4 : S_MAX;
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int SMEM = (RING > RED_BYTES ? RING : RED_BYTES) + MISC;
  static_assert(STAGES >= 2, "dwconv wgrad tile does not leave room for a double-buffered ring");
};

template <class G, typename TDY, typename TX>
__global__ void __launch_bounds__(WgCfg<G, TDY, TX>::NT, 1)
dwconv7_wgrad_v2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, int C, int tiles_x,
                        int tiles_y, int ntiles, int rows, float* __restrict__ partial) {
  typedef WgCfg<G, TDY, TX> Cfg;
  constexpr int STAGES = Cfg::STAGES, NWC = Cfg::NWC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* base_p = smem_raw + (base_u - smem_u32(smem_raw));
  const uint32_t bars = base_u + (Cfg::SMEM - Cfg::MISC);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = C / CH;
  const int chunk = (int)blockIdx.x % nchunks, row = (int)blockIdx.x / nchunks;
  const int my_tiles = (row < ntiles) ? (ntiles - row + rows - 1) / rows : 0;

  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), NWC); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int g = 0; g < my_tiles; ++g) {
        const TileCoord t = decode_tile<G>(row + g * rows, tiles_x, tiles_y);
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t dst = base_u + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(full_bar(s), Cfg::HALO_BYTES + Cfg::DY_BYTES);
        tma_load_4d(dst, &tmX, full_bar(s), chunk * CH, t.x0 - 3, t.y0 - 3, t.n0);
        tma_load_4d(dst + Cfg::HALO_PAD, &tmDY, full_bar(s), chunk * CH, t.x0, t.y0, t.n0);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    return;                                   // the compute warps synchronise among themselves with a named barrier
  }
  const int ct = tid - 32;
  const int worker = ct >> 4, cp = ct & 15;
  const int wx = worker % G::WX, tt = worker / G::WX;
  const int wy = tt % G::WY, img = tt / G::WY;
  const int hbase = ((img * G::HH + wy * G::TH) * G::HW + wx * G::CPW) * CH + 2 * cp;
  const int dbase = ((img * G::ROWS + wy * G::TH) * G::TW + wx * G::CPW) * CH + 2 * cp;

  float2 accw[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) accw[t] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);
  int s = 0; uint32_t ph = 0;
  for (int g = 0; g < my_tiles; ++g) {
    mbar_wait(full_bar(s), ph);
    const TX* halo = reinterpret_cast<const TX*>(base_p + s * Cfg::STAGE_BYTES);
    const TDY* dsm = reinterpret_cast<const TDY*>(base_p + s * Cfg::STAGE_BYTES + Cfg::HALO_PAD);
    float2 d[G::CPW][G::TH];
#pragma unroll
    for (int q = 0; q < G::CPW; ++q)
#pragma unroll
      for (int r = 0; r < G::TH; ++r) {
        d[q][r] = ld_pair(dsm, dbase + (r * G::TW + q) * CH);
        accb = __fadd2_rn(accb, d[q][r]);
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
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));
    if (++s == STAGES) { s = 0; ph ^= 1; }
  }
  // cross-worker reduction through shared memory (fixed order), then one partial row per CTA.  Every TMA load that was
  // issued has been consumed by every compute warp once they all pass this barrier, so the ring can be reused.
  named_bar(1, NWC * 32);
  float* red = reinterpret_cast<float*>(base_p);          // [NWORK][50][32]
#pragma unroll
  for (int t = 0; t < 49; ++t) *reinterpret_cast<float2*>(red + (worker * 50 + t) * CH + 2 * cp) = accw[t];
  *reinterpret_cast<float2*>(red + (worker * 50 + 49) * CH + 2 * cp) = accb;
  named_bar(1, NWC * 32);
  for (int i = ct; i < 50 * CH; i += NWC * 32) {
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < G::NWORK; ++wv) sum += red[wv * 50 * CH + i];
    const int t = i / CH, l = i - t * CH;
    partial[((int64_t)row * 50 + t) * C + chunk * CH + l] = sum;
  }
}

template <class G, typename TDY, typename TX>
static int launch_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W, int64_t C,
                        float* partial, int P, cudaStream_t s) {
  typedef WgCfg<G, TDY, TX> Cfg;
  CUtensorMap tmX, tmDY;
  if (int rc = make_map_nhwc(&tmX, x, x_dtype, N, H, W, C, G::HW, G::HH, G::NB)) return rc;
  if (int rc = make_map_nhwc(&tmDY, dy, dy_dtype, N, H, W, C, G::TW, G::ROWS, G::NB)) return rc;
  static_assert(Cfg::SMEM <= SMEM_MAX, "dwconv wgrad tile does not fit in shared memory");
  auto k = dwconv7_wgrad_v2_kernel<G, TDY, TX>;
  if (int rc = set_smem(k, Cfg::SMEM)) return rc;
  int tiles_x, tiles_y;
  const int64_t nt = num_tiles<G>(N, H, W, &tiles_x, &tiles_y);
  CNX_REQUIRE(nt < (1ll << 30), CNX_E_SHAPE, "dwconv wgrad: too many tiles");
  const int nchunks = (int)(C / CH);
  launch_pdl(k, dim3((unsigned)(P * nchunks)), dim3(Cfg::NT), Cfg::SMEM, s, tmX, tmDY, (int)C, tiles_x, tiles_y, (int)nt, P, partial);
  return check_launch("dwconv7_wgrad");
}

}  // namespace dw2

int dwconv7_wgrad_v2(const void* dy, int dy_dtype, const void* x, int x_dtype, int64_t N, int64_t H, int64_t W, int64_t C,
                     float* partial, int P, cudaStream_t s) {
  using namespace dw2;
  const int gid = pick_geo(N, H, W);
#define CNX_WG2(TD, TX) CNX_GEO_SWITCH(gid, return (launch_wgrad<G, TD, TX>(dy, dy_dtype, x, x_dtype, N, H, W, C, partial, P, s)))
  if (dy_dtype == CNX_F32 && x_dtype == CNX_F32) CNX_WG2(float, float);
  if (dy_dtype == CNX_BF16 && x_dtype == CNX_F32) CNX_WG2(bf16, float);
  if (dy_dtype == CNX_BF16 && x_dtype == CNX_BF16) CNX_WG2(bf16, bf16);
#undef CNX_WG2
  set_error("dwconv7_wgrad: unsupported dtype combination");
  return CNX_E_BADARG;
}


// entry points used by the extern "C" layer in dwconv.cu
int dwconv7_ln_fwd_v2(const void* x, int x_dtype, const float* wt, const float* bias, const float* ln_w, const float* ln_b,
                      float eps, int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn, int act_dtype, float* mean,
                      float* rstd, cudaStream_t s) {
  using namespace dw2;
  if (x_dtype == CNX_F32 && act_dtype == CNX_F32)
    return pick_conv<MODE_FWD, float, float>(x, x_dtype, wt, bias, nullptr, y, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  if (x_dtype == CNX_F32 && act_dtype == CNX_BF16)
    return pick_conv<MODE_FWD, float, bf16>(x, x_dtype, wt, bias, nullptr, y, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  if (x_dtype == CNX_BF16 && act_dtype == CNX_BF16)
    return pick_conv<MODE_FWD, bf16, bf16>(x, x_dtype, wt, bias, nullptr, y, ln_w, ln_b, eps, xn, mean, rstd, N, H, W, C, s);
  set_error("dwconv7_ln_fwd: bf16 stream with fp32 activations is not a supported combination");
  return CNX_E_BADARG;
}

// fp32 stream, fp32 conv + LayerNorm; the normalised rows leave as the split operand xn3 bf16 [M, 3C]; y is fp32 scratch
int dwconv7_ln_fwd_x3_v2(const void* x, const float* wt, const float* bias, const float* ln_w, const float* ln_b, float eps,
                         int64_t N, int64_t H, int64_t W, int64_t C, void* y, void* xn3, int segments, float* mean, float* rstd,
                         cudaStream_t s) {
  using namespace dw2;
  return pick_conv<MODE_FWD, float, float>(x, CNX_F32, wt, bias, nullptr, y, ln_w, ln_b, eps, y, mean, rstd, N, H, W, C, s, xn3,
                                           segments);
}

int dwconv7_dgrad_v2(const void* dy, int dy_dtype, const float* wt, const void* dres, void* dx, int stream_dtype, int64_t N,
                     int64_t H, int64_t W, int64_t C, void* dz_up, const float* dp_up, cudaStream_t s) {
  using namespace dw2;
  if (dy_dtype == CNX_F32 && stream_dtype == CNX_F32) {
    CNX_REQUIRE(dz_up == nullptr, CNX_E_BADARG, "dwconv7_dgrad: the folded operand copy is a bf16-activation feature");
    return pick_conv<MODE_DGRAD, float, float>(dy, dy_dtype, wt, nullptr, dres, dx, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, N, H, W, C, s);
  }
  // (dz_up, dp_up) ride in the forward-only (xn3, ln_w) parameters of the shared kernel
  if (dy_dtype == CNX_BF16 && stream_dtype == CNX_F32)
    return pick_conv<MODE_DGRAD, bf16, float>(dy, dy_dtype, wt, nullptr, dres, dx, dp_up, nullptr, 0.f, nullptr, nullptr, nullptr, N, H, W, C, s, dz_up);
  if (dy_dtype == CNX_BF16 && stream_dtype == CNX_BF16)
    return pick_conv<MODE_DGRAD, bf16, bf16>(dy, dy_dtype, wt, nullptr, dres, dx, dp_up, nullptr, 0.f, nullptr, nullptr, nullptr, N, H, W, C, s, dz_up);
  set_error("dwconv7_dgrad: fp32 activations with a bf16 stream is not a supported combination");
  return CNX_E_BADARG;
}

}  // namespace cnx

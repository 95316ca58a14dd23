// dwconv_common.cuh — shared pieces of the depthwise 7x7 kernels (dwconv_fwd.cu, dwconv_dgrad.cu, dwconv_wgrad.cu).
//
// Common design (measured rationale in DESIGN.md §dwconv, profiles/r01b_*):
//   * warp-specialised CTAs: NWC compute warps + ONE producer warp.  The producer's elected lane issues every TMA
//     load (halo tiles: cp.async.bulk.tensor.4d over a {C, W, H, N} map, out-of-bounds elements zero-filled by the
//     TMA unit == the conv's padding=3) and, where the output leaves through shared memory, every TMA store.
//     Compute warps only ever touch shared memory and mbarriers: no per-thread global address arithmetic, no
//     bounds branches, no CTA-wide __syncthreads in the steady state (the r01b profile showed more address/bounds
//     instructions than FMAs and every warp stalling on the slowest one at each chunk).
//   * a half-warp ("worker") owns 32 channels as 16 channel PAIRS and a CPW x TH strip of output pixels: every
//     shared-memory access is a conflict-free 8-byte (fp32) / 4-byte (bf16) row and every FMA is the packed
//     fma.rn.f32x2; an input value read once from shared memory feeds up to 7*CPW packed FMAs.  The isolated core
//     sustains ~100 FMA/clk/SM with 8 warps (profiles/micro/dwconv_core.cu).
//   * tile geometries are chosen per feature-map size so that 56/28/14/7-pixel maps tile with ZERO padded work
//     (14 workers: 8x28, 7x28, 14x14, 2 x 7x7) instead of the 12-31 % that power-of-two tiles waste.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace cnx {
namespace dw2 {

constexpr int CH = 32;                        // channels per chunk (16 pairs per half-warp)
constexpr int W_BYTES = 49 * CH * 4;          // one chunk of tap-major weights
constexpr int W_STRIDE = 6400;                // padded to a multiple of 128 B
constexpr int SMEM_MAX = 227 * 1024;

__host__ __device__ constexpr int round128(int x) { return (x + 127) / 128 * 128; }

// TH x CPW output pixels per worker; WX x WY workers per image tile; NB images per tile
template <int TH_, int CPW_, int WX_, int WY_, int NB_>
struct Geo {
  static constexpr int TH = TH_, CPW = CPW_, WX = WX_, WY = WY_, NB = NB_;
  static constexpr int TW = WX * CPW;                 // output columns per tile
  static constexpr int ROWS = WY * TH;                // output rows per tile (per image)
  static constexpr int HH = ROWS + 6, HW = TW + 6;
  static constexpr int NWORK = NB * WY * WX;          // half-warp workers
  static constexpr int NWC = NWORK / 2;               // compute warps
  static constexpr int NCT = NWORK * 16;              // compute threads
  static constexpr int NT = NCT + 32;                 // + producer warp
  static constexpr int P = NB * ROWS * TW;            // output pixel slots per tile
  static constexpr int HALO_ELEMS = NB * HH * HW * CH;
  static constexpr int TILE_ELEMS = P * CH;
  static_assert(NWORK % 2 == 0, "workers come in pairs (one warp = two half-warps)");
};

typedef Geo<7, 2, 4, 4, 1> GeoT28;    // 28 x 8          H % 28 == 0, W % 8 == 0 (56, 112, 224): 16 workers = 8 compute warps, two
                                      //                 per scheduler — the 14-worker geometries leave one scheduler with a single
                                      //                 compute warp (FMA pipe capped at 7/8)
typedef Geo<8, 2, 16, 1, 1> GeoW32;   //  8 x 32         generic wide maps
typedef Geo<8, 2, 14, 1, 1> GeoW28;   //  8 x 28         W % 28 == 0, H % 8 == 0  (56, 112, 224)
typedef Geo<7, 2, 14, 1, 1> GeoS28;   //  7 x 28         28 x 28
typedef Geo<8, 2, 8, 2, 1> GeoW16;    // 16 x 16         generic medium maps
typedef Geo<7, 2, 7, 2, 1> GeoS14;    // 14 x 14         14 x 14
typedef Geo<8, 1, 8, 1, 2> GeoW8;     //  8 x 8 x 2 img  generic small maps
typedef Geo<7, 1, 7, 1, 2> GeoS7;     //  7 x 7 x 2 img  7 x 7

enum GeoId { GEO_T28 = 0, GEO_W32, GEO_W28, GEO_S28, GEO_W16, GEO_S14, GEO_W8, GEO_S7, GEO_COUNT };

// least padded work; ties go to the earlier (larger-tile) entry
inline int pick_geo(int64_t N, int64_t H, int64_t W, bool allow_t28 = true) {
  static const int tw[GEO_COUNT] = {8, 32, 28, 28, 16, 14, 8, 7};
  static const int rows[GEO_COUNT] = {28, 8, 8, 7, 16, 14, 8, 7};
  static const int nb[GEO_COUNT] = {1, 1, 1, 1, 1, 1, 2, 2};
  static int t28 = -1;                          // CNX_DW_T28=0: without the 16-worker 28 x 8 geometry (A/B measurements)
  if (t28 < 0) { const char* e = getenv("CNX_DW_T28"); t28 = (e && e[0] == '0') ? 0 : 1; }
  int best = 0;
  double bw = 1e30;
  for (int g = (t28 && allow_t28) ? 0 : 1; g < GEO_COUNT; ++g) {
    double work = (double)((W + tw[g] - 1) / tw[g] * tw[g]) * (double)((H + rows[g] - 1) / rows[g] * rows[g]) *
                  (double)((N + nb[g] - 1) / nb[g] * nb[g]);
    if (work < bw * 0.999) { bw = work; best = g; }
  }
  return best;
}

#define CNX_GEO_SWITCH(gid, ...)                                                 \
  switch (gid) {                                                                 \
    case cnx::dw2::GEO_T28: { typedef cnx::dw2::GeoT28 G; __VA_ARGS__; } break;    \
    case cnx::dw2::GEO_W32: { typedef cnx::dw2::GeoW32 G; __VA_ARGS__; } break;    \
    case cnx::dw2::GEO_W28: { typedef cnx::dw2::GeoW28 G; __VA_ARGS__; } break;    \
    case cnx::dw2::GEO_S28: { typedef cnx::dw2::GeoS28 G; __VA_ARGS__; } break;    \
    case cnx::dw2::GEO_W16: { typedef cnx::dw2::GeoW16 G; __VA_ARGS__; } break;    \
    case cnx::dw2::GEO_S14: { typedef cnx::dw2::GeoS14 G; __VA_ARGS__; } break;    \
    case cnx::dw2::GEO_W8: { typedef cnx::dw2::GeoW8 G; __VA_ARGS__; } break;      \
    default: { typedef cnx::dw2::GeoS7 G; __VA_ARGS__; } break;                   \
  }

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
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
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// channel-pair loads/stores on a [pixel][32 ch] shared-memory tile
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

// per-thread constants of the worker decomposition
template <class G>
struct Worker {
  int cp;        // channel pair within the chunk (0..15)
  int hbase;     // element offset of this worker's strip origin in the halo tile, + 2*cp
  int obase;     // element offset of this worker's strip origin in a [P][32] output tile, + 2*cp
  __device__ __forceinline__ Worker(int tid) {
    const int worker = tid >> 4;
    cp = tid & 15;
    const int wx = worker % G::WX;
    const int t = worker / G::WX;
    const int wy = t % G::WY, img = t / G::WY;
    hbase = ((img * G::HH + wy * G::TH) * G::HW + wx * G::CPW) * CH + 2 * cp;
    obase = ((img * G::ROWS + wy * G::TH) * G::TW + wx * G::CPW) * CH + 2 * cp;
  }
};

// The register-tiled 7x7 correlation for one 32-channel chunk: acc[q][r] for output column q, row r of the strip.
template <class G, typename TS, bool FLIP>
__device__ __forceinline__ void conv_chunk(const TS* __restrict__ halo, const float* __restrict__ wsm, int hbase, int cp,
                                           float2 (&acc)[G::CPW][G::TH], float2 init = make_float2(0.f, 0.f)) {
  float2 wr[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wr[t] = *reinterpret_cast<const float2*>(wsm + (FLIP ? 48 - t : t) * CH + 2 * cp);
#pragma unroll
  for (int q = 0; q < G::CPW; ++q)
#pragma unroll
    for (int r = 0; r < G::TH; ++r) acc[q][r] = init;
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
            if (r >= 0 && r < G::TH) acc[q][r] = __ffma2_rn(v, wr[ky * 7 + kx], acc[q][r]);
          }
        }
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode() {
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
inline int make_map_nhwc(CUtensorMap* map, const void* ptr, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, int bw,
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
inline int make_map_wt(CUtensorMap* map, const float* wt, int64_t C) {
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
inline int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

template <class G>
inline int64_t num_tiles(int64_t N, int64_t H, int64_t W, int* tiles_x, int* tiles_y) {
  *tiles_x = (int)((W + G::TW - 1) / G::TW);
  *tiles_y = (int)((H + G::ROWS - 1) / G::ROWS);
  return (int64_t)(*tiles_x) * (*tiles_y) * ((N + G::NB - 1) / G::NB);
}

}  // namespace dw2
}  // namespace cnx

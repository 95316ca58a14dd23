// Microbenchmark: HBM write bandwidth of the GEMM epilogues' output path — every warp of a persistent CTA fills a shared-memory
// slab with st.shared.v4 and hands it to a TMA store (cp.async.bulk.tensor.2d), double-buffered, over a [M, N] bf16 or fp32
// matrix.  Variables: slab row width in bytes (64 = the [32 x 32] bf16 slabs of the fc1 / dGELU / x3 epilogues, 128 = the fp32
// slabs and a [32 x 64] bf16 slab), warps per CTA, and how many column blocks a warp writes per row block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_store_bw tma_store_bw.cu -lcuda && ./tma_store_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ROWB: slab row bytes (64 or 128); a slab is [32 rows][ROWB bytes]; each warp owns 2 slabs
template <int ROWB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) store_kernel(const __grid_constant__ CUtensorMap tm, int64_t M, int64_t row_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int SLAB = 32 * ROWB;
  const uint32_t slab0 = smem_u32(smem) + warp * 2 * SLAB;
  const int64_t rb_total = M / 32;                      // 32-row blocks
  const int64_t cb_total = row_bytes / ROWB;            // column blocks per row
  const int64_t items = rb_total * cb_total;
  int buf = 0;
  // item -> (row block, column block): consecutive warps take consecutive column blocks of one row block (as the epilogues do)
  for (int64_t it = (int64_t)blockIdx.x * WARPS + warp; it < items; it += (int64_t)gridDim.x * WARPS) {
    const int64_t rb = it / cb_total, cb = it - rb * cb_total;
    const uint32_t slab = slab0 + buf * SLAB;
    if (lane == 0) bulk_wait_read1();
    __syncwarp();
    const uint32_t v = (uint32_t)it;
#pragma unroll
    for (int j = 0; j < ROWB / 16; ++j)
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(slab + lane * ROWB + (((uint32_t)j ^ (lane & (ROWB / 16 - 1))) << 4)), "r"(v));
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tm, slab, (int32_t)(cb * (ROWB / 2)), (int32_t)(rb * 32));     // map in 2-byte elements
      bulk_commit();
    }
    buf ^= 1;
  }
  if (lane == 0) bulk_wait0();
}

// the plain way: each thread writes 16 bytes, fully coalesced rows (what torch's fill does)
__global__ void __launch_bounds__(512, 1) plain_kernel(uint4* out, int64_t n16) {
  for (int64_t i = (int64_t)blockIdx.x * 512 + threadIdx.x; i < n16; i += (int64_t)gridDim.x * 512) out[i] = make_uint4(1, 2, 3, (uint32_t)i);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWB, int WARPS>
static void run(EncodeTiledFn enc, void* buf, int64_t M, int64_t N, const char* tag) {
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t gstr[1] = {(cuuint64_t)N * 2};
  cuuint32_t box[2] = {ROWB / 2, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
  auto k = store_kernel<ROWB, WARPS>;
  const int smem = WARPS * 2 * 32 * ROWB;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int i = 0; i < 6; ++i) {
    cudaEventRecord(e0);
    k<<<148, WARPS * 32, smem>>>(tm, M, N * 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (i > 0 && ms < best) best = ms;
  }
  printf("%-34s rows of %3d B, %2d warps/CTA: %.3f ms  %.0f GB/s  (%s)\n", tag, ROWB, WARPS, best, (double)M * N * 2 / best / 1e6,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int64_t M = 802816;                              // 256 x 56 x 56 rows
  void* buf; cudaMalloc(&buf, M * 1536 * 2);
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    const int64_t n16 = M * 768 * 2 / 16;
    for (int i = 0; i < 6; ++i) {
      cudaEventRecord(e0); plain_kernel<<<148 * 4, 512>>>((uint4*)buf, n16); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (i > 0 && ms < best) best = ms;
    }
    printf("%-34s %.3f ms  %.0f GB/s\n", "plain 16-byte stores [M,768] bf16", best, (double)M * 768 * 2 / best / 1e6);
  }
  run<64, 16>(enc, buf, M, 768, "TMA slabs [M,768] bf16");
  run<128, 16>(enc, buf, M, 768, "TMA slabs [M,768] bf16");
  run<64, 8>(enc, buf, M, 768, "TMA slabs [M,768] bf16");
  run<128, 8>(enc, buf, M, 768, "TMA slabs [M,768] bf16");
  run<64, 16>(enc, buf, M, 384, "TMA slabs [M,384] bf16");
  run<128, 16>(enc, buf, M, 384, "TMA slabs [M,384] bf16");
  run<128, 16>(enc, buf, M, 192, "TMA slabs [M,96] fp32 (as bf16 x2)");
  run<64, 16>(enc, buf, M, 1536, "TMA slabs [M,1536] bf16");
  run<128, 16>(enc, buf, M, 1536, "TMA slabs [M,1536] bf16");
  return 0;
}

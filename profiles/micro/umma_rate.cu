// Microbenchmark: what one tcgen05.mma (cta_group::1, kind::f16, bf16 operands from shared memory, M = 128, K = 16) costs as a
// function of N and of the dependency between consecutive instructions — the same TMEM accumulator every time (what a GEMM's K
// loop does), or 2 / 4 accumulators in rotation.  One CTA per SM on every SM (the operands' shared-memory reads and TMEM traffic
// are per SM), one thread issues ITERS instructions and one tcgen05.commit; cycles from the first issue to the commit's arrival.
// Also: the issuing thread's own time for the same sequence (clock after the last issue).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B, K-major
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// NACC accumulators of N columns each in rotation; KSTEPS k-slices (of 16) of a [rows][64] operand box are walked in rotation
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 128 * 128, bar = sB + 256 * 128, slot = bar + 16;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(128, N);
    const uint64_t adesc = make_smem_desc(sA, 16, 1024), bdesc = make_smem_desc(sB, 16, 1024);
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      const int k = i & 3, a = i % NACC;
      umma_f16(tmem + a * N, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, i >= NACC);
    }
    umma_commit(bar);
    const long long t1 = clock64();
    while (!mbar_try_wait(bar, 0)) {}
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int NACC>
static void run(long long* dout) {
  static_assert(N * NACC <= 512, "TMEM columns");
  auto k = rate_kernel<N, NACC>;
  const int smem = (128 + 256) * 128 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 3; ++rep) {
    k<<<148, 128, smem>>>(iters, dout);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d  accumulators in rotation=%d : %6.1f cycles per MMA to completion (pipe floor %3d), issuing thread alone %6.1f   %s\n", N,
         NACC, (double)h[1] / iters, 128 * N / 256, (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 16);
  run<32, 1>(dout);  run<32, 2>(dout);  run<32, 4>(dout);
  run<64, 1>(dout);  run<64, 2>(dout);  run<64, 4>(dout);
  run<96, 1>(dout);  run<96, 2>(dout);  run<96, 4>(dout);
  run<128, 1>(dout); run<128, 2>(dout); run<128, 4>(dout);
  run<256, 1>(dout); run<256, 2>(dout);
  return 0;
}

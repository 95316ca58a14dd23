// Microbenchmark of the depthwise-7x7 register-tiled FMA core (no global traffic): how many FMA/clk/SM the
// inner loop of csrc/dwconv.cu can sustain as a function of strip shape, operand type, packing and occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dwconv_core dwconv_core.cu && ./dwconv_core
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
typedef __nv_bfloat16 bf16;
constexpr int CH = 32;

__device__ __forceinline__ float2 ld_pair(const float* sm, int idx) { return *reinterpret_cast<const float2*>(sm + idx); }
__device__ __forceinline__ float2 ld_pair(const bf16* sm, int idx) {
  uint32_t u = *reinterpret_cast<const uint32_t*>(sm + idx);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ float ld_one(const float* sm, int idx) { return sm[idx]; }
__device__ __forceinline__ float ld_one(const bf16* sm, int idx) { return __bfloat162float(sm[idx]); }

// PACK=2: lane = channel pair (16 lanes per worker), FFMA2.  PACK=1: lane = channel (32 lanes per worker), FFMA.
template <int TH, int CPW, typename TS, int NT, int MINB, int PACK, bool WREG>
__global__ void __launch_bounds__(NT, MINB) core(float* out, int iters) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int LPW = 32 / PACK;                 // lanes per worker
  constexpr int NWK = NT / LPW;                  // workers per CTA
  constexpr int HW = 6 + CPW * 8, HH = TH + 6;   // 8 workers side by side in x share one halo; others alias it
  TS* halo = reinterpret_cast<TS*>(smem);
  float* wsm = reinterpret_cast<float*>(smem + ((HH * HW * CH * sizeof(TS) + 127) / 128) * 128);
  for (int i = threadIdx.x; i < HH * HW * CH; i += NT) halo[i] = (TS)(float)((i * 7) % 13 - 6);
  for (int i = threadIdx.x; i < 49 * CH; i += NT) wsm[i] = (float)((i * 5) % 11 - 5) * 0.01f;
  __syncthreads();
  const int worker = threadIdx.x / LPW, cp = threadIdx.x % LPW;
  const int col0 = (worker % 8) * CPW;
  const int base = col0 * CH + PACK * cp;
  if constexpr (PACK == 2) {
    float2 acc[CPW][TH];
#pragma unroll
    for (int q = 0; q < CPW; ++q)
#pragma unroll
      for (int r = 0; r < TH; ++r) acc[q][r] = make_float2(0.f, 0.f);
    float2 wr[49];
    if (WREG) {
#pragma unroll
      for (int t = 0; t < 49; ++t) wr[t] = *reinterpret_cast<const float2*>(wsm + t * CH + 2 * cp);
    }
    for (int it = 0; it < iters; ++it) {
      if (!WREG) {
#pragma unroll
        for (int t = 0; t < 49; ++t) wr[t] = *reinterpret_cast<const float2*>(wsm + t * CH + 2 * cp);
      }
#pragma unroll
      for (int j = 0; j < 6 + CPW; ++j) {
#pragma unroll
        for (int iy = 0; iy < TH + 6; ++iy) {
          const float2 v = ld_pair(halo, base + (iy * HW + j) * CH);
#pragma unroll
          for (int q = 0; q < CPW; ++q) {
            const int kx = j - q;
            if (kx >= 0 && kx <= 6) {
#pragma unroll
              for (int ky = 0; ky < 7; ++ky) {
                const int r = iy - ky;
                if (r >= 0 && r < TH) acc[q][r] = __ffma2_rn(v, wr[ky * 7 + kx], acc[q][r]);
              }
            }
          }
        }
      }
      asm volatile("" ::: "memory");
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < CPW; ++q)
#pragma unroll
      for (int r = 0; r < TH; ++r) { s.x += acc[q][r].x; s.y += acc[q][r].y; }
    out[(size_t)blockIdx.x * NT + threadIdx.x] = s.x + s.y;
  } else {
    float acc[CPW][TH];
#pragma unroll
    for (int q = 0; q < CPW; ++q)
#pragma unroll
      for (int r = 0; r < TH; ++r) acc[q][r] = 0.f;
    float wr[49];
    if (WREG) {
#pragma unroll
      for (int t = 0; t < 49; ++t) wr[t] = wsm[t * CH + cp];
    }
    for (int it = 0; it < iters; ++it) {
      if (!WREG) {
#pragma unroll
        for (int t = 0; t < 49; ++t) wr[t] = wsm[t * CH + cp];
      }
#pragma unroll
      for (int j = 0; j < 6 + CPW; ++j) {
#pragma unroll
        for (int iy = 0; iy < TH + 6; ++iy) {
          const float v = ld_one(halo, base + (iy * HW + j) * CH);
#pragma unroll
          for (int q = 0; q < CPW; ++q) {
            const int kx = j - q;
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
      asm volatile("" ::: "memory");
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < CPW; ++q)
#pragma unroll
      for (int r = 0; r < TH; ++r) s += acc[q][r];
    out[(size_t)blockIdx.x * NT + threadIdx.x] = s;
  }
}

template <int TH, int CPW, typename TS, int NT, int MINB, int PACK, bool WREG>
void run(const char* ts, float* out, int nsm) {
  auto k = core<TH, CPW, TS, NT, MINB, PACK, WREG>;
  constexpr int HW = 6 + CPW * 8, HH = TH + 6;
  size_t smem = ((HH * HW * CH * sizeof(TS) + 127) / 128) * 128 + 49 * CH * 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k);
  const int iters = 200;
  const int grid = nsm * occ;
  k<<<grid, NT, smem>>>(out, 10);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<<<grid, NT, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double fma = (double)grid * NT * iters * 49.0 * CPW * TH * PACK;
  double tf = 2.0 * fma / (best * 1e-3) / 1e12;
  printf("TH=%d CPW=%d %-5s NT=%d minb=%d occ=%d PACK=%d WREG=%d regs=%3d smem=%6zu  %.3f ms  %.1f TFLOP/s  (%.0f FMA/clk/SM @1.965GHz)  err=%s\n",
         TH, CPW, ts, NT, MINB, occ, PACK, (int)WREG, fa.numRegs, smem, best, tf, fma / (best * 1e-3) / 1.965e9 / nsm,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, (size_t)nsm * 8 * 1024 * 4);
#define R(TH, CPW, TS, NT, MINB, PACK, WREG) run<TH, CPW, TS, NT, MINB, PACK, WREG>(#TS, out, nsm)
  R(8, 2, float, 256, 1, 2, false); R(8, 2, float, 256, 2, 2, false); R(8, 2, float, 256, 1, 2, true); R(8, 2, float, 384, 1, 2, true);
  R(8, 2, bf16, 256, 1, 2, false); R(8, 2, bf16, 256, 2, 2, false); R(8, 2, bf16, 256, 1, 2, true); R(8, 2, bf16, 384, 1, 2, true);
  R(4, 2, float, 256, 2, 2, false); R(4, 2, float, 256, 3, 2, false); R(4, 2, float, 256, 4, 2, false); R(4, 2, float, 256, 2, 2, true); R(4, 2, float, 256, 3, 2, true);
  R(4, 2, bf16, 256, 2, 2, false); R(4, 2, bf16, 256, 3, 2, false); R(4, 2, bf16, 256, 4, 2, false); R(4, 2, bf16, 256, 2, 2, true);
  R(4, 4, float, 256, 2, 2, false); R(4, 4, bf16, 256, 2, 2, false); R(4, 4, bf16, 256, 2, 2, true);
  R(8, 1, float, 256, 2, 2, false); R(8, 1, float, 256, 3, 2, false); R(8, 1, bf16, 256, 3, 2, false);
  R(8, 2, float, 256, 2, 1, false); R(8, 2, float, 256, 3, 1, false); R(8, 2, float, 256, 4, 1, false); R(8, 2, float, 256, 4, 1, true);
  R(8, 2, bf16, 256, 3, 1, false); R(8, 2, bf16, 256, 4, 1, false); R(8, 2, bf16, 256, 4, 1, true);
  R(8, 4, float, 256, 2, 1, false); R(8, 4, float, 256, 3, 1, false); R(8, 4, bf16, 256, 3, 1, false); R(8, 4, bf16, 256, 3, 1, true);
  R(4, 4, float, 256, 4, 1, false); R(4, 4, float, 256, 4, 1, true); R(4, 4, bf16, 256, 4, 1, false);
  return 0;
}

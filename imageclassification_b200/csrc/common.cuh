// common.cuh — shared device/host helpers for libcnx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/cnx.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcnx is written for sm_100a (B200) only"
#endif

typedef __nv_bfloat16 bf16;

namespace cnx {

// ---- error reporting (thread-local message, errno-style return codes) -------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError() -> return code (+message)
int sm_count();

#define CNX_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      cnx::set_error(__VA_ARGS__);              \
      return (code);                            \
    }                                           \
  } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// Every hot-path kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may start as soon as
// the CTAs of the previous kernel in the stream have exited (SMs that finished their share of a persistent kernel pick up
// the next kernel's CTAs, which run their prologue — shared-memory carve-up, mbarrier init, TMEM allocation, tensor-map
// prefetch — under the previous kernel's tail).  A kernel launched this way MUST execute pdl_wait() in every thread before
// its first global-memory access: it blocks until the previous grid has completed and its writes are visible (and the
// kernel's own writes cannot overtake the previous grid's reads).  Without the launch attribute pdl_wait() is a no-op.
// CNX_PDL=0 in the environment launches everything the ordinary way (for A/B measurements).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KP, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KP>(args)...);
}
#endif

inline bool dtype_ok(int d) { return d == CNX_F32 || d == CNX_BF16; }
inline int dtype_size(int d) { return d == CNX_BF16 ? 2 : 4; }

// ---- typed load/store with fp32 math ----------------------------------------------------------
template <typename T> struct Vec4;   // 4 consecutive elements
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<bf16> { typedef uint2 type; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
// value after a round trip through T (what the reference sees after an autocast cast)
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f32(from_f32<T>(v)); }

__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (nn.GELU() default, convnext.py:37) and its derivative, fp32 math
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float kAlpha = 0.70710678118654752440f;        // 1/sqrt(2)
  const float kBeta = 0.39894228040143267794f;         // 1/sqrt(2*pi)
  float cdf = 0.5f * (1.0f + erff(x * kAlpha));
  float pdf = __expf(-0.5f * x * x) * kBeta;
  return cdf + x * pdf;
}

// dispatch helper: call f.template operator()<T>() for a dtype enum
#define CNX_DISPATCH_DTYPE(dt, T, ...)                \
  do {                                                \
    if ((dt) == CNX_F32) { typedef float T; __VA_ARGS__; } \
    else { typedef bf16 T; __VA_ARGS__; }             \
  } while (0)

}  // namespace cnx

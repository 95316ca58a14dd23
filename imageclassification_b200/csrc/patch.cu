// patch.cu — the data movement of the patchify convolutions (stem 4x4 stride 4, downsample 2x2 stride 2; §8f-2,
// semantic_segmentation/backbone/convnext.py:79-89).  A k x k stride-k convolution is a GEMM over non-overlapping patches:
// these kernels only re-order activations into (and out of) the [patches, k*k*C] row-major operand the tcgen05 GEMMs read;
// the arithmetic runs in gemm_tc.cu.  Pure bandwidth kernels: 16-byte vectors, grid-stride, coalesced on the wide side.
#include "common.cuh"

namespace cnx {

// x [N,Cin,H,W] fp32 (NCHW, the loader's layout) -> out [N*(H/4)*(W/4), Cin*16] (TOUT), k = (ci*4 + ky)*4 + kx: the
// flattening of the canonical [Cout,Cin,4,4] weight, so the GEMM's B operand is the parameter itself
template <typename TOUT>
__global__ void __launch_bounds__(256) patchify4_kernel(const float* __restrict__ x, int N, int Cin, int H, int W,
                                                        TOUT* __restrict__ out) {
  pdl_wait();
  const int OW = W >> 2, OH = H >> 2;
  const int64_t total = (int64_t)N * OH * OW * Cin * 4;        // one float4 (4 kx) per work item
  const int K = Cin * 16;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    // item order: ox fastest (coalesced float4 reads along an image row), then ky, ci, oy, n
    int64_t r = i;
    const int ox = (int)(r % OW); r /= OW;
    const int ky = (int)(r & 3); r >>= 2;
    const int ci = (int)(r % Cin); r /= Cin;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)n * Cin + ci) * H + (oy * 4 + ky)) * W + ox * 4));
    const float f[4] = {v.x, v.y, v.z, v.w};
    store4(out + (((int64_t)n * OH + oy) * OW + ox) * K + (ci * 4 + ky) * 4, f);
  }
}

// in [N,H,W,C] -> out [N,H/2,W/2,(2,2,C)]  (GATHER = true)   or the inverse (GATHER = false); VEC-byte elements moved as 16 B
__global__ void __launch_bounds__(256) patch2_kernel(const uint4* __restrict__ in, int N, int H, int W, int vpc /*16B vectors per pixel*/,
                                                     uint4* __restrict__ out, int gather) {
  pdl_wait();
  const int64_t total = (int64_t)N * H * W * vpc;
  const int OW = W >> 1, OH = H >> 1;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t r = i;
    const int v = (int)(r % vpc); r /= vpc;
    const int xx = (int)(r % W); r /= W;
    const int yy = (int)(r % H);
    const int n = (int)(r / H);
    const int64_t j = ((((int64_t)n * OH + (yy >> 1)) * OW + (xx >> 1)) * 4 + ((yy & 1) * 2 + (xx & 1))) * vpc + v;
    if (gather) out[j] = in[i];
    else out[i] = in[j];
  }
}

// ------------------------------------------------------------------------------------------------
// head: global average pool over the H*W pixels of a channels-last map (timm SelectAdaptivePool2d('avg'), the first op of
// NormMlpClassifierHead) and its backward.  One thread per (sample, 4 channels): the HW rows of a sample are read with
// coalesced 16-byte (fp32) / 8-byte (bf16) vectors, fp32 accumulation in pixel order (deterministic).
// ------------------------------------------------------------------------------------------------
template <typename TX>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const TX* __restrict__ x, int64_t N, int HW, int C, float* __restrict__ out) {
  pdl_wait();
  const int c4 = C >> 2;
  const int64_t total = N * c4;
  const float inv = 1.0f / (float)HW;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t n = i / c4;
    const int c = (int)(i - n * c4) * 4;
    const TX* p = x + n * HW * C + c;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < HW; ++r) {
      float v[4];
      load4(p + (int64_t)r * C, v);
      a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += v[3];
    }
    a[0] *= inv; a[1] *= inv; a[2] *= inv; a[3] *= inv;
    store4(out + n * C + c, a);
  }
}
template <typename TX>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dout, int64_t N, int HW, int C, TX* __restrict__ dx) {
  pdl_wait();
  const int c4 = C >> 2;
  const int64_t total = N * HW * c4;
  const float inv = 1.0f / (float)HW;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t m = i / c4;
    const int c = (int)(i - m * c4) * 4;
    const int64_t n = m / HW;
    float v[4];
    load4(dout + n * C + c, v);
    v[0] *= inv; v[1] *= inv; v[2] *= inv; v[3] *= inv;
    store4(dx + m * C + c, v);
  }
}

}  // namespace cnx

using namespace cnx;

extern "C" {

int cnx_patchify4_nchw(const float* x, int64_t N, int64_t Cin, int64_t H, int64_t W, void* out, int out_dtype, void* stream) {
  CNX_REQUIRE(x && out, CNX_E_BADARG, "patchify4: null pointer");
  CNX_REQUIRE(N > 0 && Cin > 0 && H > 0 && W > 0 && dtype_ok(out_dtype), CNX_E_BADARG, "patchify4: bad shape/dtype");
  CNX_REQUIRE(H % 4 == 0 && W % 4 == 0, CNX_E_SHAPE, "patchify4: H=%lld and W=%lld must be multiples of 4", (long long)H,
              (long long)W);
  const int64_t items = N * (H / 4) * (W / 4) * Cin * 4;
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == CNX_F32) launch_pdl(patchify4_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, s, x, (int)N, (int)Cin, (int)H, (int)W, (float*)out);
  else launch_pdl(patchify4_kernel<bf16>, dim3((unsigned)blocks), dim3(256), 0, s, x, (int)N, (int)Cin, (int)H, (int)W, (bf16*)out);
  return check_launch("patchify4");
}

int cnx_patch2(const void* in, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, void* out, int gather, void* stream) {
  CNX_REQUIRE(in && out, CNX_E_BADARG, "patch2: null pointer");
  CNX_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && dtype_ok(dtype), CNX_E_BADARG, "patch2: bad shape/dtype");
  CNX_REQUIRE(H % 2 == 0 && W % 2 == 0, CNX_E_SHAPE, "patch2: H=%lld and W=%lld must be even", (long long)H, (long long)W);
  const int64_t row_bytes = C * dtype_size(dtype);
  CNX_REQUIRE(row_bytes % 16 == 0, CNX_E_SHAPE, "patch2: C=%lld rows must be a multiple of 16 bytes", (long long)C);
  CNX_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, CNX_E_SHAPE, "patch2: pointers must be 16-byte aligned");
  const int vpc = (int)(row_bytes / 16);
  const int64_t items = N * H * W * vpc;
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  launch_pdl(patch2_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const uint4*)in, (int)N, (int)H, (int)W, vpc, (uint4*)out,
                                                                     gather);
  return check_launch("patch2");
}

int cnx_avgpool_nhwc_fwd(const void* x, int x_dtype, int64_t N, int64_t HW, int64_t C, float* out, void* stream) {
  CNX_REQUIRE(x && out && N > 0 && HW > 0 && C > 0 && dtype_ok(x_dtype), CNX_E_BADARG, "avgpool_fwd: bad argument");
  CNX_REQUIRE(C % 4 == 0, CNX_E_SHAPE, "avgpool_fwd: C=%lld must be a multiple of 4", (long long)C);
  int64_t blocks = (N * (C / 4) + 255) / 256;
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == CNX_F32) launch_pdl(avgpool_fwd_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, s, (const float*)x, N, (int)HW, (int)C, out);
  else launch_pdl(avgpool_fwd_kernel<bf16>, dim3((unsigned)blocks), dim3(256), 0, s, (const bf16*)x, N, (int)HW, (int)C, out);
  return check_launch("avgpool_fwd");
}

int cnx_avgpool_nhwc_bwd(const float* dout, int64_t N, int64_t HW, int64_t C, void* dx, int dx_dtype, void* stream) {
  CNX_REQUIRE(dout && dx && N > 0 && HW > 0 && C > 0 && dtype_ok(dx_dtype), CNX_E_BADARG, "avgpool_bwd: bad argument");
  CNX_REQUIRE(C % 4 == 0, CNX_E_SHAPE, "avgpool_bwd: C=%lld must be a multiple of 4", (long long)C);
  int64_t blocks = (N * HW * (C / 4) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (dx_dtype == CNX_F32) launch_pdl(avgpool_bwd_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, s, dout, N, (int)HW, (int)C, (float*)dx);
  else launch_pdl(avgpool_bwd_kernel<bf16>, dim3((unsigned)blocks), dim3(256), 0, s, dout, N, (int)HW, (int)C, (bf16*)dx);
  return check_launch("avgpool_bwd");
}

}  // extern "C"

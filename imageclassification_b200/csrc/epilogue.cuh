// epilogue.cuh — the fused GEMM epilogues of the Block MLP, shared by the tcgen05 kernel (gemm_tc.cu)
// and the fp32 CUDA-core kernel (gemm_simt.cu).  An epilogue consumes 8 consecutive accumulator columns
// of one output row and performs all the elementwise work the reference does in separate ATen kernels
// (convnext.py:48-55): bias, exact GELU, layer-scale, drop-path scale, residual add, GELU'.
#pragma once
#include "common.cuh"

namespace cnx {

enum {
  EPI_PLAIN = 0,        // out0 = acc (+bias)                                  [out dtype = TOUT]
  EPI_BIAS_GELU = 1,    // h = round(acc + b1); GELU'(h) -> out0 (optional, saved for backward), g = GELU(h) -> out1 [act dtype]
  EPI_SCALE_RES = 2,    // out0 = shortcut + dp[row/rps] * gamma[n] * (acc + b2[n])  [stream dtype]
  EPI_DGELU = 3,        // out0 = acc * gp[m,n]                                 [act dtype], aux = gp = GELU'(h) saved by BIAS_GELU
  EPI_BIAS_GELU3 = 4,   // fp32-accurate forward: g = GELU_erf(acc + b1) in fp32 -> out1 bf16 [M, 2N] = [hi(g) | mid(g)]
                        // (split operand of the next GEMM, read with a_wrap = 2N; tcgen05 slab epilogue only)
  EPI_DGELU_RC = 5,     // out0 = acc * GELU'(round(acc2 + b1)), acc2 = A2.B2^T RECOMPUTED in the same kernel (A2 = xn, B2 = W1): the
                        // forward then stores g only, not GELU'(h) (tcgen05 CTA-pair slab kernel only; HBM-bound stages)
  EPI_DGELU3 = 6        // fp32 training: dh = acc * gp[m,n] with gp = GELU'(h) in FP32 (aux) -> out0 bf16 [M, 2N] = [hi(dh) | mid(dh)]
                        // (split operand of the two GEMMs that consume dh; tcgen05 slab epilogue only)
};

struct EpiParams {
  const float* bias;        // [N] or null
  const float* gamma;       // [N] or null
  const float* dp;          // [num_samples] or null
  int64_t rows_per_sample;  // H*W
  const void* aux;          // shortcut (TOUT) for SCALE_RES, GELU'(h) (TOUT) for DGELU
  void* out0;
  void* out1;
  int64_t ld;               // row stride of out0/out1/aux (= N)
  const void* a2;           // EPI_DGELU_RC: second operand pair A2 [M,K], B2 [N,K] of the recomputed pre-activation
  const void* b2;
  int32_t a_wrap;           // tcgen05 kernels: A is stored with only this many columns and the K loop wraps around it (0 = off):
                            // the split operand [hi | mid | hi] kept as [hi | mid], its third segment re-reads the first
  int32_t round_bf16;       // EPI_PLAIN with an fp32 output: store fp32(bf16(acc + bias)) — what a bf16 GEMM output widened by the
                            // consumer holds, without the bf16 tensor and the cast pass (downsample conv -> fp32 residual stream)
  int32_t pf_tiles;         // tcgen05 kernels: the TMA producer prefetches into L2 the A rows (and the epilogue's input slab rows)
                            // of the tile this many persistent-loop steps ahead (0 = off; set by the launcher)
};

// acc[8] -> global, columns n..n+7 of row m (n % 8 == 0, N % 8 == 0 so vectors never straddle a row end)
template <int KIND, typename TOUT>
__device__ __forceinline__ void epilogue_store8(const EpiParams& p, int64_t m, int64_t n, float (&acc)[8]) {
  const int64_t off = m * p.ld + n;
  if (KIND == EPI_PLAIN) {
    if (p.bias) {
      float b[8];
      load8(p.bias + n, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += b[i];
    }
    if (sizeof(TOUT) == 4 && p.round_bf16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __bfloat162float(__float2bfloat16_rn(acc[i]));
    }
    store8(reinterpret_cast<TOUT*>(p.out0) + off, acc);
  } else if (KIND == EPI_BIAS_GELU) {
    float b[8], g[8];
    load8(p.bias + n, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // h is rounded to the activation dtype before GELU, as autocast's Linear output is
      acc[i] = round_to<TOUT>(acc[i] + b[i]);
      g[i] = gelu_erf(acc[i]);
    }
    store8(reinterpret_cast<TOUT*>(p.out1) + off, g);
    if (p.out0) {
      // what backward needs of h is only GELU'(h): saved instead of h, so the dgrad epilogue is one multiply
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = gelu_erf_grad(acc[i]);
      store8(reinterpret_cast<TOUT*>(p.out0) + off, g);
    }
  } else if (KIND == EPI_SCALE_RES) {
    float s = 1.0f;
    if (p.dp) s = p.dp[m / p.rows_per_sample];
    float b[8], gm[8], sc[8];
    if (p.bias) load8(p.bias + n, b);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i] = 0.f;
    }
    if (p.gamma) load8(p.gamma + n, gm);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) gm[i] = 1.f;
    }
    if (p.aux) load8(reinterpret_cast<const TOUT*>(p.aux) + off, sc);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) sc[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = sc[i] + s * (gm[i] * (acc[i] + b[i]));
    store8(reinterpret_cast<TOUT*>(p.out0) + off, acc);
  } else {  // EPI_DGELU
    float gp[8];
    load8(reinterpret_cast<const TOUT*>(p.aux) + off, gp);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= gp[i];
    store8(reinterpret_cast<TOUT*>(p.out0) + off, acc);
  }
}

}  // namespace cnx

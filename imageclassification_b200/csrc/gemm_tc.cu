// gemm_tc.cu — a3/a5 and their four backward GEMMs on the 5th-gen tensor cores (sm_100a):
// TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (one elected thread,
// kind::f16, bf16 operands, fp32 accumulators in TMEM) -> tcgen05.ld -> fused epilogue (epilogue.cuh).
//
//   gemm_tn_tc_kernel    acc[M,N] = A[M,K].B[N,K]^T, both operands K-major (nn.Linear layout).
//                        Persistent (one CTA per SM, static round-robin tiles, N fastest so the A tile
//                        is shared through L2), STAGES-deep smem ring, TWO TMEM accumulator buffers so
//                        the epilogue of tile i overlaps the MMAs of tile i+1.  Warp roles: 0 = TMA
//                        producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue (8 warps: the
//                        exact-erf GELU epilogue at K=C=96 costs more issue slots than the MMAs).
//   gemm_wgrad_tc_kernel out[N1,N2] = X[M,N1]^T.Y[M,N2]: both operands MN-major straight from the
//                        row-major activations (no transposes), split-K over M with fp32 partials
//                        (deterministic), bias gradient = column sums of X computed BY THE TENSOR CORE
//                        with a constant all-ones B tile (one extra N=16 MMA per k-step).
//
// Descriptor encodings follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor"
// tables (bit positions cross-checked against cute/arch/mma_sm100_desc.hpp in the vendored CUTLASS tree).
#include "common.cuh"
#include "epilogue.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace cnx {
int reduce_partials2(const float* pa, int64_t La, float* oa, const float* pb, int64_t Lb, float* ob, int P, int accumulate,
                     cudaStream_t s);
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;                       // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = (kFirstEpiWarp + kEpiWarps) * 32;

// ---- PTX wrappers ------------------------------------------------------------------------------
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
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: one MMA spans two SMs (M = 256), each CTA feeds its own 128 rows of A and half of
// the B tile, and holds 128 rows of the accumulator in its own TMEM ---------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's even CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster.  RELAXED: the only thing the waiter (the MMA issuer)
// consumes is the TMEM accumulator, ordered by tcgen05.fence::before_thread_sync; a release at cluster scope would drain
// every outstanding global store of the warp first (measured: 22 % of all warp stalls in the dGELU kernel)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(bar), "r"(rank)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -------------------------------------------------------------------------------
// shared-memory matrix descriptor: [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
// [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: c_format F32 (bit 4), a/b format BF16 (bits 7, 10), majors (15, 16),
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// BN = columns of one output tile; NCTA = 1: one CTA owns a 128 x BN tile.  NCTA = 2: a CTA PAIR (cluster of 2, one
// tcgen05.mma.cta_group::2 spans both SMs) owns a 256 x BN tile — each CTA stages its own 128 rows of A and HALF of the
// B tile, so per flop a CTA moves half the bytes through L2 and shared memory that the single-CTA tile does.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
// L2 prefetch of one box of a tensor map (no shared memory, no barrier): the later TMA load of the same box hits in L2
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ float exp2f_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Phi(x) - 0.5 = xc * P(xc^2) on |x| <= 4.5 (clamped beyond: |Phi - {0,1}| < 3.4e-6 there), degree-9 minimax fit of the
// normal CDF, max abs error 1.4e-5 (profiles/fit_phi.py) — far below the bf16 resolution of the stored results.  No MUFU, no
// division; evaluated two elements at a time with the packed fp32x2 FMA so the epilogue costs half the issue slots.
// copysign(min(|x|, r), x) in ONE instruction (min.xorsign.abs: sign = sign(x) ^ sign(r), r > 0)
__device__ __forceinline__ float clamp_sym(float x, float r) {
  float y;
  asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(r));
  return y;
}
__device__ __forceinline__ float2 phi_minus_half2(float2 x) {
  // symmetric input clamp: one FMNMX per element instead of a min and a max.  No clamp of the result: the fit leaves
  // Phi in [-1.4e-5, 1 + 1.4e-5], i.e. |g - GELU(x)| <= 1.4e-5 |x| <= 6.3e-5 on the clamped range — four FMNMX per pair less
  // in an epilogue that is bound by its instruction count (profiles/r02p_*: 8 of ~35 math instructions per pair were clamps)
  const float R = 4.5f;
  float2 xc = make_float2(clamp_sym(x.x, R), clamp_sym(x.y, R));
  const float2 t = __fmul2_rn(xc, xc);
  float2 p = make_float2(-1.726107430e-12f, -1.726107430e-12f);
  p = __ffma2_rn(p, t, make_float2(2.022235545e-10f, 2.022235545e-10f));
  p = __ffma2_rn(p, t, make_float2(-1.056632382e-08f, -1.056632382e-08f));
  p = __ffma2_rn(p, t, make_float2(3.278768088e-07f, 3.278768088e-07f));
  p = __ffma2_rn(p, t, make_float2(-6.813567067e-06f, -6.813567067e-06f));
  p = __ffma2_rn(p, t, make_float2(1.017335529e-04f, 1.017335529e-04f));
  p = __ffma2_rn(p, t, make_float2(-1.142722919e-03f, -1.142722919e-03f));
  p = __ffma2_rn(p, t, make_float2(9.891897850e-03f, 9.891897850e-03f));
  p = __ffma2_rn(p, t, make_float2(-6.642163740e-02f, -6.642163740e-02f));
  p = __ffma2_rn(p, t, make_float2(3.989305611e-01f, 3.989305611e-01f));
  return __fmul2_rn(p, xc);
}
// g = GELU(x) = x * Phi(x); if WITH_GRAD also gp = GELU'(x) = Phi(x) + x * pdf(x)
template <bool WITH_GRAD>
__device__ __forceinline__ void gelu_pair(float2 x, float2& g, float2& gp) {
  const float2 s = phi_minus_half2(x);
  const float2 phi = __fadd2_rn(s, make_float2(0.5f, 0.5f));
  g = __fmul2_rn(x, phi);
  if (WITH_GRAD) {
    // pdf(x) = exp2(-x^2 * log2(e)/2) / sqrt(2 pi)
    const float2 t = __fmul2_rn(x, x);
    const float kc = -0.72134752044448170368f;
    const float2 ta = __fmul2_rn(t, make_float2(kc, kc));
    float2 e;
    e.x = exp2f_fast(ta.x);
    e.y = exp2f_fast(ta.y);
    const float2 xk = __fmul2_rn(x, make_float2(0.39894228040143267794f, 0.39894228040143267794f));
    gp = __ffma2_rn(xk, e, phi);
  }
}

// fp32-accurate GELU for the split-operand ("x3") forward: Phi(h) through erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7)
// on MUFU rcp / ex2, 15 instructions instead of erff()'s ~40.  q = 0.5 * (a1 t + ... + a5 t^5) exp(-h^2/2), t = 1/(1 + p|h|/sqrt2);
// Phi = h >= 0 ? 1 - q : q.  Measured against float64 over [-8, 8]: |g - GELU(h)| <= 4.3e-7.
__device__ __forceinline__ float gelu_as(float h) {
  const float t = __frcp_rn(fmaf(0.23164190398f, fabsf(h), 1.0f));          // p / sqrt(2)
  float poly = fmaf(0.5307027145f, t, -0.7265760135f);
  poly = fmaf(poly, t, 0.7107068705f);
  poly = fmaf(poly, t, -0.142248368f);
  poly = fmaf(poly, t, 0.127414796f);
  const float q = poly * t * exp2f_fast(h * h * -0.72134752044448170368f);
  const float phi = h >= 0.f ? 1.0f - q : q;
  return h * phi;
}

// the same on two elements at a time: packed fp32x2 FMAs for everything but the two MUFU pairs and the sign select
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 gelu_as2(float2 h) {
  const float2 d = __ffma2_rn(make_float2(fabsf(h.x), fabsf(h.y)), make_float2(0.23164190398f, 0.23164190398f), make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_fast(d.x), rcp_fast(d.y));
  float2 p = __ffma2_rn(make_float2(0.5307027145f, 0.5307027145f), t, make_float2(-0.7265760135f, -0.7265760135f));
  p = __ffma2_rn(p, t, make_float2(0.7107068705f, 0.7107068705f));
  p = __ffma2_rn(p, t, make_float2(-0.142248368f, -0.142248368f));
  p = __ffma2_rn(p, t, make_float2(0.127414796f, 0.127414796f));
  p = __fmul2_rn(p, t);
  const float2 a = __fmul2_rn(__fmul2_rn(h, h), make_float2(-0.72134752044448170368f, -0.72134752044448170368f));
  const float2 q = __fmul2_rn(p, make_float2(exp2f_fast(a.x), exp2f_fast(a.y)));
  const float2 phi = make_float2(h.x >= 0.f ? 1.0f - q.x : q.x, h.y >= 0.f ? 1.0f - q.y : q.y);
  return __fmul2_rn(h, phi);
}

// the same with the derivative: gp = GELU'(h) = Phi(h) + h * pdf(h), pdf(h) = exp(-h^2/2) / sqrt(2 pi) from the exponential already formed
__device__ __forceinline__ float2 gelu_as2_grad(float2 h, float2& gp) {
  const float2 d = __ffma2_rn(make_float2(fabsf(h.x), fabsf(h.y)), make_float2(0.23164190398f, 0.23164190398f), make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_fast(d.x), rcp_fast(d.y));
  float2 p = __ffma2_rn(make_float2(0.5307027145f, 0.5307027145f), t, make_float2(-0.7265760135f, -0.7265760135f));
  p = __ffma2_rn(p, t, make_float2(0.7107068705f, 0.7107068705f));
  p = __ffma2_rn(p, t, make_float2(-0.142248368f, -0.142248368f));
  p = __ffma2_rn(p, t, make_float2(0.127414796f, 0.127414796f));
  p = __fmul2_rn(p, t);
  const float2 a = __fmul2_rn(__fmul2_rn(h, h), make_float2(-0.72134752044448170368f, -0.72134752044448170368f));
  const float2 e = make_float2(exp2f_fast(a.x), exp2f_fast(a.y));
  const float2 q = __fmul2_rn(p, e);
  const float2 phi = make_float2(h.x >= 0.f ? 1.0f - q.x : q.x, h.y >= 0.f ? 1.0f - q.y : q.y);
  gp = __ffma2_rn(__fmul2_rn(h, make_float2(0.39894228040143267794f, 0.39894228040143267794f)), e, phi);
  return __fmul2_rn(h, phi);
}

constexpr int kSlabBytes = 4096;                 // one epilogue slab: [32 rows][32 columns] of <= 4-byte elements
template <int BN, int NCTA, bool SLAB, int NEPI = kEpiWarps, bool DUAL = false> struct TnCfg {
  static constexpr int THREADS = (kFirstEpiWarp + NEPI) * 32;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = BN / NCTA;                             // B rows staged by one CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB_BYTES = SLAB ? NEPI * 2 * kSlabBytes : 0;        // two slabs per epilogue warp
  static constexpr int RING_BUDGET = 224 * 1024 - SLAB_BYTES;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int NCHUNK = BN / 32;                               // 32-column epilogue chunks per tile
  // BN <= 256: TWO accumulator buffers (the epilogue of tile i overlaps the MMAs of tile i+1).  BN = 384 (CTA pairs only): ONE
  // 384-column accumulator — a CTA then stages 40 KB per k-block for 128 x 384 x 64 MACs (79 MAC/B) where the 256 x 192 tile
  // stages 28 KB for half the MACs (55 MAC/B): the long-K GEMMs with N = 384 were bound by the bytes an SM can take in per
  // clock, not by the tensor pipe (profiles/r02h_*), and their epilogue is a small fraction of a K >= 1024 tile.
  static constexpr int NACC = (BN <= 256) ? 2 : 1;
  // DUAL (EPI_DGELU_RC): two accumulators per tile, [acc | recomputed pre-activation], BN = 128 columns each
  static constexpr int ACC_STRIDE = (BN <= 128 && !DUAL) ? 128 : 256;
  static constexpr int TMEM_COLS = (BN <= 256) ? 2 * ACC_STRIDE : 512;
  static_assert(!DUAL || (BN == 128 && SLAB), "the recompute epilogue is a 128-column slab configuration");
  static_assert(BN <= 256 || (BN == 384 && NCTA == 2 && SLAB), "BN = 384 is a CTA-pair slab-epilogue configuration");
  static constexpr int EPI_COLS = BN / 2;                              // columns per epilogue warp
  static constexpr int CHUNK = (EPI_COLS % 32 == 0) ? 32 : 16;
  static constexpr int SMEM = STAGES * STAGE_BYTES + SLAB_BYTES + 1024 /*align*/ + 512 /*barriers*/;
  static_assert(B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for the 128B swizzle");
  static_assert(2 * STAGES + 5 + 2 * NEPI <= 64, "barrier block");
  static_assert(STAGES >= 2, "ring too small");
  static_assert(!SLAB || BN % 32 == 0, "slab epilogue works on 32-column chunks");
};

// ================================================================================================
template <int BN, int KIND, typename TOUT, int NCTA, bool SLAB, int NEPI>
__global__ void __launch_bounds__((kFirstEpiWarp + NEPI) * 32, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmIn,
                  const __grid_constant__ CUtensorMap tmB2, int64_t M, int64_t N, int64_t K, EpiParams ep) {
  constexpr bool DUAL = (KIND == EPI_DGELU_RC);        // second operand pair: A2 through tmIn, B2 through tmB2
  typedef TnCfg<BN, NCTA, SLAB, NEPI, DUAL> Cfg;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int TM = BM * NCTA;
  constexpr int NGRP = NEPI / 4;                     // epilogue warps per TMEM lane quarter = column groups
  static_assert(SLAB || NEPI == kEpiWarps, "the direct-store epilogue is written for 8 warps");                      // rows of one output tile (per CTA pair when NCTA = 2)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * Cfg::A_BYTES;
  const uint32_t sSlab = sB + STAGES * Cfg::B_BYTES;
  const uint32_t bars = sSlab + Cfg::SLAB_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  auto in_bar = [&](int w, int b) { return bars + 8u * (2 * STAGES + 5 + 2 * w + b); };   // slab-input barriers, per warp
  constexpr bool HAS_IN = SLAB && (KIND == EPI_SCALE_RES || KIND == EPI_DGELU || KIND == EPI_DGELU3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (NCTA == 2) ? cluster_ctarank() : 0u;       // 0 = the pair's leader (issues the MMAs)
  const int64_t m_tiles = (M + TM - 1) / TM, n_tiles = (N + BN - 1) / BN;
  const int64_t num_tiles = m_tiles * n_tiles;
  const int64_t first_tile = blockIdx.x / NCTA, tile_step = gridDim.x / NCTA;
  const int nkb = (int)((K + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NEPI * NCTA); }
    if (SLAB)
      for (int w = 0; w < NEPI; ++w) { mbar_init(in_bar(w, 0), 1); mbar_init(in_bar(w, 1), 1); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (NCTA == 2) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();                 // the peer's barriers must exist before anything lands on them
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                        // barrier init / TMEM allocation above ran under the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs of a pair: own A rows, own half of the B tile; bytes are counted on the leader's barrier) =====
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int64_t mt = tile / n_tiles, nt = tile - mt * n_tiles;
        const int32_t arow = (int32_t)(mt * TM + rank * BM);
        const int32_t brow = (int32_t)(nt * BN + rank * Cfg::B_ROWS);
        // The ring holds 3-5 k-blocks (the epilogue slabs take 64-128 KB): too little data in flight to cover an HBM round trip
        // under the epilogue's write stream.  Prefetch the A rows (and the rows of the epilogue's input slabs) of a later tile
        // of this CTA into L2, so that the ring's own loads are L2 hits (profiles/r02w_*: the epilogue warps of the K <= 384
        // GEMMs waited 25 % of their time for an accumulator whose operands had not arrived).
        if (ep.pf_tiles > 0) {
          const int64_t tp = tile + (int64_t)ep.pf_tiles * tile_step;
          if (tp < num_tiles) {
            const int64_t mtp = tp / n_tiles, ntp = tp - mtp * n_tiles;
            const int32_t prow = (int32_t)(mtp * TM + rank * BM);
            const int ka = ep.a_wrap > 0 ? (int)ep.a_wrap : (int)K;
            for (int c = 0; c < ka; c += BK) tma_prefetch_2d(&tmA, c, prow);     // (CTAs that share an m-tile repeat it: L2 hits)
            if (HAS_IN) {
              for (int q = 0; q < BM / 32; ++q)
                for (int c = 0; c < Cfg::NCHUNK; ++c) tma_prefetch_2d(&tmIn, (int32_t)(ntp * BN + c * 32), prow + q * 32);
            }
          }
        }
        for (int kk = 0; kk < (DUAL ? 2 : 1) * nkb; ++kk) {
          const int kb = (DUAL && kk >= nkb) ? kk - nkb : kk;
          const CUtensorMap* mA = (DUAL && kk >= nkb) ? &tmIn : &tmA;
          const CUtensorMap* mB = (DUAL && kk >= nkb) ? &tmB2 : &tmB;
          mbar_wait(empty_bar(s), ph ^ 1);
          int32_t acol = kb * BK;
          if (ep.a_wrap > 0 && acol >= ep.a_wrap) acol -= ep.a_wrap;      // third segment of a split operand = its first
          if (NCTA == 2) {
            if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
            tma_load_2d_pair(sA + s * Cfg::A_BYTES, mA, full_bar(s), acol, arow);
            if constexpr (BN == 384) {
              // this CTA's half of the N = 256 MMA's B rows (two 64-row boxes), then its half of the N = 128 MMA's (one box)
              const int32_t b0 = (int32_t)(nt * BN + rank * 128), b1 = (int32_t)(nt * BN + 256 + rank * 64);
              tma_load_2d_pair(sB + s * Cfg::B_BYTES, mB, full_bar(s), kb * BK, b0);
              tma_load_2d_pair(sB + s * Cfg::B_BYTES + 8192, mB, full_bar(s), kb * BK, b0 + 64);
              tma_load_2d_pair(sB + s * Cfg::B_BYTES + 16384, mB, full_bar(s), kb * BK, b1);
            } else {
              tma_load_2d_pair(sB + s * Cfg::B_BYTES, mB, full_bar(s), kb * BK, brow);
            }
          } else {
            mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
            tma_load_2d(sA + s * Cfg::A_BYTES, mA, full_bar(s), acol, arow);
            tma_load_2d(sB + s * Cfg::B_BYTES, mB, full_bar(s), kb * BK, brow);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (the leader CTA's elected thread issues for the pair) =====
      constexpr uint32_t idesc = make_idesc(TM, BN == 384 ? 256 : BN, 0, 0);
      constexpr uint32_t idesc_b = make_idesc(TM, 128, 0, 0);          // BN = 384: the second MMA of a k-step (columns 256..383)
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      for (int64_t tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(as), aph ^ 1);          // epilogue(s) have drained this accumulator buffer
        tc_fence_after();
        for (int kk = 0; kk < (DUAL ? 2 : 1) * nkb; ++kk) {
          const int kb = (DUAL && kk >= nkb) ? kk - nkb : kk;
          // DUAL: the second pass accumulates the recomputed pre-activation in the upper 128 columns of the buffer
          const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE + ((DUAL && kk >= nkb) ? 128u : 0u);
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(sA + s * Cfg::A_BYTES, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sB + s * Cfg::B_BYTES, 16, 1024);
          int64_t krem = K - (int64_t)kb * BK;
          const int kmma = krem >= BK ? BK / 16 : (int)((krem + 15) / 16);
          for (int k = 0; k < kmma; ++k) {
            if (NCTA == 2) umma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            else umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            if constexpr (BN == 384) {
              const uint64_t bdesc2 = make_smem_desc(sB + s * Cfg::B_BYTES + 16384, 16, 1024);
              umma_f16_pair(d_tmem + 256, adesc + (uint64_t)(k * 2), bdesc2 + (uint64_t)(k * 2), idesc_b, (kb | k) != 0);
            }
          }
          if (NCTA == 2) umma_commit_pair(empty_bar(s));   // slot s is free in BOTH CTAs once these MMAs retire
          else umma_commit(empty_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (NCTA == 2) umma_commit_pair(tfull_bar(as));    // accumulator complete -> both epilogues
        else umma_commit(tfull_bar(as));
        if (++as == Cfg::NACC) { as = 0; aph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= kFirstEpiWarp && !SLAB) {
    // ===== epilogue: TMEM -> registers -> fused math -> global (each CTA drains its own 128 accumulator rows) =====
    const int quarter = warp & 3;                    // TMEM lanes [32*quarter, +32) are this warp's
    const int half = (warp - kFirstEpiWarp) >> 2;    // which half of the tile's columns
    int as = 0; uint32_t aph = 0;
    for (int64_t tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int64_t mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int64_t m = mt * TM + rank * BM + quarter * 32 + lane;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + as * Cfg::ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int c = 0; c < Cfg::EPI_COLS; c += Cfg::CHUNK) {
        const int col = half * Cfg::EPI_COLS + c;
        float v[Cfg::CHUNK];
        if constexpr (Cfg::CHUNK == 32) tmem_ld32(t_row + col, *reinterpret_cast<float(*)[32]>(v));
        else tmem_ld16(t_row + col, *reinterpret_cast<float(*)[16]>(v));
        tmem_ld_wait();
        const int64_t n = nt * BN + col;
        if (m < M) {
#pragma unroll
          for (int j = 0; j < Cfg::CHUNK; j += 8)
            if (n + j < N) epilogue_store8<KIND, TOUT>(ep, m, n + j, *reinterpret_cast<float(*)[8]>(&v[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(tempty_bar(as), 0);     // the leader's MMA issuer waits for both CTAs
        else mbar_arrive(tempty_bar(as));
      }
      if (++as == Cfg::NACC) { as = 0; aph ^= 1; }
    }
  } else if (warp >= kFirstEpiWarp && SLAB) {
    // ===== slab epilogue: every epilogue warp is autonomous.  It owns TMEM lanes [32q, 32q+32) and every other 32-column
    // chunk of the tile; per chunk: (residual / GELU' slab prefetched by ITS OWN TMA load) -> tcgen05.ld -> fused math in
    // registers -> swizzled shared-memory slab -> ITS OWN TMA store.  Global traffic is whole 128-byte rows moved by the TMA
    // unit (ragged M / N edges clipped by the tensor maps); no CTA-wide barrier, no per-thread global addressing. =====
    const int ew = warp - kFirstEpiWarp;
    const int quarter = warp & 3;
    const int half = ew >> 2;                          // column group: chunks half, half + NGRP, ...
    const uint32_t slab0 = sSlab + (uint32_t)ew * 2u * kSlabBytes;
    constexpr uint32_t ROWB = 32 * sizeof(TOUT);                       // slab row: 128 B (fp32) or 64 B (bf16)
    const uint32_t row_off = (uint32_t)lane * ROWB;
    const uint32_t swz = (sizeof(TOUT) == 4) ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
    constexpr int NV = (int)ROWB / 16;                                 // 16-byte vectors per slab row
    constexpr uint32_t IN_BYTES = (KIND == EPI_DGELU3) ? 4096u : 32u * ROWB;   // EPI_DGELU3: the input slab is fp32 [32][32]
    int as = 0; uint32_t aph = 0;
    int buf = 0; uint32_t inph[2] = {0u, 0u};
    int64_t t_cur = first_tile; int c_cur = half;
    // Column chunks go to the NGRP warps of a lane quarter round-robin.  Where NCHUNK is not a multiple of NGRP (BN = 192 with
    // 16 epilogue warps: 6 chunks over 4 groups) the round-robin continues ACROSS tiles instead of restarting at every tile —
    // restarting gave two groups 2 chunks and two groups 1 chunk of every tile, i.e. half the warps idle a quarter of the time
    // (profiles/r02w_ncu_source_stalls.txt: 25 % of the x3 fc1 epilogue warps' samples waiting for the next accumulator).
    constexpr bool ROT = (Cfg::NCHUNK % NGRP) != 0;
    static_assert(!ROT || NGRP <= Cfg::NCHUNK, "every epilogue warp must visit every tile");
    bool new_tile = true;
    auto chunk_coords = [&](int64_t tile, int c, int32_t& x, int32_t& y) {
      const int64_t mt = tile / n_tiles, nt = tile - mt * n_tiles;
      x = (int32_t)(nt * BN + c * 32);
      y = (int32_t)(mt * TM + rank * BM + quarter * 32);
    };
    if (HAS_IN && lane == 0 && t_cur < num_tiles) {
      int32_t x, y;
      chunk_coords(t_cur, c_cur, x, y);
      mbar_expect_tx(in_bar(ew, 0), IN_BYTES);
      tma_load_2d(slab0, &tmIn, in_bar(ew, 0), x, y);
    }
    while (t_cur < num_tiles) {
      int64_t t_nxt = t_cur; int c_nxt = c_cur + NGRP;
      if (c_nxt >= Cfg::NCHUNK) { c_nxt = ROT ? c_nxt - Cfg::NCHUNK : half; t_nxt += tile_step; }
      const uint32_t slab = slab0 + (uint32_t)buf * kSlabBytes;
      if (lane == 0) {
        if (HAS_IN) {
          bulk_wait_read0();                         // the store that last read the other slab is done with it
          if (t_nxt < num_tiles) {
            int32_t x, y;
            chunk_coords(t_nxt, c_nxt, x, y);
            mbar_expect_tx(in_bar(ew, buf ^ 1), IN_BYTES);
            tma_load_2d(slab0 + (uint32_t)(buf ^ 1) * kSlabBytes, &tmIn, in_bar(ew, buf ^ 1), x, y);
          }
        } else if (KIND == EPI_BIAS_GELU3 && ep.out0 != nullptr) {
          bulk_wait_read0();                         // with the fp32 GELU' output a chunk uses BOTH slabs: all earlier stores done
        } else {
          bulk_wait_read1();                         // the store issued two chunks ago has finished reading this slab
        }
      }
      if (new_tile) {                                // first chunk of a tile for this warp
        mbar_wait(tfull_bar(as), aph);
        tc_fence_after();
      }
      new_tile = (t_nxt != t_cur);
      float v[32];
      float u[DUAL ? 32 : 1];       // DUAL: the recomputed pre-activation chunk (upper 128 columns of the accumulator buffer)
      tmem_ld32(tmem_base + as * Cfg::ACC_STRIDE + ((uint32_t)(quarter * 32) << 16) + c_cur * 32, v);
      if constexpr (DUAL)
        tmem_ld32(tmem_base + as * Cfg::ACC_STRIDE + ((uint32_t)(quarter * 32) << 16) + 128 + c_cur * 32,
                  *reinterpret_cast<float(*)[32]>(u));
      tmem_ld_wait();
      if (t_nxt != t_cur) {         // last chunk of this tile for this warp: hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_cluster(tempty_bar(as), 0);
          else mbar_arrive(tempty_bar(as));
        }
        if (++as == Cfg::NACC) { as = 0; aph ^= 1; }
      }
      const int64_t n0 = (t_cur % n_tiles) * BN + c_cur * 32;
      const int64_t m = (t_cur / n_tiles) * TM + rank * BM + quarter * 32 + lane;
      // ---- fused math on this lane's row: 32 columns ----
      if (KIND == EPI_PLAIN) {
        if (ep.bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (sizeof(TOUT) == 4 && ep.round_bf16) {        // fp32 output holding bf16-rounded values (EpiParams::round_bf16)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint32_t u = pack_bf16(v[i], v[i + 1]);
            v[i] = bf16_lo(u); v[i + 1] = bf16_hi(u);
          }
        }
      } else if (KIND == EPI_SCALE_RES) {
        float sc = 1.0f;
        if (ep.dp) sc = __ldg(ep.dp + (m < M ? m : M - 1) / ep.rows_per_sample);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (ep.bias) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + i));
          if (ep.gamma) g4 = __ldg(reinterpret_cast<const float4*>(ep.gamma + n0 + i));
          v[i] = sc * (g4.x * (v[i] + b4.x));
          v[i + 1] = sc * (g4.y * (v[i + 1] + b4.y));
          v[i + 2] = sc * (g4.z * (v[i + 2] + b4.z));
          v[i + 3] = sc * (g4.w * (v[i + 3] + b4.w));
        }
      }
      if constexpr (DUAL) {
        // dh = acc * GELU'(h), h = round_bf16(xn.W1^T + b1) exactly as the forward's fc1 epilogue rounded it
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + i));
          const float2 ha = __fadd2_rn(make_float2(u[i], u[i + 1]), make_float2(b4.x, b4.y));
          const float2 hb = __fadd2_rn(make_float2(u[i + 2], u[i + 3]), make_float2(b4.z, b4.w));
          const uint32_t ua = pack_bf16(ha.x, ha.y), ub = pack_bf16(hb.x, hb.y);
          float2 ga, gb, da, db;
          gelu_pair<true>(make_float2(bf16_lo(ua), bf16_hi(ua)), ga, da);
          gelu_pair<true>(make_float2(bf16_lo(ub), bf16_hi(ub)), gb, db);
          // GELU'(h) rounded to bf16 as the stored tensor of the other path is, then the same multiply
          const uint32_t pa = pack_bf16(da.x, da.y), pb = pack_bf16(db.x, db.y);
          v[i] *= bf16_lo(pa); v[i + 1] *= bf16_hi(pa); v[i + 2] *= bf16_lo(pb); v[i + 3] *= bf16_hi(pb);
        }
      }
      uint32_t pg[16], pd[16];
      const bool want_gp = (KIND == EPI_BIAS_GELU3) || (KIND == EPI_DGELU3) || ((KIND == EPI_BIAS_GELU) && (ep.out0 != nullptr));
      if (KIND == EPI_BIAS_GELU3) {
        // fp32 h, exact-erf GELU in fp32, g split into two bf16 pieces: g ~ hi + mid to 2^-17 relative.  Training (ep.out0): the
        // fp32 GELU'(h) leaves through the OTHER slab of this warp ([32 rows][128 B], fp32 swizzle) as it is formed
        const bool gp3 = ep.out0 != nullptr;
        const uint32_t oslab = slab0 + (uint32_t)(buf ^ 1) * kSlabBytes + (uint32_t)lane * 128u;
        if (gp3) __syncwarp();                       // lane 0's wait above covers the whole warp before the first store
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + i));
          float2 ga, gb;
          if (gp3) {
            float2 da, db;
            ga = gelu_as2_grad(__fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y)), da);
            gb = gelu_as2_grad(__fadd2_rn(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w)), db);
            sts128(oslab + ((((uint32_t)i >> 2) ^ (uint32_t)(lane & 7)) << 4), __float_as_uint(da.x), __float_as_uint(da.y),
                   __float_as_uint(db.x), __float_as_uint(db.y));
          } else {
            ga = gelu_as2(__fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y)));
            gb = gelu_as2(__fadd2_rn(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w)));
          }
          const uint32_t ha = pack_bf16(ga.x, ga.y), hb = pack_bf16(gb.x, gb.y);
          pg[i / 2] = ha;
          pg[i / 2 + 1] = hb;
          const float2 ra = __fadd2_rn(ga, make_float2(-bf16_lo(ha), -bf16_hi(ha)));
          const float2 rb = __fadd2_rn(gb, make_float2(-bf16_lo(hb), -bf16_hi(hb)));
          pd[i / 2] = pack_bf16(ra.x, ra.y);
          pd[i / 2 + 1] = pack_bf16(rb.x, rb.y);
        }
      }
      if (KIND == EPI_BIAS_GELU) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + i));
          const float2 ha = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y));
          const float2 hb = __fadd2_rn(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w));
          const uint32_t ua = pack_bf16(ha.x, ha.y);        // h rounded to bf16, as autocast's Linear output
          const uint32_t ub = pack_bf16(hb.x, hb.y);
          float2 ga, gb, da, db;
          if (want_gp) {
            gelu_pair<true>(make_float2(bf16_lo(ua), bf16_hi(ua)), ga, da);
            gelu_pair<true>(make_float2(bf16_lo(ub), bf16_hi(ub)), gb, db);
            pd[i / 2] = pack_bf16(da.x, da.y);              // GELU'(h): all that backward needs of h
            pd[i / 2 + 1] = pack_bf16(db.x, db.y);
          } else {
            gelu_pair<false>(make_float2(bf16_lo(ua), bf16_hi(ua)), ga, da);
            gelu_pair<false>(make_float2(bf16_lo(ub), bf16_hi(ub)), gb, db);
          }
          pg[i / 2] = pack_bf16(ga.x, ga.y);
          pg[i / 2 + 1] = pack_bf16(gb.x, gb.y);
        }
      }
      if constexpr (KIND == EPI_DGELU3) {
        // fp32 GELU'(h) slab ([32 rows][128 B], fp32 swizzle) -> dh = acc * gp in fp32 -> two bf16 pieces
        mbar_wait(in_bar(ew, buf), inph[buf]);
        inph[buf] ^= 1;
        const uint32_t irow = slab + (uint32_t)lane * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t a0, a1, a2, a3;
          lds128(irow + (((uint32_t)j ^ (uint32_t)(lane & 7)) << 4), a0, a1, a2, a3);
          const float2 da = __fmul2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(__uint_as_float(a0), __uint_as_float(a1)));
          const float2 db = __fmul2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(__uint_as_float(a2), __uint_as_float(a3)));
          const uint32_t ha = pack_bf16(da.x, da.y), hb = pack_bf16(db.x, db.y);
          pg[2 * j] = ha;
          pg[2 * j + 1] = hb;
          const float2 ra = __fadd2_rn(da, make_float2(-bf16_lo(ha), -bf16_hi(ha)));
          const float2 rb = __fadd2_rn(db, make_float2(-bf16_lo(hb), -bf16_hi(hb)));
          pd[2 * j] = pack_bf16(ra.x, ra.y);
          pd[2 * j + 1] = pack_bf16(rb.x, rb.y);
        }
      } else if (HAS_IN) {
        mbar_wait(in_bar(ew, buf), inph[buf]);
        inph[buf] ^= 1;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          uint32_t a0, a1, a2, a3;
          lds128(slab + row_off + (((uint32_t)j ^ swz) << 4), a0, a1, a2, a3);
          if (sizeof(TOUT) == 4) {
            if (KIND == EPI_SCALE_RES) {
              v[4 * j] += __uint_as_float(a0); v[4 * j + 1] += __uint_as_float(a1);
              v[4 * j + 2] += __uint_as_float(a2); v[4 * j + 3] += __uint_as_float(a3);
            } else {
              v[4 * j] *= __uint_as_float(a0); v[4 * j + 1] *= __uint_as_float(a1);
              v[4 * j + 2] *= __uint_as_float(a2); v[4 * j + 3] *= __uint_as_float(a3);
            }
          } else {
            const uint32_t a[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (KIND == EPI_SCALE_RES) { v[8 * j + 2 * q] += bf16_lo(a[q]); v[8 * j + 2 * q + 1] += bf16_hi(a[q]); }
              else { v[8 * j + 2 * q] *= bf16_lo(a[q]); v[8 * j + 2 * q + 1] *= bf16_hi(a[q]); }
            }
          }
        }
      }
      __syncwarp();                                  // lane 0's wait on the slab's previous store covers the whole warp
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const uint32_t addr = slab + row_off + (((uint32_t)j ^ swz) << 4);
        if (KIND == EPI_BIAS_GELU || KIND == EPI_BIAS_GELU3 || KIND == EPI_DGELU3) {   // two bf16 outputs share the slab: g (hi) in the first 2 KB, GELU' (mid) in the second
          sts128(addr, pg[4 * j], pg[4 * j + 1], pg[4 * j + 2], pg[4 * j + 3]);
          if (want_gp) sts128(addr + 2048u, pd[4 * j], pd[4 * j + 1], pd[4 * j + 2], pd[4 * j + 3]);
        } else if (sizeof(TOUT) == 4)
          sts128(addr, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                 __float_as_uint(v[4 * j + 3]));
        else
          sts128(addr, pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                 pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        int32_t x, y;
        chunk_coords(t_cur, c_cur, x, y);
        tma_store_2d(&tmOut, slab, x, y);
        if (KIND == EPI_DGELU3) {                  // [hi | mid] column blocks of the [M, 2N] split operand
          tma_store_2d(&tmOut, slab + 2048u, x + (int32_t)N, y);
        } else if (KIND == EPI_BIAS_GELU3) {
          tma_store_2d(&tmOut, slab + 2048u, x + (int32_t)N, y);
          if (ep.out0 != nullptr) tma_store_2d(&tmIn, slab0 + (uint32_t)(buf ^ 1) * kSlabBytes, x, y);      // fp32 GELU'(h) [M, N]
        } else if (want_gp) {
          tma_store_2d(&tmIn, slab + 2048u, x, y);
        }
        bulk_commit();
      }
      buf ^= 1;
      t_cur = t_nxt; c_cur = c_nxt;
    }
    if (lane == 0) bulk_wait0();
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();                 // no CTA may exit (or free TMEM) while its peer still uses it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}


// ================================================================================================
// wgrad: part[split][i][j] = sum_{m in split} X[m,i] * Y[m,j] ; cs_part[split][i] = sum_m X[m,i]
// grid = (i_tiles * j_tiles, splits).  A = X^T and B = Y^T are both MN-major operands: the TMA box is
// [64 m-rows][64 contiguous channels] (128 B rows, 128B swizzle), two boxes per operand per stage.
// ================================================================================================
template <int BN> struct WgCfg {
  static constexpr int A_BYTES = 2 * 64 * BK * 2;      // two 64-wide boxes
  static constexpr int B_BYTES = 2 * 64 * BK * 2;
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 256;                // BN (<=128) accumulator + 16 colsum columns at 128
  static constexpr int CS_COL = 128;
  static constexpr int EPI_COLS = BN / 2;
  static constexpr int CHUNK = (EPI_COLS % 32 == 0) ? 32 : 16;
  static constexpr int ONES_BYTES = 2048;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + ONES_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, int64_t Mtot,
                     int64_t N1, int64_t N2, int kb_per_split, float* __restrict__ part, float* __restrict__ cs_part, int passes) {
  typedef WgCfg<BN> Cfg;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * Cfg::A_BYTES;
  const uint32_t sOnes = sB + STAGES * Cfg::B_BYTES;
  const uint32_t bars = sOnes + Cfg::ONES_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j_tiles = (N2 + BN - 1) / BN;
  const int64_t it = blockIdx.x / j_tiles, jt = blockIdx.x - it * j_tiles;
  const int split = blockIdx.y;
  // passes = 3: split operands X2 = [hi | mid] (2 N1 columns), Y2 = [hi | mid] (2 N2 columns): ONE K loop over the three products
  // hi.hi, mid.hi, hi.mid (pass p reads X at column offset (p == 1) N1 and Y at (p == 2) N2) instead of three launches with
  // three sets of split-K partials; the column sums of X come from passes 0 and 1 (hi + mid)
  const int kb_rows = (int)((Mtot + BK - 1) / BK);
  const int kb_total = passes * kb_rows;
  const int kb0 = split * kb_per_split;
  int kb1 = kb0 + kb_per_split;
  if (kb1 > kb_total) kb1 = kb_total;
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;
  const bool do_colsum = (cs_part != nullptr) && (jt == 0);

  // constant all-ones B tile for the bias-gradient MMA (any layout of ones is ones)
  {
    uint8_t* gen = smem_raw + (sOnes - smem_u32(smem_raw));
    for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += kThreads) reinterpret_cast<uint32_t*>(gen)[i] = 0x3F803F80u;
    fence_proxy_async();     // generic-proxy writes must be visible to the tensor core's async proxy
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int g = kb0 + kb, pass = g / kb_rows;
        const int32_t mrow = (g - pass * kb_rows) * BK;
        const int32_t xo = (pass == 1) ? (int32_t)N1 : 0, yo = (pass == 2) ? (int32_t)N2 : 0;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), Cfg::A_BYTES + Cfg::B_BYTES);
        tma_load_2d(sA + s * Cfg::A_BYTES, &tmX, full_bar(s), (int32_t)(it * BM) + xo, mrow);
        tma_load_2d(sA + s * Cfg::A_BYTES + 8192, &tmX, full_bar(s), (int32_t)(it * BM + 64) + xo, mrow);
        tma_load_2d(sB + s * Cfg::B_BYTES, &tmY, full_bar(s), (int32_t)(jt * BN) + yo, mrow);
        tma_load_2d(sB + s * Cfg::B_BYTES + 8192, &tmY, full_bar(s), (int32_t)(jt * BN + 64) + yo, mrow);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, 1, 1);
      constexpr uint32_t idesc_ones = make_idesc(BM, 16, 1, 1);
      const uint64_t odesc = make_smem_desc(sOnes, 8192, 1024);
      int s = 0; uint32_t ph = 0;
      uint32_t cs_acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        // MN-major SW128: 64-channel atoms are LBO = 8192 B apart (one TMA box), 8-row k-groups SBO = 1024 B
        const uint64_t adesc = make_smem_desc(sA + s * Cfg::A_BYTES, 8192, 1024);
        const uint64_t bdesc = make_smem_desc(sB + s * Cfg::B_BYTES, 8192, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // one UMMA_K = 16 rows = two 8-row groups = 2048 B further into the tile
          umma_f16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (kb | k) != 0);
          if (do_colsum && (kb0 + kb) < 2 * kb_rows) {       // (passes = 1: always; passes = 3: the hi and mid passes of X)
            umma_f16(tmem_base + Cfg::CS_COL, adesc + (uint64_t)(k * 128), odesc, idesc_ones, cs_acc);
            cs_acc = 1;
          }
        }
        umma_commit(empty_bar(s));
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  } else if (warp >= kFirstEpiWarp) {
    const int quarter = warp & 3;
    const int half = (warp - kFirstEpiWarp) >> 2;
    const int64_t i = it * BM + quarter * 32 + lane;
    float* po = part + (int64_t)split * N1 * N2;
    if (nkb > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
    for (int c = 0; c < Cfg::EPI_COLS; c += Cfg::CHUNK) {
      const int col = half * Cfg::EPI_COLS + c;
      float v[Cfg::CHUNK];
      if (nkb > 0) {
        if constexpr (Cfg::CHUNK == 32) tmem_ld32(t_row + col, *reinterpret_cast<float(*)[32]>(v));
        else tmem_ld16(t_row + col, *reinterpret_cast<float(*)[16]>(v));
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < Cfg::CHUNK; ++j) v[j] = 0.f;
      }
      const int64_t jj = jt * BN + col;
      if (i < N1) {
#pragma unroll
        for (int j = 0; j < Cfg::CHUNK; j += 8)
          if (jj + j < N2) store8(po + i * N2 + jj + j, *reinterpret_cast<float(*)[8]>(&v[j]));
      }
    }
    if (do_colsum && half == 0) {
      float v[16];
      if (nkb > 0 && kb0 < 2 * kb_rows) {              // this split issued at least one column-sum MMA
        tmem_ld16(t_row + Cfg::CS_COL, v);
        tmem_ld_wait();
      } else {
        v[0] = 0.f;
      }
      if (i < N1) cs_part[(int64_t)split * N1 + i] = v[0];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ================================================================================================
// Fused MLP forward for the HBM-bound stages (C <= 192), no-grad pass:
//     out = x + dp * gamma * ( GELU(xn . W1^T + b1) . W2^T + b2 )
// in ONE kernel: the [M, 4C] hidden activation never leaves the SM.  Per 128-row tile, the hidden dimension is walked in
// 64-column chunks: GEMM1 (tcgen05, K = C) -> TMEM acc1 (double-buffered) -> 16 epilogue warps (bias + GELU, bf16) write the
// chunk into shared memory in the K-major 128B-swizzled layout of an MMA A operand -> GEMM2 (K = 64) accumulates the [128, C]
// output tile in TMEM acc2.  The final epilogue adds bias / layer-scale / drop-path / residual through per-warp TMA slabs.
// HBM traffic per token: C*2 (xn) + C*4 (x) + C*4 (out) bytes instead of that plus 2 * 4C*2 for the hidden round trip.
// ================================================================================================
constexpr int kFuHCH = 64;                         // hidden columns per chunk
constexpr int kFuEpiWarps = 16;
constexpr int kFuThreads = (kFirstEpiWarp + kFuEpiWarps) * 32;

template <int C> struct FuCfg {
  static constexpr int KBX = (C + 63) / 64;                 // 64-wide K boxes of the xn / W1 tiles
  static constexpr int XN_BYTES = KBX * BM * 128;           // [128 rows][64 k] boxes
  static constexpr int W1_BYTES = KBX * kFuHCH * 128;       // [64 hidden rows][64 k] boxes
  static constexpr int W2_BYTES = C * 128;                  // [C out rows][64 hidden k]
  static constexpr int WST_BYTES = W1_BYTES + W2_BYTES;
  static constexpr int G_BYTES = BM * 128;                  // [128 rows][64 hidden] bf16
  static constexpr int NXBUF = (C <= 96) ? 2 : 1;
  static constexpr int NCH32 = C / 32;                      // 32-column chunks of the output tile
  static constexpr int NSLABW = (NCH32 * 4 <= kFuEpiWarps) ? NCH32 * 4 : 8;   // warps that run the final epilogue
  static constexpr int ITEMS = NCH32 * 4 / NSLABW;          // (quarter, chunk) items per such warp
  static constexpr int NBARS = 2 * NXBUF + 4 + 4 + 4 + 2 + NSLABW + 1;
  static constexpr int SMEM = NXBUF * XN_BYTES + 2 * WST_BYTES + 2 * G_BYTES + NSLABW * kSlabBytes + 1024 + 8 * NBARS + 64;
  static constexpr int ACC2_COL = 128;
  static_assert(C % 32 == 0 && C <= 256 && (NCH32 * 4) % NSLABW == 0, "fused MLP tile shape");
  static_assert(SMEM <= 227 * 1024, "fused MLP does not fit in shared memory");
};

template <int C>
__global__ void __launch_bounds__(kFuThreads, 1)
mlp_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmXn, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRes,
                     const __grid_constant__ CUtensorMap tmOut, int64_t M, const float* __restrict__ b1,
                     const float* __restrict__ b2, const float* __restrict__ gamma, const float* __restrict__ dp,
                     int64_t rows_per_sample) {
  typedef FuCfg<C> Cfg;
  constexpr int NJ = 4 * C / kFuHCH;                 // hidden chunks per tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sXn = base;
  const uint32_t sW = sXn + Cfg::NXBUF * Cfg::XN_BYTES;
  const uint32_t sG = sW + 2 * Cfg::WST_BYTES;
  const uint32_t sSlab = sG + 2 * Cfg::G_BYTES;
  const uint32_t bars = sSlab + Cfg::NSLABW * kSlabBytes;
  int bi = 0;
  auto nb = [&](int n) { const uint32_t a = bars + 8u * bi; bi += n; return a; };
  const uint32_t xn_full = nb(Cfg::NXBUF), xn_empty = nb(Cfg::NXBUF);
  const uint32_t w_full = nb(2), w_empty = nb(2);
  const uint32_t a1_full = nb(2), a1_empty = nb(2);
  const uint32_t g_full = nb(2), g_empty = nb(2);
  const uint32_t a2_full = nb(1), a2_empty = nb(1);
  const uint32_t res_full = nb(Cfg::NSLABW);
  const uint32_t tmem_slot = nb(1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (M + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmXn); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmRes); tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::NXBUF; ++i) { mbar_init(xn_full + 8 * i, 1); mbar_init(xn_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1);
      mbar_init(a1_full + 8 * i, 1); mbar_init(a1_empty + 8 * i, kFuEpiWarps);
      mbar_init(g_full + 8 * i, kFuEpiWarps); mbar_init(g_empty + 8 * i, 1);
    }
    mbar_init(a2_full, 1); mbar_init(a2_empty, Cfg::NSLABW);
    for (int i = 0; i < Cfg::NSLABW; ++i) mbar_init(res_full + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: the xn tile once per tile, one (W1 chunk, W2 chunk) weight stage per hidden chunk =====
      uint32_t xc = 0, wc = 0;
      for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        const uint32_t xb = xc % Cfg::NXBUF, xu = xc / Cfg::NXBUF;
        mbar_wait(xn_empty + 8 * xb, (xu & 1) ^ 1);
        mbar_expect_tx(xn_full + 8 * xb, Cfg::XN_BYTES);
        for (int kb = 0; kb < Cfg::KBX; ++kb)
          tma_load_2d(sXn + xb * Cfg::XN_BYTES + kb * (BM * 128), &tmXn, xn_full + 8 * xb, kb * 64, (int32_t)(tile * BM));
        ++xc;
        for (int j = 0; j < NJ; ++j, ++wc) {
          const uint32_t ws = wc & 1, wu = wc >> 1;
          mbar_wait(w_empty + 8 * ws, (wu & 1) ^ 1);
          mbar_expect_tx(w_full + 8 * ws, Cfg::WST_BYTES);
          for (int kb = 0; kb < Cfg::KBX; ++kb)
            tma_load_2d(sW + ws * Cfg::WST_BYTES + kb * (kFuHCH * 128), &tmW1, w_full + 8 * ws, kb * 64, j * kFuHCH);
          tma_load_2d(sW + ws * Cfg::WST_BYTES + Cfg::W1_BYTES, &tmW2, w_full + 8 * ws, j * kFuHCH, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: GEMM1(j) runs one chunk ahead of GEMM2(j-1) so the GELU epilogue of chunk j-1 overlaps it =====
      constexpr uint32_t idesc1 = make_idesc(BM, kFuHCH, 0, 0);
      constexpr uint32_t idesc2 = make_idesc(BM, C, 0, 0);
      uint32_t xc = 0, cc = 0, tc_ = 0;             // tile / chunk counters (chunk counter is shared by acc1, G and weight rings)
      for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++xc, ++tc_) {
        const uint32_t xb = xc % Cfg::NXBUF, xu = xc / Cfg::NXBUF;
        mbar_wait(xn_full + 8 * xb, xu & 1);
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) {
            const uint32_t c = cc + j, b = c & 1, u = c >> 1;
            mbar_wait(w_full + 8 * b, u & 1);
            mbar_wait(a1_empty + 8 * b, (u & 1) ^ 1);
            tc_fence_after();
            const uint32_t d1 = tmem_base + b * kFuHCH;
#pragma unroll
            for (int kb = 0; kb < Cfg::KBX; ++kb) {
              const uint64_t adesc = make_smem_desc(sXn + xb * Cfg::XN_BYTES + kb * (BM * 128), 16, 1024);
              const uint64_t bdesc = make_smem_desc(sW + b * Cfg::WST_BYTES + kb * (kFuHCH * 128), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (kb * 64 + k * 16 < C) umma_f16(d1, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc1, (kb | k) != 0);
            }
            umma_commit(a1_full + 8 * b);
            if (j == NJ - 1) umma_commit(xn_empty + 8 * xb);
          }
          if (j > 0) {
            const uint32_t c = cc + j - 1, b = c & 1, u = c >> 1;
            mbar_wait(g_full + 8 * b, u & 1);
            if (j == 1) mbar_wait(a2_empty, (tc_ & 1) ^ 1);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(sG + b * Cfg::G_BYTES, 16, 1024);
            const uint64_t bdesc = make_smem_desc(sW + b * Cfg::WST_BYTES + Cfg::W1_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + Cfg::ACC2_COL, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc2, ((j - 1) | k) != 0);
            umma_commit(g_empty + 8 * b);
            umma_commit(w_empty + 8 * b);
            if (j == NJ) umma_commit(a2_full);
          }
        }
        cc += NJ;
      }
    }
    __syncwarp();
  } else if (warp >= kFirstEpiWarp) {
    const int ew = warp - kFirstEpiWarp;
    const int quarter = warp & 3;
    const int cg = ew >> 2;                           // 16 of the chunk's 64 hidden columns
    const int r = quarter * 32 + lane;                // row within the tile
    const uint32_t lane_t = (uint32_t)(quarter * 32) << 16;
    const uint32_t g_row = (uint32_t)r * 128u;
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t slab = sSlab + (uint32_t)ew * kSlabBytes;
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t rswz = (uint32_t)(lane & 7);
    const bool fin = ew < Cfg::NSLABW;
    uint32_t cc = 0, tc_ = 0, rc = 0;                 // chunk counter, tile counter, residual-slab load counter
    for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++tc_) {
      const int32_t y0 = (int32_t)(tile * BM + quarter * 32);
      if (fin && lane == 0) {
        // the residual slab of this warp's first output item: in flight during the whole chunk loop
        bulk_wait_read0();
        mbar_expect_tx(res_full + 8 * ew, kSlabBytes);
        tma_load_2d(slab, &tmRes, res_full + 8 * ew, (int32_t)((ew >> 2) * 32), y0);
      }
      for (int j = 0; j < NJ; ++j, ++cc) {
        const uint32_t b = cc & 1, u = cc >> 1;
        mbar_wait(a1_full + 8 * b, u & 1);
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem_base + b * kFuHCH + lane_t + cg * 16, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_empty + 8 * b);
        uint32_t pg[8];
        const float* bp = b1 + j * kFuHCH + cg * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + i));
          const float2 ha = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y));
          const float2 hb = __fadd2_rn(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w));
          const uint32_t ua = pack_bf16(ha.x, ha.y), ub = pack_bf16(hb.x, hb.y);   // h rounded to bf16 (autocast's Linear output)
          float2 ga, gb, da, db;
          gelu_pair<false>(make_float2(bf16_lo(ua), bf16_hi(ua)), ga, da);
          gelu_pair<false>(make_float2(bf16_lo(ub), bf16_hi(ub)), gb, db);
          pg[i / 2] = pack_bf16(ga.x, ga.y);
          pg[i / 2 + 1] = pack_bf16(gb.x, gb.y);
        }
        mbar_wait(g_empty + 8 * b, (u & 1) ^ 1);      // GEMM2 of two chunks ago has consumed this buffer
        const uint32_t gb_ = sG + b * Cfg::G_BYTES + g_row;
        sts128(gb_ + (((uint32_t)(cg * 2) ^ swz) << 4), pg[0], pg[1], pg[2], pg[3]);
        sts128(gb_ + (((uint32_t)(cg * 2 + 1) ^ swz) << 4), pg[4], pg[5], pg[6], pg[7]);
        fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(g_full + 8 * b);
      }
      if (fin) {
        // ===== final epilogue: this warp's (quarter, 32-column chunk) items of the [128, C] output tile =====
        mbar_wait(a2_full, tc_ & 1);
        tc_fence_after();
        const int64_t m = tile * BM + r;
        float sc = 1.0f;
        if (dp) sc = __ldg(dp + (m < M ? m : M - 1) / rows_per_sample);
#pragma unroll 1
        for (int it = 0; it < Cfg::ITEMS; ++it) {
          const int ch = (ew >> 2) + (Cfg::NSLABW / 4) * it;
          float v[32];
          tmem_ld32(tmem_base + Cfg::ACC2_COL + lane_t + ch * 32, v);
          tmem_ld_wait();
          if (it == Cfg::ITEMS - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a2_empty);
          }
          const int n0 = ch * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + n0 + i));
            float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
            if (gamma) g4 = __ldg(reinterpret_cast<const float4*>(gamma + n0 + i));
            v[i] = sc * (g4.x * (v[i] + b4.x));
            v[i + 1] = sc * (g4.y * (v[i + 1] + b4.y));
            v[i + 2] = sc * (g4.z * (v[i + 2] + b4.z));
            v[i + 3] = sc * (g4.w * (v[i + 3] + b4.w));
          }
          mbar_wait(res_full + 8 * ew, rc & 1);
          ++rc;
#pragma unroll
          for (int jv = 0; jv < 8; ++jv) {
            uint32_t a0, a1, a2, a3;
            const uint32_t addr = slab + row_off + (((uint32_t)jv ^ rswz) << 4);
            lds128(addr, a0, a1, a2, a3);
            sts128(addr, __float_as_uint(v[4 * jv] + __uint_as_float(a0)), __float_as_uint(v[4 * jv + 1] + __uint_as_float(a1)),
                   __float_as_uint(v[4 * jv + 2] + __uint_as_float(a2)), __float_as_uint(v[4 * jv + 3] + __uint_as_float(a3)));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, slab, (int32_t)n0, y0);
            bulk_commit();
            if (it + 1 < Cfg::ITEMS) {
              bulk_wait_read0();
              mbar_expect_tx(res_full + 8 * ew, kSlabBytes);
              tma_load_2d(slab, &tmRes, res_full + 8 * ew, (int32_t)(((ew >> 2) + (Cfg::NSLABW / 4) * (it + 1)) * 32), y0);
            }
          }
          __syncwarp();
        }
      }
    }
    if (fin && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// Fused split-operand ("x3", fp32-accurate) MLP forward for C = 96, no-grad pass:
//     out = x + dp * gamma * ( GELU_erf(xn . W1^T + b1) . W2^T + b2 )      with every product as hi.hi + mid.hi + hi.mid
// The unfused pair writes the hidden activation as two bf16 pieces and reads it back: 2 x 8 bytes per hidden element, 6.1 KB of
// the 7.2 KB a token moves at C = 96 — both GEMMs sit on the HBM roof there.  Here the [128, 64] hidden chunk goes
// TMEM -> registers (bias, erf GELU in fp32, split into bf16 hi / mid) -> shared memory in the K-major 128B-swizzled layout of an
// MMA A operand -> GEMM2, and never reaches HBM: a token moves 2C*2 (xn pieces) + C*4 (x) + C*4 (out) = 1.2 KB.
// Structure as mlp_fused_fwd_kernel (producer warp, MMA issuer, 16 epilogue warps, acc1 double-buffered in TMEM, acc2
// persistent over the tile); differences: operands arrive as (hi, mid) box pairs, each GEMM issues three products, the hidden
// chunk has ONE shared-memory buffer pair (208 KB of shared memory are taken by the xn pieces, two weight stages and it), and the
// final epilogue reads the residual / writes the output rows straight from registers (a lane owns a 128-byte line of each).
// Operand tensors: xn2 [M, 2C] = [hi | mid] (cnx_dwconv7_ln_fwd_x3, segments = 2); W1x3 [4C, 3C] and W2x3 [C, 12C] = [hi | hi | mid]
// (cnx_weight_prep mode 3): hi at column 0, mid at column 2K.
// ================================================================================================
template <int C> struct Fu3Cfg {
  static constexpr int KBX = (C + 63) / 64;                 // 64-wide K boxes per piece of the xn / W1 tiles
  static constexpr int XN_BYTES = 2 * KBX * BM * 128;       // hi and mid pieces of the [128, C] xn tile
  static constexpr int W1_BYTES = 2 * KBX * kFuHCH * 128;   // hi and mid pieces of the [64 hidden rows, C] W1 chunk
  static constexpr int W2_BYTES = 2 * C * 128;              // hi and mid pieces of the [C rows, 64 hidden k] W2 chunk
  static constexpr int WST_BYTES = W1_BYTES + W2_BYTES;
  static constexpr int G_BYTES = 2 * BM * 128;              // hi and mid pieces of the [128, 64] GELU chunk
  static constexpr int NCH32 = C / 32;
  static constexpr int NFIN = NCH32 * 4;                    // warps of the final epilogue: one (lane quarter, 32-column chunk) each
  static constexpr int NBARS = 2 + 8 + 4 + 2 + 2 + 1;
  // final epilogue: one 4 KB transposition slab per warp; the first G_BYTES / 4096 of them reuse the hidden buffer pair (idle
  // between a tile's last GEMM2 and the next tile's first chunk), the rest get their own space
  static constexpr int XSLAB_BYTES = (NFIN * 4096 > G_BYTES) ? NFIN * 4096 - G_BYTES : 0;
  static constexpr int SMEM = XN_BYTES + 2 * WST_BYTES + G_BYTES + XSLAB_BYTES + 1024 + 8 * NBARS + 64;
  static constexpr int ACC2_COL = 128;
  static_assert(C % 32 == 0 && NFIN <= kFuEpiWarps && ACC2_COL + C <= 256, "fused x3 MLP tile shape");
  static_assert((C * 128) % 1024 == 0, "W2 pieces must stay 1024-byte aligned for the 128B swizzle");
  static_assert(SMEM <= 227 * 1024, "fused x3 MLP does not fit in shared memory");
};

template <int C>
__global__ void __launch_bounds__(kFuThreads, 1)
mlp_fused_x3_fwd_kernel(const __grid_constant__ CUtensorMap tmXn, const __grid_constant__ CUtensorMap tmW1,
                        const __grid_constant__ CUtensorMap tmW2, int64_t M, const float* __restrict__ b1,
                        const float* __restrict__ b2, const float* __restrict__ gamma, const float* __restrict__ dp,
                        int64_t rows_per_sample, const float* __restrict__ shortcut, float* __restrict__ out) {
  typedef Fu3Cfg<C> Cfg;
  constexpr int NJ = 4 * C / kFuHCH;                 // hidden chunks per tile
  constexpr int KBX = Cfg::KBX;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sXn = base;
  const uint32_t sW = sXn + Cfg::XN_BYTES;
  const uint32_t sG = sW + 2 * Cfg::WST_BYTES;
  const uint32_t bars = sG + Cfg::G_BYTES + Cfg::XSLAB_BYTES;     // (the extra slabs follow the hidden buffer pair: one slab array)
  int bi = 0;
  auto nb = [&](int n) { const uint32_t a = bars + 8u * bi; bi += n; return a; };
  const uint32_t xn_full = nb(1), xn_empty = nb(1);
  // the W1 and W2 pieces of a weight stage have their own barrier pairs: W1(j) is free again as soon as GEMM1(j) has retired,
  // a whole chunk before GEMM2(j) lets go of W2(j) — with one pair per stage every GEMM1 waited for a reload that could only
  // start when the GEMM2 of two chunks ago had retired (tensor pipe 22 % active, profiles/r02z6_stalls.txt)
  const uint32_t w1_full = nb(2), w1_empty = nb(2), w2_full = nb(2), w2_empty = nb(2);
  const uint32_t a1_full = nb(2), a1_empty = nb(2);
  const uint32_t g_full = nb(1), g_empty = nb(1);
  const uint32_t a2_full = nb(1), a2_empty = nb(1);
  const uint32_t tmem_slot = nb(1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (M + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmXn); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(xn_full, 1); mbar_init(xn_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(w1_full + 8 * i, 1); mbar_init(w1_empty + 8 * i, 1);
      mbar_init(w2_full + 8 * i, 1); mbar_init(w2_empty + 8 * i, 1);
      mbar_init(a1_full + 8 * i, 1); mbar_init(a1_empty + 8 * i, kFuEpiWarps);
    }
    mbar_init(g_full, kFuEpiWarps); mbar_init(g_empty, 1);
    mbar_init(a2_full, 1); mbar_init(a2_empty, Cfg::NFIN);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: the (hi, mid) xn tile once per tile; per hidden chunk the (hi, mid) W1 pieces — one chunk AHEAD of the
      // (hi, mid) W2 pieces, since GEMM1 runs a chunk ahead of GEMM2 =====
      const int64_t my_tiles = ((int64_t)blockIdx.x < m_tiles) ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const int64_t total = my_tiles * NJ;                         // hidden chunks this CTA walks
      auto load_xn = [&](int64_t t) {
        const int64_t tile = blockIdx.x + t * gridDim.x;
        mbar_wait(xn_empty, (uint32_t)((t & 1) ^ 1));
        mbar_expect_tx(xn_full, Cfg::XN_BYTES);
        for (int seg = 0; seg < 2; ++seg)
          for (int kb = 0; kb < KBX; ++kb)
            tma_load_2d(sXn + (seg * KBX + kb) * (BM * 128), &tmXn, xn_full, seg * C + kb * 64, (int32_t)(tile * BM));
        // the xn tile has ONE buffer: the next tile's load can only be issued when this tile's last GEMM1 has retired, so at
        // least make it an L2 hit
        if (tile + gridDim.x < m_tiles)
          for (int c = 0; c < 2 * C; c += 64) tma_prefetch_2d(&tmXn, c, (int32_t)((tile + gridDim.x) * BM));
      };
      auto load_w1 = [&](int64_t c) {
        const uint32_t ws = (uint32_t)(c & 1), wu = (uint32_t)(c >> 1);
        const int j = (int)(c % NJ);
        mbar_wait(w1_empty + 8 * ws, (wu & 1) ^ 1);
        mbar_expect_tx(w1_full + 8 * ws, Cfg::W1_BYTES);
        const uint32_t wb = sW + ws * Cfg::WST_BYTES;
        for (int seg = 0; seg < 2; ++seg)
          for (int kb = 0; kb < KBX; ++kb)
            tma_load_2d(wb + (seg * KBX + kb) * (kFuHCH * 128), &tmW1, w1_full + 8 * ws, seg * 2 * C + kb * 64, j * kFuHCH);
      };
      auto load_w2 = [&](int64_t c) {
        const uint32_t ws = (uint32_t)(c & 1), wu = (uint32_t)(c >> 1);
        const int j = (int)(c % NJ);
        mbar_wait(w2_empty + 8 * ws, (wu & 1) ^ 1);
        mbar_expect_tx(w2_full + 8 * ws, Cfg::W2_BYTES);
        const uint32_t wb = sW + ws * Cfg::WST_BYTES + Cfg::W1_BYTES;
        for (int seg = 0; seg < 2; ++seg) tma_load_2d(wb + seg * (C * 128), &tmW2, w2_full + 8 * ws, seg * 8 * C + j * kFuHCH, 0);
      };
      if (total > 0) {
        load_xn(0);
        load_w1(0);
      }
      for (int64_t c = 0; c < total; ++c) {
        if (c + 1 < total) {
          if ((c + 1) % NJ == 0) load_xn((c + 1) / NJ);            // first chunk of the next tile: its xn pieces first
          load_w1(c + 1);
        }
        load_w2(c);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: GEMM1(j) runs one chunk ahead of GEMM2(j-1); each is three products (hi.hi, mid.hi, hi.mid) =====
      constexpr uint32_t idesc1 = make_idesc(BM, kFuHCH, 0, 0);
      constexpr uint32_t idesc2 = make_idesc(BM, C, 0, 0);
      uint32_t tcnt = 0, cc = 0;
      for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++tcnt) {
        mbar_wait(xn_full, tcnt & 1);
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) {
            const uint32_t c = cc + j, b = c & 1, u = c >> 1;
            mbar_wait(w1_full + 8 * b, u & 1);
            mbar_wait(a1_empty + 8 * b, (u & 1) ^ 1);
            tc_fence_after();
            const uint32_t d1 = tmem_base + b * kFuHCH;
            uint32_t acc = 0;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              const int aseg = (p == 1) ? 1 : 0, bseg = (p == 2) ? 1 : 0;
#pragma unroll
              for (int kb = 0; kb < KBX; ++kb) {
                const uint64_t adesc = make_smem_desc(sXn + (aseg * KBX + kb) * (BM * 128), 16, 1024);
                const uint64_t bdesc = make_smem_desc(sW + b * Cfg::WST_BYTES + (bseg * KBX + kb) * (kFuHCH * 128), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (kb * 64 + k * 16 < C) { umma_f16(d1, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc1, acc); acc = 1; }
              }
            }
            umma_commit(a1_full + 8 * b);
            umma_commit(w1_empty + 8 * b);
            if (j == NJ - 1) umma_commit(xn_empty);
          }
          if (j > 0) {
            const uint32_t c = cc + j - 1, b = c & 1;
            mbar_wait(w2_full + 8 * b, (c >> 1) & 1);
            mbar_wait(g_full, c & 1);
            if (j == 1) mbar_wait(a2_empty, (tcnt & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              const int aseg = (p == 1) ? 1 : 0, bseg = (p == 2) ? 1 : 0;
              const uint64_t adesc = make_smem_desc(sG + aseg * (BM * 128), 16, 1024);
              const uint64_t bdesc = make_smem_desc(sW + b * Cfg::WST_BYTES + Cfg::W1_BYTES + bseg * (C * 128), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base + Cfg::ACC2_COL, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc2, ((j - 1) | p | k) != 0);
            }
            umma_commit(g_empty);
            umma_commit(w2_empty + 8 * b);
            if (j == NJ) umma_commit(a2_full);
          }
        }
        cc += NJ;
      }
    }
    __syncwarp();
  } else if (warp >= kFirstEpiWarp) {
    const int ew = warp - kFirstEpiWarp;
    const int quarter = warp & 3;
    const int cg = ew >> 2;                           // 16 of the chunk's 64 hidden columns; 32 of the output tile's columns
    const int r = quarter * 32 + lane;                // row within the tile
    const uint32_t lane_t = (uint32_t)(quarter * 32) << 16;
    const uint32_t g_row = sG + (uint32_t)r * 128u;
    const uint32_t swz = (uint32_t)(r & 7);
    const bool fin = ew < Cfg::NFIN;
    uint32_t cc = 0, tcnt = 0;
    for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++tcnt) {
      if (fin) {
        // the residual line of this lane's output row: pulled into L2 now, read after the tile's last chunk
        const int64_t m = tile * BM + r;
        if (m < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(shortcut + m * C + cg * 32));
      }
      for (int j = 0; j < NJ; ++j, ++cc) {
        const uint32_t b = cc & 1, u = cc >> 1;
        // the chunk's bias slice is fetched BEFORE the wait for its accumulator (it was 8 % of all warp samples as the first use
        // after the wait, profiles/r02z6_stalls.txt)
        float4 bq[4];
        const float* bp = b1 + j * kFuHCH + cg * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) bq[i] = __ldg(reinterpret_cast<const float4*>(bp + 4 * i));
        mbar_wait(a1_full + 8 * b, u & 1);
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem_base + b * kFuHCH + lane_t + cg * 16, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_empty + 8 * b);
        uint32_t ph[8], pm[8];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = bq[i / 4];
          // fp32 h (no rounding), erf GELU in fp32, g ~ hi + mid to 2^-17 relative
          const float2 ga = gelu_as2(__fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y)));
          const float2 gb = gelu_as2(__fadd2_rn(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w)));
          const uint32_t ha = pack_bf16(ga.x, ga.y), hb = pack_bf16(gb.x, gb.y);
          ph[i / 2] = ha;
          ph[i / 2 + 1] = hb;
          const float2 ra = __fadd2_rn(ga, make_float2(-bf16_lo(ha), -bf16_hi(ha)));
          const float2 rb = __fadd2_rn(gb, make_float2(-bf16_lo(hb), -bf16_hi(hb)));
          pm[i / 2] = pack_bf16(ra.x, ra.y);
          pm[i / 2 + 1] = pack_bf16(rb.x, rb.y);
        }
        mbar_wait(g_empty, (cc & 1) ^ 1);             // GEMM2 of the previous chunk has consumed the buffer pair
        const uint32_t o0 = ((uint32_t)(cg * 2) ^ swz) << 4, o1 = ((uint32_t)(cg * 2 + 1) ^ swz) << 4;
        sts128(g_row + o0, ph[0], ph[1], ph[2], ph[3]);
        sts128(g_row + o1, ph[4], ph[5], ph[6], ph[7]);
        sts128(g_row + BM * 128 + o0, pm[0], pm[1], pm[2], pm[3]);
        sts128(g_row + BM * 128 + o1, pm[4], pm[5], pm[6], pm[7]);
        fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(g_full);
      }
      if (fin) {
        // ===== final epilogue: this warp's (lane quarter, 32-column chunk) of the [128, C] output tile =====
        mbar_wait(a2_full, tcnt & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + Cfg::ACC2_COL + lane_t + cg * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_empty);
        // A lane holds 32 columns of ONE row.  Reading the residual and writing the output from that layout touches 32
        // different 128-byte lines per instruction (measured: 38 % of the kernel's time, profiles/r02z13_*).  The values are
        // transposed through a 4 KB shared-memory slab instead: written row-per-lane (swizzled, conflict-free), read back so
        // that one instruction covers 4 rows x 128 contiguous bytes, and the residual load / output store use that layout.
        const int64_t m = tile * BM + r;
        const float sc = dp ? __ldg(dp + (m < M ? m : M - 1) / rows_per_sample) : 1.0f;
        const int n0 = cg * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + n0 + i));
          float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (gamma) g4 = __ldg(reinterpret_cast<const float4*>(gamma + n0 + i));
          v[i] = sc * (g4.x * (v[i] + b4.x));
          v[i + 1] = sc * (g4.y * (v[i + 1] + b4.y));
          v[i + 2] = sc * (g4.z * (v[i + 2] + b4.z));
          v[i + 3] = sc * (g4.w * (v[i + 3] + b4.w));
        }
        const uint32_t slab = sG + (uint32_t)ew * 4096u;
        const uint32_t rsw = (uint32_t)(lane & 7);
#pragma unroll
        for (int jv = 0; jv < 8; ++jv)
          sts128(slab + (uint32_t)lane * 128u + (((uint32_t)jv ^ rsw) << 4), __float_as_uint(v[4 * jv]), __float_as_uint(v[4 * jv + 1]),
                 __float_as_uint(v[4 * jv + 2]), __float_as_uint(v[4 * jv + 3]));
        __syncwarp();
        const int64_t row0 = tile * BM + quarter * 32;
        const int c16 = lane & 7;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int rr = t * 4 + (lane >> 3);
          uint32_t a0, a1, a2, a3;
          lds128(slab + (uint32_t)rr * 128u + (((uint32_t)c16 ^ (uint32_t)(rr & 7)) << 4), a0, a1, a2, a3);
          const int64_t mm = row0 + rr;
          if (mm < M) {
            const float4 x4 = __ldg(reinterpret_cast<const float4*>(shortcut + mm * C + n0 + c16 * 4));
            float4 o;
            o.x = x4.x + __uint_as_float(a0);
            o.y = x4.y + __uint_as_float(a1);
            o.z = x4.z + __uint_as_float(a2);
            o.w = x4.w + __uint_as_float(a3);
            *reinterpret_cast<float4*>(out + mm * C + n0 + c16 * 4) = o;
          }
        }
      }
      // no warp may write the next tile's first hidden chunk while a slab in the hidden buffer pair is still being read
      named_bar(1, kFuEpiWarps * 32);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ================================================================================================
// wgrad on CTA pairs: one tcgen05.mma.cta_group::2 spans a 256 x BN tile of out = X^T.Y (BN = 128, 256 or 384).  Each CTA stages
// its own 128 channels of X (two 64-channel MN-major boxes) and HALF of the BN channels of Y per 64-row k-block, so a CTA moves
// 24 / 32 KB per k-block where the single-CTA 128 x 128 tile moves 32 KB for a quarter / half of the MACs.  Bias gradient,
// split-K partials and the deterministic reduction are as in gemm_wgrad_tc_kernel.  grid = (2 * tiles, splits), cluster (2,1,1).
// BN = 384 (N2 a multiple of 384: the 4C x C and C x 4C weights of C = 96 k): two MMAs per k-step, N = 256 into TMEM columns
// 0..255 and N = 128 into 256..383 — 40 KB staged per CTA and k-block for 1.5x the MACs of the 256 x 256 tile (these long-K
// GEMMs are bound by the bytes an SM takes in and reads back per clock), and N2 = 384 is ONE tile instead of
// two 256-wide ones of which the second is half zero fill.
// ================================================================================================
template <int BN> struct WgPairCfg {
  static constexpr int A_BYTES = 2 * 64 * BK * 2;                      // own 128 channels of X
  static constexpr int B_BOXES = BN / 2 / 64;                          // 64-channel boxes of Y staged by one CTA
  static constexpr int B_BYTES = B_BOXES * 64 * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN <= 256) ? 6 : 5;
  static constexpr int CS_COL = BN;                                    // 16 column-sum columns after the accumulator
  static constexpr int TMEM_COLS = (BN + 16 <= 256) ? 256 : 512;
  static constexpr int EPI_COLS = BN / 2;
  static constexpr int ONES_BYTES = 2048;
  static constexpr int SMEM = STAGES * STAGE_BYTES + ONES_BYTES + 1024 + 256;
  static_assert(BN == 128 || BN == 256 || BN == 384, "the pair's half of the Y tile must be whole 64-channel boxes");
  static_assert(SMEM <= 227 * 1024, "wgrad pair ring does not fit in shared memory");
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, int64_t Mtot,
                       int64_t N1, int64_t N2, int kb_per_split, float* __restrict__ part, float* __restrict__ cs_part, int passes) {
  typedef WgPairCfg<BN> Cfg;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * Cfg::A_BYTES;
  const uint32_t sOnes = sB + STAGES * Cfg::B_BYTES;
  const uint32_t bars = sOnes + Cfg::ONES_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t j_tiles = (N2 + BN - 1) / BN;
  const int64_t tile = blockIdx.x >> 1;
  const int64_t it = tile / j_tiles, jt = tile - it * j_tiles;
  const int split = blockIdx.y;
  const int kb_rows = (int)((Mtot + BK - 1) / BK);     // passes = 3: one K loop over hi.hi, mid.hi, hi.mid (see gemm_wgrad_tc_kernel)
  const int kb_total = passes * kb_rows;
  const int kb0 = split * kb_per_split;
  int kb1 = kb0 + kb_per_split;
  if (kb1 > kb_total) kb1 = kb_total;
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;
  const bool do_colsum = (cs_part != nullptr) && (jt == 0);

  {
    uint8_t* gen = smem_raw + (sOnes - smem_u32(smem_raw));
    for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += kThreads) reinterpret_cast<uint32_t*>(gen)[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const int32_t a_ch = (int32_t)(it * 256 + rank * 128);
      // BN <= 256: this CTA's half of the tile's channels.  BN = 384: the halves of the N = 256 MMA's channels (boxes 0, 1)
      // and of the N = 128 MMA's (box 2), so that accumulator columns stay in channel order
      const int32_t b_ch = (int32_t)(jt * BN + rank * (BN == 384 ? 128 : BN / 2));
      const int32_t b_ch2 = (int32_t)(jt * BN + 256 + rank * 64);
      for (int kb = 0; kb < nkb; ++kb) {
        const int g = kb0 + kb, pass = g / kb_rows;
        const int32_t mrow = (g - pass * kb_rows) * BK;
        const int32_t xo = (pass == 1) ? (int32_t)N1 : 0, yo = (pass == 2) ? (int32_t)N2 : 0;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
        tma_load_2d_pair(sA + s * Cfg::A_BYTES, &tmX, full_bar(s), a_ch + xo, mrow);
        tma_load_2d_pair(sA + s * Cfg::A_BYTES + 8192, &tmX, full_bar(s), a_ch + 64 + xo, mrow);
#pragma unroll
        for (int b = 0; b < Cfg::B_BOXES; ++b)
          tma_load_2d_pair(sB + s * Cfg::B_BYTES + b * 8192, &tmY, full_bar(s), ((BN == 384 && b == 2) ? b_ch2 : b_ch + 64 * b) + yo, mrow);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0 && nkb > 0) {
      constexpr uint32_t idesc = make_idesc(256, BN == 384 ? 256 : BN, 1, 1);
      constexpr uint32_t idesc_hi = make_idesc(256, 128, 1, 1);         // BN = 384: columns 256..383
      constexpr uint32_t idesc_ones = make_idesc(256, 16, 1, 1);
      const uint64_t odesc = make_smem_desc(sOnes, 8192, 1024);
      int s = 0; uint32_t ph = 0;
      uint32_t cs_acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(sA + s * Cfg::A_BYTES, 8192, 1024);
        const uint64_t bdesc = make_smem_desc(sB + s * Cfg::B_BYTES, 8192, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_f16_pair(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (kb | k) != 0);
          if constexpr (BN == 384) {
            const uint64_t bdesc2 = make_smem_desc(sB + s * Cfg::B_BYTES + 2 * 8192, 8192, 1024);
            umma_f16_pair(tmem_base + 256, adesc + (uint64_t)(k * 128), bdesc2 + (uint64_t)(k * 128), idesc_hi, (kb | k) != 0);
          }
          if (do_colsum && (kb0 + kb) < 2 * kb_rows) {
            umma_f16_pair(tmem_base + Cfg::CS_COL, adesc + (uint64_t)(k * 128), odesc, idesc_ones, cs_acc);
            cs_acc = 1;
          }
        }
        umma_commit_pair(empty_bar(s));
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit_pair(tfull_bar);
    }
    __syncwarp();
  } else if (warp >= kFirstEpiWarp) {
    const int quarter = warp & 3;
    const int half = (warp - kFirstEpiWarp) >> 2;
    const int64_t i = it * 256 + rank * 128 + quarter * 32 + lane;
    float* po = part + (int64_t)split * N1 * N2;
    if (nkb > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // A lane holds 32 columns of ONE accumulator row: stored from that layout every instruction touches 32 different lines of the
    // partial buffer.  Each 32 x 32 chunk is transposed through a 4 KB slab of the (now idle) operand ring instead, so that one
    // store instruction covers 4 rows x 128 contiguous bytes.
    const uint32_t slab = sA + (uint32_t)(warp - kFirstEpiWarp) * 4096u;
    static_assert(STAGES * Cfg::A_BYTES >= kEpiWarps * 4096, "the A ring must hold one slab per epilogue warp");
    const int64_t row0 = it * 256 + rank * 128 + quarter * 32;
    const uint32_t rsw = (uint32_t)(lane & 7);
    const int c16 = lane & 7;
#pragma unroll 1
    for (int c = 0; c < Cfg::EPI_COLS; c += 32) {
      const int col = half * Cfg::EPI_COLS + c;
      float v[32];
      if (nkb > 0) {
        tmem_ld32(t_row + col, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int jv = 0; jv < 8; ++jv)
        sts128(slab + (uint32_t)lane * 128u + (((uint32_t)jv ^ rsw) << 4), __float_as_uint(v[4 * jv]), __float_as_uint(v[4 * jv + 1]),
               __float_as_uint(v[4 * jv + 2]), __float_as_uint(v[4 * jv + 3]));
      __syncwarp();
      const int64_t jj = jt * BN + col + c16 * 4;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int rr = t * 4 + (lane >> 3);
        uint32_t a0, a1, a2, a3;
        lds128(slab + (uint32_t)rr * 128u + (((uint32_t)c16 ^ (uint32_t)(rr & 7)) << 4), a0, a1, a2, a3);
        if (row0 + rr < N1 && jj < N2)
          *reinterpret_cast<uint4*>(po + (row0 + rr) * N2 + jj) = make_uint4(a0, a1, a2, a3);
      }
      __syncwarp();
    }
    if (do_colsum && half == 0) {
      float v[16];
      if (nkb > 0 && kb0 < 2 * kb_rows) {              // this split issued at least one column-sum MMA
        tmem_ld16(t_row + Cfg::CS_COL, v);
        tmem_ld_wait();
      } else {
        v[0] = 0.f;
      }
      if (i < N1) cs_part[(int64_t)split * N1 + i] = v[0];
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- host side -----------------------------------------------------------------------------------
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

// 2-D bf16 row-major tensor [rows, cols]; box = [box_rows][64 cols], 128-byte swizzle, OOB -> zeros
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows, int64_t ld = 0) {
  EncodeTiledFn enc = get_encode();
  CNX_REQUIRE(enc != nullptr, CNX_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  CNX_REQUIRE((((uintptr_t)ptr) & 15) == 0, CNX_E_SHAPE, "GEMM operand must be 16-byte aligned");
  CNX_REQUIRE(cols % 8 == 0, CNX_E_SHAPE, "GEMM operand inner dimension %lld must be a multiple of 8", (long long)cols);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld > 0 ? ld : cols) * 2};     // ld: row stride in elements when the operand is a column block
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CNX_REQUIRE(r == CUDA_SUCCESS, CNX_E_DRIVER, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d",
              (int)r, (long long)rows, (long long)cols, box_rows);
  return 0;
}

// 2-D row-major tensor [rows, cols] of elem_bytes elements; box = [box_rows][32 cols] for the epilogue slabs
static int make_slab_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int elem_bytes) {
  EncodeTiledFn enc = get_encode();
  CNX_REQUIRE(enc != nullptr, CNX_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  CNX_REQUIRE((((uintptr_t)ptr) & 15) == 0, CNX_E_SHAPE, "GEMM output must be 16-byte aligned");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   elem_bytes == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CNX_REQUIRE(r == CUDA_SUCCESS, CNX_E_DRIVER, "cuTensorMapEncodeTiled (slab) failed (%d) rows=%lld cols=%lld", (int)r,
              (long long)rows, (long long)cols);
  return 0;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%d): %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_GEMM_NCTA");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

// CNX_GEMM_PF=<n>: L2 prefetch distance of the persistent GEMM's producer, in tiles of its own loop (0 = off).  Measured
// (profiles/r02x_kbench_gemm_pf*.jsonl): -3..-8 % time for the split-output fc1 (EPI_BIAS_GELU3) at distance 1, +5..+45 % for
// every kernel whose epilogue loads an input slab and for the plain data-gradient GEMM — so it is on for the former only.
static int l2_prefetch_tiles(int kind) {
  static int v = -2;
  if (v == -2) { const char* e = getenv("CNX_GEMM_PF"); v = e ? atoi(e) : -1; }
  if (v >= 0) return v;
  return kind == EPI_BIAS_GELU3 ? 1 : 0;
}
static bool bn384_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_GEMM_BN384");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
static bool slab_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_GEMM_SLAB");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int BN, int KIND, typename TOUT, int NCTA, bool SLAB, int NEPI = kEpiWarps>
static int launch_tn_impl(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s) {
  typedef TnCfg<BN, NCTA, SLAB, NEPI, KIND == EPI_DGELU_RC> Cfg;
  CUtensorMap tmA, tmB, tmOut, tmIn, tmB2;
  if (int rc = make_map(&tmA, A, M, ep.a_wrap > 0 ? (int64_t)ep.a_wrap : K, BM)) return rc;
  if (int rc = make_map(&tmB, B, N, K, BN == 384 ? 64 : Cfg::B_ROWS)) return rc;
  if (SLAB) {
    constexpr bool kGelu = (KIND == EPI_BIAS_GELU || KIND == EPI_BIAS_GELU3);
    void* o = kGelu ? ep.out1 : ep.out0;
    const void* in = kGelu ? (ep.out0 ? ep.out0 : ep.out1) : (ep.aux ? ep.aux : ep.out0);
    if (int rc = make_slab_map(&tmOut, o, M, (KIND == EPI_BIAS_GELU3 || KIND == EPI_DGELU3) ? 2 * N : N, (int)sizeof(TOUT))) return rc;
    if (KIND == EPI_DGELU3) {
      if (int rc = make_slab_map(&tmIn, ep.aux, M, N, 4)) return rc;            // fp32 GELU'(h)
    } else if (KIND == EPI_BIAS_GELU3) {
      tmIn = tmOut;
      if (ep.out0 != nullptr)
        if (int rc = make_slab_map(&tmIn, ep.out0, M, N, 4)) return rc;         // fp32 GELU'(h), training
    } else if (KIND == EPI_DGELU_RC) {
      if (int rc = make_map(&tmIn, ep.a2, M, K, BM)) return rc;               // A2 = xn
    } else if (int rc = make_slab_map(&tmIn, in, M, N, (int)sizeof(TOUT))) return rc;
  } else {
    tmOut = tmA;
    tmIn = tmA;
  }
  tmB2 = tmB;
  if (KIND == EPI_DGELU_RC)
    if (int rc = make_map(&tmB2, ep.b2, N, K, Cfg::B_ROWS)) return rc;        // B2 = W1
  auto k = gemm_tn_tc_kernel<BN, KIND, TOUT, NCTA, SLAB, NEPI>;
  if (int rc = set_smem(k, Cfg::SMEM)) return rc;
  EpiParams epk = ep;
  epk.pf_tiles = l2_prefetch_tiles(KIND);
  const int64_t tiles = ((M + BM * NCTA - 1) / (BM * NCTA)) * ((N + BN - 1) / BN);
  int64_t grid = sm_count() / NCTA;
  {
    // experiments (profiles/kbench.py): CNX_GEMM_MAXSMS caps the number of SMs the persistent grid occupies — if throughput
    // per launch barely drops with fewer SMs, the kernel is bound by what the SMs share (L2 -> SM bandwidth), not by the SMs
    static int cap = -1;
    if (cap < 0) { const char* e = getenv("CNX_GEMM_MAXSMS"); cap = e ? atoi(e) : 0; }
    if (cap > 0 && grid > cap / NCTA) grid = cap / NCTA;
  }
  if (grid > tiles) grid = tiles;
  grid *= NCTA;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, tmA, tmB, tmOut, tmIn, tmB2, M, N, K, epk);
  if (e != cudaSuccess) {
    set_error("gemm_tn_tc launch: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return (int)e;
  }
  return check_launch("gemm_tn_tc");
}

template <int BN, int KIND, typename TOUT, int NCTA>
static int launch_tn(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s) {
  // slab epilogue: whole 32-column chunks, an input slab only where the epilogue has one
  constexpr bool kGelu = (KIND == EPI_BIAS_GELU || KIND == EPI_BIAS_GELU3);
  constexpr bool kSlabKind = (BN % 32 == 0) && (!kGelu || sizeof(TOUT) == 2);
  if constexpr (kSlabKind) {
    const bool need_in = (KIND == EPI_SCALE_RES || KIND == EPI_DGELU || KIND == EPI_DGELU3);
    const void* o = kGelu ? ep.out1 : ep.out0;
    if ((slab_enabled() || KIND == EPI_BIAS_GELU3 || KIND == EPI_DGELU3) && N % 32 == 0 && (!need_in || ep.aux != nullptr) && o != nullptr) {
      // the GELU epilogue is issue-bound: 16 epilogue warps (4 per scheduler) where the tile has >= 4 column chunks
      if constexpr (kGelu && BN >= 128) {
        // 16 warps' slabs (128 KB) leave a 3-stage operand ring.  The split-output fc1 with K' >= 576 (C >= 192) is tensor-
        // bound, not epilogue-bound: 8 warps and a 5-stage ring are 2-4 % faster there (profiles/r02z1_kbench_nepi8_*.jsonl);
        // CNX_GEMM_NEPI8=0/1 forces 16 / 8 warps for every GELU-epilogue GEMM
        static int nepi8 = -2;
        if (nepi8 == -2) { const char* e = getenv("CNX_GEMM_NEPI8"); nepi8 = e ? (e[0] == '1' ? 1 : 0) : -1; }
        const bool eight = nepi8 >= 0 ? nepi8 == 1 : (KIND == EPI_BIAS_GELU3 && K >= 576);
        if (eight) return launch_tn_impl<BN, KIND, TOUT, NCTA, true>(A, B, M, N, K, ep, s);
        return launch_tn_impl<BN, KIND, TOUT, NCTA, true, 16>(A, B, M, N, K, ep, s);
      } else return launch_tn_impl<BN, KIND, TOUT, NCTA, true>(A, B, M, N, K, ep, s);
    }
  }
  if constexpr (KIND == EPI_BIAS_GELU3 || KIND == EPI_DGELU3) {
    set_error("split-output GEMM (x3): N=%lld must be a multiple of 32 and the output / GELU' pointers non-null", (long long)N);
    return CNX_E_SHAPE;
  } else {
    return launch_tn_impl<BN, KIND, TOUT, NCTA, false>(A, B, M, N, K, ep, s);
  }
}


// 2-D bf16 row-major tensor [rows, cols]; box = [box_rows][64 cols], 128-byte swizzle (make_map), used by the fused MLP
template <int C>
static int launch_mlp_fused(const void* xn, const void* W1, const float* b1, const void* W2, const float* b2, const float* gamma,
                            const float* dp, int64_t rows_per_sample, const void* shortcut, void* out, int64_t M, cudaStream_t s) {
  typedef FuCfg<C> Cfg;
  CUtensorMap tmXn, tmW1, tmW2, tmRes, tmOut;
  if (int rc = make_map(&tmXn, xn, M, C, BM)) return rc;
  if (int rc = make_map(&tmW1, W1, 4 * C, C, kFuHCH)) return rc;
  if (int rc = make_map(&tmW2, W2, C, 4 * C, C)) return rc;
  if (int rc = make_slab_map(&tmRes, shortcut, M, C, 4)) return rc;
  if (int rc = make_slab_map(&tmOut, out, M, C, 4)) return rc;
  auto k = mlp_fused_fwd_kernel<C>;
  if (int rc = set_smem(k, Cfg::SMEM)) return rc;
  int64_t grid = sm_count();
  const int64_t tiles = (M + BM - 1) / BM;
  if (grid > tiles) grid = tiles;
  launch_pdl(k, dim3((unsigned)grid), dim3(kFuThreads), Cfg::SMEM, s, tmXn, tmW1, tmW2, tmRes, tmOut, M, b1, b2, gamma, dp, rows_per_sample);
  return check_launch("mlp_fused_fwd");
}

template <int C>
static int launch_mlp_fused_x3(const void* xn2, const void* W1x3, const float* b1, const void* W2x3, const float* b2, const float* gamma,
                               const float* dp, int64_t rows_per_sample, const float* shortcut, float* out, int64_t M, cudaStream_t s) {
  typedef Fu3Cfg<C> Cfg;
  CUtensorMap tmXn, tmW1, tmW2;
  if (int rc = make_map(&tmXn, xn2, M, 2 * C, BM)) return rc;                 // [hi | mid]
  if (int rc = make_map(&tmW1, W1x3, 4 * C, 3 * C, kFuHCH)) return rc;        // [hi | hi | mid]: pieces at columns 0 and 2C
  if (int rc = make_map(&tmW2, W2x3, C, 12 * C, C)) return rc;                // [hi | hi | mid]: pieces at columns 0 and 8C
  auto k = mlp_fused_x3_fwd_kernel<C>;
  if (int rc = set_smem(k, Cfg::SMEM)) return rc;
  int64_t grid = sm_count();
  const int64_t tiles = (M + BM - 1) / BM;
  if (grid > tiles) grid = tiles;
  launch_pdl(k, dim3((unsigned)grid), dim3(kFuThreads), Cfg::SMEM, s, tmXn, tmW1, tmW2, M, b1, b2, gamma, dp, rows_per_sample, shortcut, out);
  return check_launch("mlp_fused_fwd_x3");
}

}  // namespace tc

int mlp_fused_fwd_tc(const void* xn, const void* W1, const float* b1, const void* W2, const float* b2, const float* gamma,
                     const float* dp, int64_t rows_per_sample, const void* shortcut, void* out, int64_t M, int64_t C,
                     cudaStream_t s) {
  if (C == 96) return tc::launch_mlp_fused<96>(xn, W1, b1, W2, b2, gamma, dp, rows_per_sample, shortcut, out, M, s);
  if (C == 128) return tc::launch_mlp_fused<128>(xn, W1, b1, W2, b2, gamma, dp, rows_per_sample, shortcut, out, M, s);
  if (C == 192) return tc::launch_mlp_fused<192>(xn, W1, b1, W2, b2, gamma, dp, rows_per_sample, shortcut, out, M, s);
  set_error("mlp_fused_fwd: C=%lld is not a fused shape (96, 128, 192)", (long long)C);
  return CNX_E_SHAPE;
}


int mlp_fused_fwd_x3_tc(const void* xn2, const void* W1x3, const float* b1, const void* W2x3, const float* b2, const float* gamma,
                        const float* dp, int64_t rows_per_sample, const float* shortcut, float* out, int64_t M, int64_t C,
                        cudaStream_t s) {
  if (C == 96) return tc::launch_mlp_fused_x3<96>(xn2, W1x3, b1, W2x3, b2, gamma, dp, rows_per_sample, shortcut, out, M, s);
  set_error("mlp_fused_fwd_x3: C=%lld is not a fused shape (96)", (long long)C);
  return CNX_E_SHAPE;
}

template <int KIND, typename TOUT>
int gemm_tn_tc(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s) {
  CNX_REQUIRE(N % 8 == 0 && K % 8 == 0, CNX_E_SHAPE, "gemm_tc: N=%lld and K=%lld must be multiples of 8", (long long)N,
              (long long)K);
  // CTA-pair tiles (256 x BN) wherever a whole tile fits; CNX_GEMM_NCTA=1 in the environment selects the single-CTA kernel
  if (tc::pair_enabled() && M >= 256) {
    if constexpr (KIND == EPI_PLAIN || KIND == EPI_SCALE_RES) {
      // long-K GEMMs onto N = 384 * j columns (fc2 and the fc1 data gradient at C = 384 / 768, and their split-operand
      // forms): 256 x 384 single-accumulator tiles when they fill the machine's waves about as well as the alternative
      if (N % 384 == 0 && K >= 1024 && tc::bn384_enabled()) {
        const int64_t pairs = sm_count() / 2, mt = (M + 255) / 256;
        auto eff = [&](int64_t tiles) { return (double)tiles / (double)(((tiles + pairs - 1) / pairs) * pairs); };
        const int alt = (N % 256 == 0) ? 256 : 192;
        if (eff(mt * (N / 384)) + 0.05 >= eff(mt * (N / alt))) {
          const bool need_in = (KIND == EPI_SCALE_RES);
          if (N % 32 == 0 && (!need_in || ep.aux != nullptr) && ep.out0 != nullptr && tc::slab_enabled())
            return tc::launch_tn_impl<384, KIND, TOUT, 2, true>(A, B, M, N, K, ep, s);
        }
      }
    }
    if (N % 256 == 0) return tc::launch_tn<256, KIND, TOUT, 2>(A, B, M, N, K, ep, s);
    if (N % 192 == 0) return tc::launch_tn<192, KIND, TOUT, 2>(A, B, M, N, K, ep, s);
  }
  if (N % 128 == 0) return tc::launch_tn<128, KIND, TOUT, 1>(A, B, M, N, K, ep, s);
  if (N % 96 == 0) return tc::launch_tn<96, KIND, TOUT, 1>(A, B, M, N, K, ep, s);
  if (N % 64 == 0) return tc::launch_tn<64, KIND, TOUT, 1>(A, B, M, N, K, ep, s);
  return tc::launch_tn<128, KIND, TOUT, 1>(A, B, M, N, K, ep, s);   // ragged N: TMA zero-fills, epilogue guards
}

// dh = (dz . Bt^T) * GELU'(round(xn . W1^T + b1)): CTA-pair 256 x 128 tiles, two accumulators, 16 epilogue warps
int gemm_dgelu_recompute_tc(const void* dz, const void* Bt, int64_t M, int64_t N, int64_t K, const EpiParams& ep, cudaStream_t s) {
  CNX_REQUIRE(N % 128 == 0 && K % 8 == 0 && M >= 256, CNX_E_SHAPE,
              "gemm_dgrad_gelu_recompute: needs N %% 128 == 0, K %% 8 == 0, M >= 256 (N=%lld K=%lld M=%lld)", (long long)N,
              (long long)K, (long long)M);
  return tc::launch_tn_impl<128, EPI_DGELU_RC, bf16, 2, true, 16>(dz, Bt, M, N, K, ep, s);
}

#define CNX_INST(KIND, TOUT) \
  template int gemm_tn_tc<KIND, TOUT>(const void*, const void*, int64_t, int64_t, int64_t, const EpiParams&, cudaStream_t);
CNX_INST(EPI_PLAIN, bf16)
CNX_INST(EPI_PLAIN, float)
CNX_INST(EPI_BIAS_GELU, bf16)
CNX_INST(EPI_SCALE_RES, float)
CNX_INST(EPI_SCALE_RES, bf16)
CNX_INST(EPI_DGELU, bf16)
CNX_INST(EPI_BIAS_GELU3, bf16)
CNX_INST(EPI_DGELU3, bf16)
#undef CNX_INST

static bool wgrad384_enabled() {               // CNX_WGRAD_BN384=0: 256 x 256 pair tiles only (A/B measurements)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CNX_WGRAD_BN384");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// tile plan of the wgrad GEMM: CTA pairs (256 x BN) where both dimensions carry at least a tile, single CTAs otherwise
struct WgPlan { int pair; int bn; int splits; int64_t tiles; };
static WgPlan wgrad_plan_tc(int64_t M, int64_t N1, int64_t N2) {
  WgPlan p;
  p.pair = (tc::pair_enabled() && N1 >= 192 && N2 >= 128) ? 1 : 0;
  if (p.pair) {
    // 256-wide tiles even when the last one is ragged (TMA zero fill): measured faster than exact 128-wide tiles (384 -> 2 x 256)
    p.bn = (N2 >= 192) ? 256 : 128;
    // 256 x 384 only where 256-wide tiles would be ragged (N2 = 384: one tile instead of 1.5): measured +16 % there and
    // -2..-5 % where N2 is a multiple of 256 as well (profiles/r02v_kbench_wgrad_bn384_*.jsonl)
    if (N2 % 384 == 0 && N2 % 256 != 0 && wgrad384_enabled()) p.bn = 384;
    p.tiles = ((N1 + 255) / 256) * ((N2 + p.bn - 1) / p.bn);
  } else {
    p.bn = (N2 % 128 == 0) ? 128 : ((N2 % 96 == 0) ? 96 : 128);
    p.tiles = ((N1 + tc::BM - 1) / tc::BM) * ((N2 + p.bn - 1) / p.bn);
  }
  const int64_t slots = p.pair ? sm_count() / 2 : sm_count();   // one wave: tiles * splits <= CTAs (pairs) that fit
  int64_t want = slots / p.tiles;
  const int64_t kb_total = (M + tc::BK - 1) / tc::BK;
  const int64_t maxs = (kb_total + 7) / 8;                      // at least 8 k-blocks (512 rows) per split
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  p.splits = (int)want;
  return p;
}

int64_t wgrad_workspace_bytes_tc(int64_t M, int64_t N1, int64_t N2) {
  return (int64_t)wgrad_plan_tc(M, N1, N2).splits * (N1 * N2 + N1) * 4;
}

template <int BN>
static int launch_wgrad_pair(const CUtensorMap& tmX, const CUtensorMap& tmY, int64_t M, int64_t N1, int64_t N2, int kb_per_split,
                             int64_t tiles, int splits, float* part, float* cs_part, cudaStream_t s, int passes) {
  auto k = tc::gemm_wgrad_pair_kernel<BN>;
  if (int rc = tc::set_smem(k, tc::WgPairCfg<BN>::SMEM)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * tiles), (unsigned)splits);
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = tc::WgPairCfg<BN>::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, tmX, tmY, M, N1, N2, kb_per_split, part, cs_part, passes);
  if (e != cudaSuccess) {
    set_error("gemm_wgrad_pair launch: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return (int)e;
  }
  return check_launch("gemm_wgrad_pair");
}

int gemm_wgrad_tc(const void* X, const void* Y, int64_t M, int64_t N1, int64_t N2, int accumulate, float* out,
                  float* colsum_x, void* workspace, int64_t workspace_bytes, cudaStream_t s, int64_t ldx, int64_t ldy, int passes) {
  CNX_REQUIRE(N1 % 8 == 0 && N2 % 8 == 0, CNX_E_SHAPE, "gemm_wgrad_tc: N1, N2 must be multiples of 8");
  CNX_REQUIRE((ldx == 0 || (ldx >= N1 && ldx % 8 == 0)) && (ldy == 0 || (ldy >= N2 && ldy % 8 == 0)), CNX_E_SHAPE,
              "gemm_wgrad_tc: row strides must cover the rows and be multiples of 8");
  const WgPlan pl = wgrad_plan_tc(M, N1, N2);
  const int splits = pl.splits;
  CNX_REQUIRE(workspace_bytes >= (int64_t)splits * (N1 * N2 + N1) * 4, CNX_E_WORKSPACE, "gemm_wgrad: workspace too small");
  CUtensorMap tmX, tmY;
  // passes = 3: X and Y are the split operands [hi | mid] (2 N columns, row stride 2 N); the kernels walk hi.hi, mid.hi, hi.mid
  // in one K loop.  (A ragged last tile then reads the neighbouring piece instead of TMA zero fill: those rows / columns of
  // the accumulator are never stored.)
  CNX_REQUIRE(passes == 1 || (passes == 3 && ldx == 2 * N1 && ldy == 2 * N2), CNX_E_BADARG,
              "gemm_wgrad_tc: passes must be 1, or 3 with split operands of row stride 2 N1 / 2 N2");
  if (int rc = tc::make_map(&tmX, X, M, passes == 3 ? 2 * N1 : N1, tc::BK, ldx)) return rc;
  if (int rc = tc::make_map(&tmY, Y, M, passes == 3 ? 2 * N2 : N2, tc::BK, ldy)) return rc;
  const int64_t kb_total = passes * ((M + tc::BK - 1) / tc::BK);
  const int kb_per_split = (int)((kb_total + splits - 1) / splits);
  float* part = (float*)workspace;
  float* cs_part = part + (int64_t)splits * N1 * N2;
  if (pl.pair) {
    int rc = (pl.bn == 384)
                 ? launch_wgrad_pair<384>(tmX, tmY, M, N1, N2, kb_per_split, pl.tiles, splits, part, colsum_x ? cs_part : nullptr, s, passes)
             : (pl.bn == 256)
                 ? launch_wgrad_pair<256>(tmX, tmY, M, N1, N2, kb_per_split, pl.tiles, splits, part, colsum_x ? cs_part : nullptr, s, passes)
                 : launch_wgrad_pair<128>(tmX, tmY, M, N1, N2, kb_per_split, pl.tiles, splits, part, colsum_x ? cs_part : nullptr, s, passes);
    if (rc) return rc;
  } else {
    dim3 grid((unsigned)pl.tiles, (unsigned)splits);
    if (pl.bn == 128) {
      auto k = tc::gemm_wgrad_tc_kernel<128>;
      if (int rc = tc::set_smem(k, tc::WgCfg<128>::SMEM)) return rc;
      launch_pdl(k, grid, dim3(tc::kThreads), tc::WgCfg<128>::SMEM, s, tmX, tmY, M, N1, N2, kb_per_split, part, colsum_x ? cs_part : nullptr, passes);
    } else {
      auto k = tc::gemm_wgrad_tc_kernel<96>;
      if (int rc = tc::set_smem(k, tc::WgCfg<96>::SMEM)) return rc;
      launch_pdl(k, grid, dim3(tc::kThreads), tc::WgCfg<96>::SMEM, s, tmX, tmY, M, N1, N2, kb_per_split, part, colsum_x ? cs_part : nullptr, passes);
    }
    if (int rc = check_launch("gemm_wgrad_tc")) return rc;
  }
  if (colsum_x) return reduce_partials2(part, N1 * N2, out, cs_part, N1, colsum_x, splits, accumulate, s);
  return cnx_reduce_partials(part, splits, N1 * N2, 1.0f, accumulate, out, s);
}

}  // namespace cnx

// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (TMEM alloc / mma / commit / ld), UMMA shared-memory + instruction descriptors, and
// system-scope release/acquire used by the peer-memory (NVLink) kernels.
//
// Encodings follow the PTX ISA for tcgen05 and were cross-checked against the CuTe headers
// vendored in this image (cute/arch/mma_sm100_desc.hpp: SmemDescriptor / InstrDescriptor
// bit layouts; cute/atom/mma_traits_sm100.hpp: canonical K-major SWIZZLE_128B layout
// "Swizzle<3,4,3> o ((8,n),2):((8,SBO),1)" in 16-byte units).  Nothing here depends on
// CUTLASS at compile time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dfno {

// ------------------------------------------------------------------------------------------
// generic helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe of a phase (try_wait may suspend the thread for an implementation-defined time slice)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_test_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates are (c0 = innermost element index, c1 = row index)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 3-D tiled load: coordinates innermost first
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// tiled stores (shared -> global, bulk async-group completion); out-of-range parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// the issuing thread's committed stores have finished READING shared memory (it may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// ... and are complete (globally performed)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes (64 bf16
// along K), 8-row swizzle atoms of 1024 bytes stacked along M/N (SBO = 1024 B).  LBO is
// unused for swizzled K-major layouts (encoded as 1, as CuTe does).  bits[46,48) = 0b01 is
// the sm_100 descriptor version; bits[61,64) = 2 selects SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_smem_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);         // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                              // LBO (ignored)  [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                              // version = 1    [46,48)
  d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B   [61,64)
  return d;
}
// MN-major operand (the MN index is contiguous in memory), SWIZZLE_128B.  Canonical layout
// in 16-byte units: Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) -- a swizzle atom is
// 64 MN-elements (128 B) x 8 K-rows (1024 B); atoms repeat every SBO bytes along K (8 rows)
// and every LBO bytes along MN (64 elements).
__device__ __forceinline__ uint64_t umma_smem_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16 with BF16 A/B and FP32 accumulation.
//   [4,6) c_format (1 = F32) | [7,10) a_format (1 = BF16) | [10,13) b_format (1 = BF16)
//   [15] a_major (0 = K) | [16] b_major (0 = K, 1 = MN) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(uint32_t M, uint32_t N,
                                                           uint32_t a_mn_major = 0,
                                                           uint32_t b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// WHOLE-WARP variants: the converged warp calls these with warp-uniform operands and one elected lane issues.
// Computing the descriptors inside an `if (lane == 0)` region keeps them in vector registers, and the compiler
// then wraps every UTCHMMA in an ELECT / 5 x R2UR "waterfall" loop (~120 cycles per MMA on the single issuing
// thread, measured: profiles/r2_spin_probe.txt); warp-uniform operands live in uniform registers instead.
__device__ __forceinline__ void umma_bf16_ss_warp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  if (elect_one_sync()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// Same with the two K-major SWIZZLE_128B descriptors given by their LOW words (start address field + LBO); the
// high word (SBO = 1024 B, version, swizzle mode) is shared.  32-bit descriptor arithmetic stays in uniform registers.
constexpr uint32_t kUmmaK128Hi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_k128_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_bf16_ss_k128_warp(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                       uint32_t accumulate) {
  if (elect_one_sync()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaK128Hi)
        : "memory");
  }
}
// General form: both descriptors as (low, high) words.  MN-major SWIZZLE_128B: low = start address | LBO << 16,
// high = SBO | version | swizzle mode (umma_mn128_lo / umma_mn128_hi).
__device__ __forceinline__ uint32_t umma_mn128_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t umma_mn128_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ void umma_f16_ss_lohi_warp(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                      uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  if (elect_one_sync()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(b_hi)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  const uint32_t addr = smem_u32(bar);
  if (elect_one_sync()) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(addr) : "memory");
  }
}

// Make an mbarrier track completion of all previously issued tcgen05.mma of this thread
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
          smem_u32(bar))
      : "memory");
}

// TMEM -> registers: this thread's lane (32*(warp%4)+laneid), 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// registers -> TMEM (used to stage a bf16 A-operand in TMEM for TS-mode MMAs)
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// system-scope synchronisation for peer (NVLink) memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() {
  asm volatile("fence.acq_rel.sys;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------
// erf-GELU and its derivative (the reference uses F.gelu's default, exact erf form).
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. below fp32 round-off of the
// surrounding arithmetic and far below bf16 resolution): one MUFU.RCP, one MUFU.EX2 and a
// degree-5 Horner polynomial.  cdf and pdf share the same exponential exp(-x^2/2), so
// value + derivative cost ~16 instructions instead of two libm calls.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
struct GeluParts { float cdf; float pdf; };
__device__ __forceinline__ GeluParts gelu_parts(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, u, 1.0f));
  const float e = ex2_approx(-1.44269504088896341f * u * u);           // exp(-x^2 / 2)
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float half_tail = 0.5f * p * t * e;                           // 0.5 * erfc(|u|)
  GeluParts r;
  r.cdf = x >= 0.f ? 1.0f - half_tail : half_tail;
  r.pdf = 0.3989422804014327f * e;
  return r;
}
struct GeluVG { float value; float grad; };                 // gelu(x) and d gelu / dx from one evaluation
__device__ __forceinline__ float gelu_erf(float x) { return x * gelu_parts(x).cdf; }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const GeluParts g = gelu_parts(x);
  return fmaf(x, g.pdf, g.cdf);
}
__device__ __forceinline__ GeluVG gelu_value_grad(float x) {
  const GeluParts g = gelu_parts(x);
  return GeluVG{x * g.cdf, fmaf(x, g.pdf, g.cdf)};
}
// ------------------------------------------------------------------------------------------
// packed fp16 GELU: two values per instruction (HFMA2 / one MUFU.TANH.F16x2 per PAIR)
// ------------------------------------------------------------------------------------------
// The pointwise epilogues evaluate 10^9..10^10 GELUs per step and were issue bound with the fp32
// erf form (~17 instr + 2 MUFU per value).  This is the tanh form fitted to the *erf* GELU
//     Phi(x) ~ 0.5 (1 + tanh(x (a + b x^2 + c x^4))),  x^2 clamped at 64   (|gelu err| <= 2.6e-5 in exact
// arithmetic) evaluated in fp16x2: 7 instr + 1 MUFU per PAIR for the value, 14 + 1 for value and
// derivative.  fp16 (11-bit significand) keeps the absolute error of gelu / gelu' near 1e-3 * max(1,|x|),
// below the bf16 rounding (2^-9 relative) applied to every stored activation.
// Inputs beyond the fp16 range are handled by the clamp (x^2 = inf -> 64; tanh saturates).
#define DFNO_H2C(v) __float2half2_rn(v)
// saturating: |x| beyond the fp16 range becomes +-65504 instead of inf, so x * cdf(x) stays finite (0 or x)
__device__ __forceinline__ __half2 h2_from_f32(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ __half2 h2_tanh(__half2 x) {
  uint32_t r, xi = *reinterpret_cast<uint32_t*>(&x);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(xi));
  return *reinterpret_cast<__half2*>(&r);
}
struct GeluH2 { __half2 value; __half2 grad; };
__device__ __forceinline__ __half2 gelu_h2(__half2 x) {
  const __half2 x2 = __hmin2(__hmul2(x, x), DFNO_H2C(64.0f));
  __half2 g = __hfma2(DFNO_H2C(-3.51519787e-4f), x2, DFNO_H2C(3.70056658e-2f));
  g = __hfma2(g, x2, DFNO_H2C(7.97507861e-1f));
  const __half2 t = h2_tanh(__hmul2(x, g));
  return __hmul2(x, __hfma2(DFNO_H2C(0.5f), t, DFNO_H2C(0.5f)));
}
__device__ __forceinline__ GeluH2 gelu_vg_h2(__half2 x) {
  const __half2 x2 = __hmin2(__hmul2(x, x), DFNO_H2C(64.0f));
  __half2 g = __hfma2(DFNO_H2C(-3.51519787e-4f), x2, DFNO_H2C(3.70056658e-2f));
  g = __hfma2(g, x2, DFNO_H2C(7.97507861e-1f));
  const __half2 t = h2_tanh(__hmul2(x, g));
  const __half2 cdf = __hfma2(DFNO_H2C(0.5f), t, DFNO_H2C(0.5f));
  __half2 gp = __hfma2(DFNO_H2C(5.0f * -3.51519787e-4f), x2, DFNO_H2C(3.0f * 3.70056658e-2f));
  gp = __hfma2(gp, x2, DFNO_H2C(7.97507861e-1f));                 // d/dx [x g(x^2)]
  const __half2 s = __hfma2(__hneg2(t), t, DFNO_H2C(1.0f));       // sech^2
  const __half2 xs = __hmul2(__hmul2(x, s), DFNO_H2C(0.5f));
  GeluH2 r;
  r.value = __hmul2(x, cdf);
  r.grad = __hfma2(xs, gp, cdf);
  return r;
}
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 h2_of_bits(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
// fp16x2 <-> bf16x2 (through fp32; values outside the fp16 range saturate to inf and are clamped by the GELU)
__device__ __forceinline__ uint32_t h2_to_bf16x2(__half2 v) {
  const float2 f = __half22float2(v);
  __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
  return *reinterpret_cast<uint32_t*>(&b);
}
__device__ __forceinline__ __half2 bf16x2_to_h2(uint32_t u) {
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u);
  const float2 f = __bfloat1622float2(b);
  return h2_from_f32(f.x, f.y);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

}  // namespace dfno

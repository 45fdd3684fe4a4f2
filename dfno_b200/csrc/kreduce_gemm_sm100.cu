// kreduce_gemm_sm100.cu -- long-K reduction GEMM on tcgen05:  D[i, j] += sum_k A[i, k] * B[j, k]
//
// A: [Ma <= 128, K], B: [Nb <= 256, K], both bf16 with the *reduced* index contiguous, K in
// the millions (all field positions).  Used for the weight gradients of the 1x1 convolutions
// (dW[o, i] = sum_pos dpre[o, pos] * h[i, pos], the SumReduce side of BroadcastedLinear,
// SURVEY.md K18): the activations are already stored channel-major with positions
// contiguous, so both operands are K-major as they are -- no transpose, no im2col.
//
// Split-K over persistent CTAs: each CTA streams a contiguous range of 64-wide K blocks
// through a TMA/mbarrier ring (rows beyond Ma/Nb are zero-filled by TMA and cost no HBM
// traffic), accumulates one 128 x Nb_pad tile in TMEM over its whole range, then adds it to
// the fp32 result with atomics.  Memory bound: ~ (Ma + Nb) * 128 B per 64 K-steps.
#include "sm100_ptx.cuh"
#include "kernels.h"
#include "tma_host.h"

namespace dfno {

namespace {
constexpr int kStages = 8;
constexpr int kThreads = 192;
constexpr uint32_t kCols = 256;

struct KrParams {
  int Ma, Nb, nb_pad;
  int a_rows;              // rows of the A box (32 when Ma <= 32: the UMMA still reads 128 rows, the
                           // extra accumulator lanes hold don't-care values that are never read)
  long long kblocks;       // total 64-wide K blocks
  float* D;
  long long ldd;
};

__global__ void __launch_bounds__(kThreads, 1)
kreduce_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KrParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t a_bytes = p.a_rows * 128, b_bytes = p.nb_pad * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes + 16384);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* done = bars + 2 * kStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // contiguous K range of this CTA
  const long long per = (p.kblocks + gridDim.x - 1) / gridDim.x;
  const long long kb0 = per * blockIdx.x;
  const long long kb1 = kb0 + per < p.kblocks ? kb0 + per : p.kblocks;
  const bool has_work = kb0 < kb1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kCols>(tmem_holder);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (has_work) {
    if (warp == 0) {
      if (lane == 0) {
        uint32_t s = 0, ph = 0;
        for (long long kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], stage_bytes);
          uint8_t* dst = smem + s * stage_bytes;
          tma_load_2d(dst, &tmA, &full[s], static_cast<int32_t>(kb * 64), 0);
          tma_load_2d(dst + a_bytes, &tmB, &full[s], static_cast<int32_t>(kb * 64), 0);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = umma_idesc_bf16_f32(128, p.nb_pad);
      uint32_t s = 0, ph = 0;
      for (long long kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_u32(smem + s * stage_bytes);
          const uint32_t b_base = a_base + a_bytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ss(tmem_base, umma_smem_desc_k128(a_base + kk * 32), umma_smem_desc_k128(b_base + kk * 32),
                         idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (kb == kb1 - 1) umma_commit(done);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    } else {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      mbar_wait(done, 0);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < p.Nb; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + c0, v);
        tmem_ld_wait();
        if (row < p.Ma) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < p.Nb) atomicAdd(p.D + row * p.ldd + c0 + i, __uint_as_float(v[i]));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kCols>(tmem_base);
}
}  // namespace

const char* kreduce_gemm(const void* A, long long lda, int Ma, const void* Bm, long long ldb, int Nb, long long K,
                         float* D, long long ldd, int num_sms, cudaStream_t stream) {
  if (Ma < 1 || Ma > 128 || Nb < 1 || Nb > 256) return "kreduce: Ma<=128, Nb<=256";
  if ((lda * 2) % 16 || (ldb * 2) % 16) return "kreduce: row pitches must be multiples of 16 bytes";
  if (K > (1ll << 31) - 64) return "kreduce: K too large for one launch";
  KrParams p;
  p.Ma = Ma; p.Nb = Nb; p.nb_pad = (Nb + 15) / 16 * 16;
  p.a_rows = Ma <= 32 ? 32 : 128;
  p.kblocks = (K + 63) / 64;
  p.D = D; p.ldd = ldd;
  CUtensorMap tmA, tmB;
  if (make_map_2d(&tmA, A, static_cast<uint64_t>(K), static_cast<uint64_t>(Ma), static_cast<uint64_t>(lda), 64,
                  static_cast<uint32_t>(p.a_rows)))
    return "cuTensorMapEncodeTiled(A) failed";
  if (make_map_2d(&tmB, Bm, static_cast<uint64_t>(K), static_cast<uint64_t>(Nb), static_cast<uint64_t>(ldb), 64,
                  static_cast<uint32_t>(p.nb_pad)))
    return "cuTensorMapEncodeTiled(B) failed";
  // + 16 KB slack: the M = 128 descriptor of the last stage reads past its 32-row box
  uint32_t smem_bytes = kStages * (p.a_rows * 128 + p.nb_pad * 128) + 256 + 16384;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kreduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "cudaFuncSetAttribute failed";
    attr_set = true;
  }
  // at least ~64 K-blocks per CTA so the per-CTA atomics stay negligible
  long long ctas = (p.kblocks + 63) / 64;
  if (ctas > num_sms) ctas = num_sms;
  if (ctas < 1) ctas = 1;
  kreduce_kernel<<<static_cast<int>(ctas), kThreads, smem_bytes, stream>>>(tmA, tmB, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

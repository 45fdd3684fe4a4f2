// head_bwd_sm100.cu -- backward of the projection head  out = W4 . gelu(W3 h + b3) + b4
// (reference linear3 -> gelu -> linear4, dfno.py:348-351) as ONE tcgen05 kernel.
//
// The 128-channel hidden layer is never written to memory: per tile of 128 field positions
//
//   MMA1   pre[pos, j]   = h[pos, :] . W3[j, :]                (A = h tile via TMA, K-major)
//   epi A  g[pos, j]     = dout[pos] * W4[j] * gelu'(pre + b3) (thread = position, TMEM lane)
//          db3 / dW4 / db4 partial sums (warp transpose-reduce), g -> bf16 -> smem tile P
//   MMA2   dh[pos, i]    = sum_j g[pos, j] * W3[j, i]          (A = P, K-major over j)
//   MMA3   dW3[j, i]    += sum_pos g[pos, j] * h[pos, i]       (A = P and B = h tile re-read as
//                                                               MN-major operands: K = positions)
//   epi B  dh -> global (channels-last), dW3 stays in TMEM across all tiles of the CTA.
//
// MMA3 is where the SumReduce side of the reference's BroadcastedLinear ends up: a K-reduction
// over all positions accumulated in TMEM, flushed once per CTA with atomics.
#include "sm100_ptx.cuh"
#include "kernels.h"
#include "tma_host.h"

namespace dfno {
namespace {

constexpr int kStagesH = 3;
constexpr int kEpiH = 4;                 // epilogue warps per TMEM lane quarter (one 32-column slice each)
constexpr int kThreadsH = 64 + 128 * kEpiH;
constexpr int kHid = 128;
constexpr uint32_t kColsH = 512;
// TMEM columns
constexpr uint32_t kD1 = 0;      // 2 x 128 : pre-activations (double buffered)
constexpr uint32_t kD2 = 256;    // 2 x 32  : dh tiles (double buffered with P)
constexpr uint32_t kD3 = 320;    // 32      : dW3 accumulator [hid lanes, c]

struct HeadBwdParams {
  long long npos;
  int C, CP;
  const float* dout;          // fp32, addressed through the row digits below (public layout)
  int nrl; int R[4]; long long SR[4];
  const float* b3; const float* W4;
  __nv_bfloat16* gcl;         // [npos, CP]
  float* gW3; float* gb3; float* gW4; float* gb4;
};

// lane l ends up with sum over lanes of v[l]  (v is destroyed)
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(kThreadsH, 1)
head_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW3,
                const __grid_constant__ CUtensorMap tmW3T, const HeadBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_w3 = smem;                          // [128 hid][64 c]   K-major (K = c)      16 KB
  uint8_t* smem_w3t = smem + 16384;                 // 2 x [32 c][64 hid] K-major (K = hid)    8 KB
  uint8_t* smem_p = smem + 24576;                   // 2 buffers x 2 x [128 pos][64 hid]      64 KB
  uint8_t* smem_a = smem + 90112;                   // stages x [128 pos][64 c]               48 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + kStagesH * 16384);
  uint64_t* a_full = bars;            // [3]
  uint64_t* a_empty = bars + 3;       // [3]
  uint64_t* w_full = bars + 6;
  uint64_t* d1_full = bars + 7;       // [2]
  uint64_t* d1_empty = bars + 9;      // [2]
  uint64_t* p_full = bars + 11;       // [2]
  uint64_t* d2_full = bars + 13;      // [2]
  uint64_t* all_done = bars + 15;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 16);
  float* s_b3 = reinterpret_cast<float*>(bars + 20);   // [128]
  float* s_w4 = s_b3 + 128;                            // [128]
  float* s_gb3 = s_w4 + 128;                           // [128] CTA partial sums
  float* s_gw4 = s_gb3 + 128;                          // [128]
  float* s_gb4 = s_gw4 + 128;                          // [1]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = static_cast<int>((p.npos + 127) / 128);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW3); tma_prefetch_desc(&tmW3T);
    for (int s = 0; s < kStagesH; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
    mbar_init(&d1_full[0], 1); mbar_init(&d1_full[1], 1);
    mbar_init(&d1_empty[0], 4 * kEpiH); mbar_init(&d1_empty[1], 4 * kEpiH);
    mbar_init(&p_full[0], 4 * kEpiH); mbar_init(&p_full[1], 4 * kEpiH);
    mbar_init(&d2_full[0], 1); mbar_init(&d2_full[1], 1);
    mbar_init(all_done, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128; i += kThreadsH) {
    s_b3[i] = p.b3[i]; s_w4[i] = p.W4[i]; s_gb3[i] = 0.f; s_gw4[i] = 0.f;
  }
  if (threadIdx.x == 0) s_gb4[0] = 0.f;
  if (warp == 1) tmem_alloc<kColsH>(tmem_holder);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 16384 + 8192);
      tma_load_2d(smem_w3, &tmW3, w_full, 0, 0);
      tma_load_2d(smem_w3t, &tmW3T, w_full, 0, 0);
      tma_load_2d(smem_w3t + 4096, &tmW3T, w_full, 64, 0);
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&a_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&a_full[s], 16384);
        tma_load_2d(smem_a + s * 16384, &tmA, &a_full[s], 0, tile * 128);
        if (++s == kStagesH) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc1 = umma_idesc_bf16_f32(128, 128);            // pre  = h . W3^T
    const uint32_t idesc2 = umma_idesc_bf16_f32(128, 32);             // dh   = P . W3T^T
    const uint32_t idesc3 = umma_idesc_bf16_f32(128, 32, 1, 1);       // dW3 += P^T . h   (MN-major A and B)
    const int k1steps = (p.C + 15) / 16;
    mbar_wait(w_full, 0);
    uint32_t s = 0, ph = 0;
    int n = 0;
    uint32_t prev_stage = 0;
    auto part2 = [&](int m, uint32_t stage) {
      const int pb = m & 1;                                   // P / D2 buffer of this tile
      mbar_wait(&p_full[pb], (m >> 1) & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t pbase = smem_u32(smem_p + pb * 32768);
        const uint32_t abase = smem_u32(smem_a + stage * 16384);
        const uint32_t wtbase = smem_u32(smem_w3t);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // K = hid, 16 per instruction; 64-wide blocks of P / W3T
          const int kb = ks >> 2, kk = ks & 3;
          umma_bf16_ss(tmem_base + kD2 + pb * 32, umma_smem_desc_k128(pbase + kb * 16384 + kk * 32),
                       umma_smem_desc_k128(wtbase + kb * 4096 + kk * 32), idesc2, ks > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // K = positions, 16 rows (2 swizzle-atom rows groups) per instruction
          umma_bf16_ss(tmem_base + kD3, umma_smem_desc_mn128(pbase + ks * 2048, 16384, 1024),
                       umma_smem_desc_mn128(abase + ks * 2048, 16384, 1024), idesc3, (m > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&d2_full[pb]);
        umma_commit(&a_empty[stage]);
      }
      __syncwarp();
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      const int buf = n & 1;
      mbar_wait(&d1_empty[buf], ((n >> 1) & 1) ^ 1);
      mbar_wait(&a_full[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t abase = smem_u32(smem_a + s * 16384);
        const uint32_t wbase = smem_u32(smem_w3);
        for (int ks = 0; ks < k1steps; ++ks)
          umma_bf16_ss(tmem_base + kD1 + buf * 128, umma_smem_desc_k128(abase + ks * 32),
                       umma_smem_desc_k128(wbase + ks * 32), idesc1, ks > 0 ? 1u : 0u);
        umma_commit(&d1_full[buf]);
      }
      __syncwarp();
      if (n > 0) part2(n - 1, prev_stage);
      prev_stage = s;
      if (++s == kStagesH) { s = 0; ph ^= 1; }
    }
    if (n > 0) part2(n - 1, prev_stage);
    if (lane == 0) umma_commit(all_done);
    __syncwarp();
  } else {
    // ===================== epilogue warps (thread = field position = TMEM lane) ==========
    const int q = warp & 3;
    const int e = (warp - 2) >> 2;               // this warp's 32-column slice of the hidden layer
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float acc_gb4 = 0.f;
    int n = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      const int buf = n & 1;
      const long long row = static_cast<long long>(tile) * 128 + r_in_tile;
      const bool row_ok = row < p.npos;
      float dout = 0.f;
      if (row_ok) {
        long long roff = 0;
        uint32_t r = static_cast<uint32_t>(row);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          if (l < p.nrl) {
            uint32_t d = r;
            if (l != p.nrl - 1) { const uint32_t qq = r / static_cast<uint32_t>(p.R[l]); d = r - qq * p.R[l]; r = qq; }
            roff += static_cast<long long>(d) * p.SR[l];
          }
        }
        dout = p.dout[roff];
      }
      if (e == 0) acc_gb4 += dout;
      mbar_wait(&d1_full[buf], (n >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t t1 = tmem_base + lane_addr + kD1 + buf * 128;
      const int pb = n & 1;
      // the buffer was last read by the MMAs of tile n-2: already retired unless we run far ahead
      if (n >= 2) mbar_wait(&d2_full[pb], ((n >> 1) - 1) & 1);
      uint8_t* prow = smem_p + pb * 32768 + r_in_tile * 128;
      {
        const int c0 = 32 * e;
        float gsum[32], wsum[32];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t1 + c0 + h2 * 16, v);
          tmem_ld_wait();
          uint32_t packed[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float g2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = c0 + h2 * 16 + i + e;
              const float pre = __uint_as_float(v[i + e]) + s_b3[j];
              const GeluVG gv = gelu_value_grad(pre);
              const float act = gv.value;
              const float g = dout * s_w4[j] * gv.grad;
              wsum[h2 * 16 + i + e] = dout * act;       // -> dW4[j]
              gsum[h2 * 16 + i + e] = g;                // -> db3[j]
              g2[e] = g;
            }
            packed[i >> 1] = pack_bf16x2(g2[0], g2[1]);
          }
          // 16 bf16 = two 16-byte chunks of this row in the 64-wide hid block, SWIZZLE_128B
          const int jbase = c0 + h2 * 16;
          const int kb = jbase >> 6;
          const int chunk = (jbase & 63) >> 3;
          uint8_t* blk = prow + kb * 16384;
          *reinterpret_cast<uint4*>(blk + (((chunk) ^ (r_in_tile & 7)) << 4)) =
              make_uint4(packed[0], packed[1], packed[2], packed[3]);
          *reinterpret_cast<uint4*>(blk + (((chunk + 1) ^ (r_in_tile & 7)) << 4)) =
              make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
        const float sg = warp_transpose_reduce32(gsum, lane);
        const float sw = warp_transpose_reduce32(wsum, lane);
        atomicAdd(&s_gb3[c0 + lane], sg);
        atomicAdd(&s_gw4[c0 + lane], sw);
      }
      // P is complete and D1[buf] fully read: publish to the async proxy / MMA warp
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&d1_empty[buf]); mbar_arrive(&p_full[pb]); }
      // ---- dh tile: only the e == 0 warps read it back
      if (e == 0) {
        mbar_wait(&d2_full[pb], (n >> 1) & 1);
        tcgen05_fence_after();
        uint32_t v[16], w[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + kD2 + pb * 32, v);
        tmem_ld_32x32b_x16(tmem_base + lane_addr + kD2 + pb * 32 + 16, w);
        tmem_ld_wait();
        if (row_ok) {
          __nv_bfloat16* o = p.gcl + row * p.CP;
          for (int c = 0; c < p.CP; c += 8) {
            uint32_t u[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int cc = c + 2 * i;
              const float a = cc < 16 ? __uint_as_float(v[cc & 15]) : __uint_as_float(w[cc & 15]);
              const float b = cc + 1 < 16 ? __uint_as_float(v[(cc + 1) & 15]) : __uint_as_float(w[(cc + 1) & 15]);
              u[i] = pack_bf16x2(cc < p.C ? a : 0.f, cc + 1 < p.C ? b : 0.f);
            }
            *reinterpret_cast<uint4*>(o + c) = make_uint4(u[0], u[1], u[2], u[3]);
          }
        }
      }
      tcgen05_fence_before();
    }
    // ---- per-CTA flush of the weight gradients
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 16);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 8);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 4);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 2);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 1);
    if (lane == 0 && e == 0) atomicAdd(s_gb4, acc_gb4);
    asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiH) : "memory");
    if (n > 0 && e == 0) {
      mbar_wait(all_done, 0);                    // every MMA of this CTA has retired
      tcgen05_fence_after();
      uint32_t v[16], w[16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + kD3, v);
      tmem_ld_32x32b_x16(tmem_base + lane_addr + kD3 + 16, w);
      tmem_ld_wait();
      const int j = r_in_tile;                       // TMEM lane = hidden unit
      for (int c = 0; c < p.C; ++c)
        atomicAdd(p.gW3 + j * p.C + c, c < 16 ? __uint_as_float(v[c & 15]) : __uint_as_float(w[c & 15]));
      atomicAdd(p.gb3 + j, s_gb3[j]);
      atomicAdd(p.gW4 + j, s_gw4[j]);
      if (j == 0) atomicAdd(p.gb4, s_gb4[0]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kColsH>(tmem_base);
}

}  // namespace

const char* head_bwd(const void* hcl, long long npos, int C, int CP, const void* W3pad, const void* W3Tpad,
                     const float* b3, const float* W4, const float* dout, int nrl, const int* R, const long long* SR,
                     void* gcl, float* gW3, float* gb3, float* gW4, float* gb4, int num_sms, cudaStream_t stream) {
  if (C > 32 || CP % 8 || CP < C || CP > 64) return "head_bwd: need C <= 32 and an 8-aligned channels-last pitch <= 64";
  if (npos > (1ll << 31) - 256) return "head_bwd: too many positions for one launch";
  HeadBwdParams p;
  p.npos = npos; p.C = C; p.CP = CP; p.dout = dout; p.nrl = nrl;
  for (int i = 0; i < 4; ++i) { p.R[i] = i < nrl ? R[i] : 1; p.SR[i] = i < nrl ? SR[i] : 0; }
  p.b3 = b3; p.W4 = W4; p.gcl = static_cast<__nv_bfloat16*>(gcl);
  p.gW3 = gW3; p.gb3 = gb3; p.gW4 = gW4; p.gb4 = gb4;
  CUtensorMap tmA, tmW3, tmW3T;
  if (make_map_2d(&tmA, hcl, static_cast<uint64_t>(C), static_cast<uint64_t>(npos), static_cast<uint64_t>(CP), 64, 128))
    return "cuTensorMapEncodeTiled(h) failed";
  if (make_map_2d(&tmW3, W3pad, 64, 128, 64, 64, 128)) return "cuTensorMapEncodeTiled(W3) failed";
  if (make_map_2d(&tmW3T, W3Tpad, 128, 32, 128, 64, 32)) return "cuTensorMapEncodeTiled(W3T) failed";
  const uint32_t smem_bytes = 90112 + kStagesH * 16384 + 4096;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "cudaFuncSetAttribute failed";
    attr_set = true;
  }
  const int num_tiles = static_cast<int>((npos + 127) / 128);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  head_bwd_kernel<<<grid, kThreadsH, smem_bytes > 120 * 1024 ? smem_bytes : 120 * 1024, stream>>>(tmA, tmW3, tmW3T, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

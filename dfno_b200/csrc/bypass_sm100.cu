// bypass_sm100.cu -- the 1x1 "bypass" convolution of a Fourier layer fused with the GELU
// (forward: SURVEY.md K2 + K14, reference dfno.py:244,291; backward: its adjoint plus the weight
// gradient) on tcgen05, TMA in *and* out.
//
//   forward   pre[o, pos] = spec[o, pos] + sum_i W[o, i] h[i, pos];   out = gelu(pre)
//   backward  g[o, pos]   = dout[o, pos] * gelu'(pre[o, pos])          (written over pre)
//             dhb[i, pos] = sum_o W[o, i] g[o, pos]
//             dW[o, i]   += sum_pos g[o, pos] h[i, pos]                (accumulated in TMEM)
//
// Activations are channel-major ([b*C + c][S positions], positions contiguous).  A tile is
// 128 consecutive positions of all C channels of one batch element: TMA drops it into
// shared memory as two SWIZZLE_128B boxes of [C rows][64 positions].  That single image is
// used three ways without ever being transposed:
//   * as an MN-major A operand (M = positions, K = channels)   -> channel mixing, positions on
//     the TMEM lanes, so the epilogue thread of a position owns all its channels;
//   * as a K-major operand (rows = channels, K = positions)     -> the weight gradient, a
//     K-reduction over every position the CTA visits, accumulated in TMEM;
//   * as the source of a TMA store after the epilogue rewrote it in place.
// Rows C..31 of every box are zeroed once and never touched again (boxes have exactly C rows),
// so the padded K range contributes exact zeros.
#include "sm100_ptx.cuh"
#include "kernels.h"
#include "tma_host.h"

namespace dfno {
namespace {

constexpr int kThreadsB = 64 + 128 * 4;        // TMA warp, MMA warp, 4 epilogue groups
constexpr int kGroups = 4;
constexpr uint32_t kColsB = 512;
constexpr uint32_t kTile = 8192;               // one tensor tile: 2 boxes x 32 rows x 128 B

__device__ __forceinline__ void tma_store_commit_and_wait() {
  tma_store_commit();
  tma_store_wait_read();
}
// byte offset of element (channel c, position t in [0,128)) inside a tile image
__device__ __forceinline__ uint32_t tile_off(int c, int t) {
  const int j = t >> 6, cp = t & 63;
  return j * 4096 + c * 128 + ((((cp >> 3) ^ (c & 7)) << 4) | ((cp & 7) << 1));
}

struct BypassParams {
  int B, C;
  long long S;              // positions per (b, c) slab; multiple of 128
  int save_pre;
  int cl_pitch;
  __nv_bfloat16* out_cl;    // forward: optional channels-last output (instead of the TMA-stored `out`)
  const __nv_bfloat16* dout_cl;   // backward: optional channels-last incoming gradient
  float* dW;                // backward: [C, C] fp32, accumulated with atomics
};

// ================================================================================ forward
constexpr int kStagesF = 6;                    // 16 KB each: h tile + spec tile

__global__ void __launch_bounds__(kThreadsB, 1)
bypass_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmS,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmW,
                     const BypassParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_w = smem;                                   // [32 o][64 i] K-major, 4 KB
  uint8_t* stage0 = smem + 4096;                            // kStagesF x {h tile 8 KB, spec tile 8 KB}
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + kStagesF * 2 * kTile);
  uint64_t* full = bars;                 // [6]
  uint64_t* empty = bars + 6;            // [6]   arrived by the epilogue after its TMA stores drained
  uint64_t* tfull = bars + 12;           // [8]
  uint64_t* tempty = bars + 20;          // [8]
  uint64_t* wfull = bars + 28;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 29);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles_per_b = p.S / 128;
  const long long num_tiles = tiles_per_b * p.B;
  constexpr int nacc = 8;

  // zero the padding rows of every tile image once
  for (uint32_t i = threadIdx.x; i < kStagesF * 2 * kTile / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(stage0)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmW);
    for (int s = 0; s < kStagesF; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < nacc; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kColsB>(tmem_holder);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, 4096);
      tma_load_2d(smem_w, &tmW, wfull, 0, 0);
      uint32_t s = 0, ph = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = static_cast<int>(tile / tiles_per_b);
        const int p0 = static_cast<int>((tile % tiles_per_b) * 128);
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], 4u * p.C * 128);
        uint8_t* st = stage0 + s * 2 * kTile;
        tma_load_2d(st, &tmH, &full[s], p0, b * p.C);
        tma_load_2d(st + 4096, &tmH, &full[s], p0 + 64, b * p.C);
        tma_load_2d(st + kTile, &tmS, &full[s], p0, b * p.C);
        tma_load_2d(st + kTile + 4096, &tmS, &full[s], p0 + 64, b * p.C);
        if (++s == kStagesF) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, 32, /*a MN-major*/ 1, 0);
    mbar_wait(wfull, 0);
    uint32_t s = 0, ph = 0;
    long long n = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      const int a = static_cast<int>(n % nacc);
      mbar_wait(&tempty[a], ((n / nacc) & 1) ^ 1);
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t hbase = smem_u32(stage0 + s * 2 * kTile);
        const uint32_t wbase = smem_u32(smem_w);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)            // K = 32 channels, 16 (= two 8-row groups) per MMA
          umma_bf16_ss(tmem_base + a * 32, umma_smem_desc_mn128(hbase + ks * 2048, 4096, 1024),
                       umma_smem_desc_k128(wbase + ks * 32), idesc, ks > 0 ? 1u : 0u);
        umma_commit(&tfull[a]);
      }
      __syncwarp();
      if (++s == kStagesF) { s = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int t = q * 32 + lane;                            // position inside the tile = TMEM lane
    long long n = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      if (n % kGroups != g) continue;
      const int a = static_cast<int>(n % nacc);
      const uint32_t s = static_cast<uint32_t>(n % kStagesF);
      const int b = static_cast<int>(tile / tiles_per_b);
      const int p0 = static_cast<int>((tile % tiles_per_b) * 128);
      uint8_t* ht = stage0 + s * 2 * kTile;
      uint8_t* st = ht + kTile;
      mbar_wait(&full[s], (n / kStagesF) & 1);              // TMA data visible to this thread
      mbar_wait(&tfull[a], (n / nacc) & 1);                 // channel mixing of the tile retired
      tcgen05_fence_after();
      uint32_t v[32];
      {
        uint32_t v0[16], v1[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 32, v0);
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 32 + 16, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = v0[i]; v[16 + i] = v1[i]; }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      __nv_bfloat16* clrow = p.out_cl ? p.out_cl + (static_cast<long long>(b) * p.S + p0 + t) * p.cl_pitch : nullptr;
      uint32_t pk[16];                                      // channels-last row, two channels per word
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        if (o < p.C) {
          const uint32_t off = tile_off(o, t);
          const float pre = __uint_as_float(v[o]) + __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(st + off));
          *reinterpret_cast<__nv_bfloat16*>(st + off) = __float2bfloat16(pre);
          const __nv_bfloat16 y = __float2bfloat16(gelu_erf(pre));
          if (clrow) {
            const uint32_t bits = __bfloat16_as_ushort(y);
            if (o & 1) pk[o >> 1] |= bits << 16; else pk[o >> 1] = bits;
          } else {
            *reinterpret_cast<__nv_bfloat16*>(ht + off) = y;
          }
        } else if ((o & 1) == 0) {
          pk[o >> 1] = 0;
        }
      }
      if (clrow) {                                          // 16-byte vectors; the pitch is a multiple of 8 channels
#pragma unroll
        for (int w = 0; w < 4; ++w)
          if (8 * w < p.cl_pitch)
            reinterpret_cast<uint4*>(clrow)[w] = make_uint4(pk[4 * w], pk[4 * w + 1], pk[4 * w + 2], pk[4 * w + 3]);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
      if (t == 0) {
        if (p.save_pre) {
          tma_store_2d(&tmS, st, p0, b * p.C);
          tma_store_2d(&tmS, st + 4096, p0 + 64, b * p.C);
        }
        if (!p.out_cl) {
          tma_store_2d(&tmO, ht, p0, b * p.C);
          tma_store_2d(&tmO, ht + 4096, p0 + 64, b * p.C);
        }
        tma_store_commit_and_wait();                        // smem may be refilled
        mbar_arrive(&empty[s]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kColsB>(tmem_base);
}

// ================================================================================ backward
constexpr int kStagesBw = 4;                   // 24 KB each: pre tile, dout tile, h tile
constexpr uint32_t kD2 = 0;                    // 4 x 32 columns: dhb tiles (one per epilogue group)
constexpr uint32_t kD3 = 128;                  // 32 columns: dW accumulator [o lanes, i]

__global__ void __launch_bounds__(kThreadsB, 1)
bypass_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmG,
                     const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmD,
                     const __grid_constant__ CUtensorMap tmWT, const BypassParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_wt = smem;                                  // [32 i][64 o] K-major, 4 KB
  uint8_t* stage0 = smem + 4096;                            // kStagesBw x {pre, dout, h} tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + kStagesBw * 3 * kTile + 16384 /*M=128 over-read slack*/);
  uint64_t* full = bars;                 // [4]
  uint64_t* empty = bars + 4;            // [4]
  uint64_t* pfull = bars + 8;            // [4] per group: g tile written
  uint64_t* d2full = bars + 12;          // [4] per group: MMA2/3 of its tile retired
  uint64_t* wfull = bars + 16;
  uint64_t* done = bars + 17;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 18);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles_per_b = p.S / 128;
  const long long num_tiles = tiles_per_b * p.B;

  for (uint32_t i = threadIdx.x; i < (kStagesBw * 3 * kTile + 16384) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(stage0)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP); tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmWT);
    for (int s = 0; s < kStagesBw; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int g = 0; g < kGroups; ++g) { mbar_init(&pfull[g], 4); mbar_init(&d2full[g], 1); }
    mbar_init(wfull, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kColsB>(tmem_holder);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const bool cl_in = p.dout_cl != nullptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, 4096);
      tma_load_2d(smem_wt, &tmWT, wfull, 0, 0);
      uint32_t s = 0, ph = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = static_cast<int>(tile / tiles_per_b);
        const int p0 = static_cast<int>((tile % tiles_per_b) * 128);
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], (cl_in ? 4u : 6u) * p.C * 128);
        uint8_t* st = stage0 + s * 3 * kTile;
        tma_load_2d(st, &tmP, &full[s], p0, b * p.C);
        tma_load_2d(st + 4096, &tmP, &full[s], p0 + 64, b * p.C);
        if (!cl_in) {
          tma_load_2d(st + kTile, &tmG, &full[s], p0, b * p.C);
          tma_load_2d(st + kTile + 4096, &tmG, &full[s], p0 + 64, b * p.C);
        }
        tma_load_2d(st + 2 * kTile, &tmH, &full[s], p0, b * p.C);
        tma_load_2d(st + 2 * kTile + 4096, &tmH, &full[s], p0 + 64, b * p.C);
        if (++s == kStagesBw) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc2 = umma_idesc_bf16_f32(128, 32, /*a MN-major*/ 1, 0);   // dhb = g^T-view . WT
    const uint32_t idesc3 = umma_idesc_bf16_f32(128, 32, 0, 0);                  // dW += g . h^T  (K = positions)
    mbar_wait(wfull, 0);
    long long n = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      const int g = static_cast<int>(n % kGroups);
      const uint32_t s = static_cast<uint32_t>(n % kStagesBw);
      mbar_wait(&pfull[g], (n / kGroups) & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t gbase = smem_u32(stage0 + s * 3 * kTile);          // g tile (over the pre tile)
        const uint32_t hbase = gbase + 2 * kTile;
        const uint32_t wbase = smem_u32(smem_wt);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)            // K = 32 output channels o
          umma_bf16_ss(tmem_base + kD2 + g * 32, umma_smem_desc_mn128(gbase + ks * 2048, 4096, 1024),
                       umma_smem_desc_k128(wbase + ks * 32), idesc2, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {          // K = 128 positions: 2 boxes x 4 steps of 16
          const int j = ks >> 2, kk = ks & 3;
          umma_bf16_ss(tmem_base + kD3, umma_smem_desc_k128(gbase + j * 4096 + kk * 32),
                       umma_smem_desc_k128(hbase + j * 4096 + kk * 32), idesc3, (n > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&d2full[g]);
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(done);
    __syncwarp();
  } else {
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int t = q * 32 + lane;
    long long n = 0, mine = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      if (n % kGroups != g) continue;
      const uint32_t s = static_cast<uint32_t>(n % kStagesBw);
      const int b = static_cast<int>(tile / tiles_per_b);
      const int p0 = static_cast<int>((tile % tiles_per_b) * 128);
      uint8_t* pt = stage0 + s * 3 * kTile;
      uint8_t* gt = pt + kTile;
      mbar_wait(&full[s], (n / kStagesBw) & 1);
      uint32_t pk[16];
      if (cl_in) {                                          // this position's channels-last gradient row
        const uint4* clrow = reinterpret_cast<const uint4*>(p.dout_cl + (static_cast<long long>(b) * p.S + p0 + t) * p.cl_pitch);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint4 u = make_uint4(0, 0, 0, 0);
          if (8 * w < p.cl_pitch) u = clrow[w];
          pk[4 * w] = u.x; pk[4 * w + 1] = u.y; pk[4 * w + 2] = u.z; pk[4 * w + 3] = u.w;
        }
      }
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        if (o < p.C) {
          const uint32_t off = tile_off(o, t);
          const float pre = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(pt + off));
          const float dy = cl_in ? __uint_as_float((o & 1) ? (pk[o >> 1] & 0xffff0000u) : (pk[o >> 1] << 16))
                                 : __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(gt + off));
          *reinterpret_cast<__nv_bfloat16*>(pt + off) = __float2bfloat16(dy * gelu_erf_grad(pre));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[g]);
      mbar_wait(&d2full[g], mine & 1);
      ++mine;
      tcgen05_fence_after();
      {
        uint32_t v0[16], v1[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kD2 + g * 32, v0);
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kD2 + g * 32 + 16, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < p.C)
            *reinterpret_cast<__nv_bfloat16*>(gt + tile_off(i, t)) =
                __float2bfloat16(__uint_as_float(i < 16 ? v0[i & 15] : v1[i & 15]));
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
      if (t == 0) {
        tma_store_2d(&tmP, pt, p0, b * p.C);                // dpre over pre
        tma_store_2d(&tmP, pt + 4096, p0 + 64, b * p.C);
        tma_store_2d(&tmD, gt, p0, b * p.C);                // dhb
        tma_store_2d(&tmD, gt + 4096, p0 + 64, b * p.C);
        tma_store_commit_and_wait();
        mbar_arrive(&empty[s]);
      }
    }
    if (g == 0) {
      mbar_wait(done, 0);
      tcgen05_fence_after();
      uint32_t v0[16], v1[16];
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kD3, v0);
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kD3 + 16, v1);
      tmem_ld_wait();
      if (t < p.C && num_tiles > blockIdx.x)
        for (int i = 0; i < p.C; ++i)
          atomicAdd(p.dW + t * p.C + i, __uint_as_float(i < 16 ? v0[i & 15] : v1[i & 15]));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kColsB>(tmem_base);
}

const char* set_attr_once(const void* fn, bool* flag) {
  if (!*flag) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "cudaFuncSetAttribute failed";
    *flag = true;
  }
  return nullptr;
}

}  // namespace

const char* bypass_fwd_tc(const void* h, void* spec_pre, const void* Wpad, void* out, void* out_cl, int cl_pitch,
                          int B, int C, long long S, int save_pre, int num_sms, cudaStream_t stream) {
  if (C > 32 || S % 128) return "bypass_fwd_tc: need C <= 32 and S % 128 == 0";
  if (S > (1ll << 31) - 256) return "bypass_fwd_tc: slab too large";
  if (!out && !out_cl) return "bypass_fwd_tc: no output";
  BypassParams p{};
  p.B = B; p.C = C; p.S = S; p.save_pre = save_pre; p.cl_pitch = cl_pitch;
  p.out_cl = static_cast<__nv_bfloat16*>(out_cl);
  CUtensorMap tmH, tmS, tmO, tmW;
  const uint64_t rows = static_cast<uint64_t>(B) * C;
  if (make_map_2d(&tmH, h, S, rows, S, 64, C)) return "tensor map (h) failed";
  if (make_map_2d(&tmS, spec_pre, S, rows, S, 64, C)) return "tensor map (spec) failed";
  if (make_map_2d(&tmO, out ? out : spec_pre, S, rows, S, 64, C)) return "tensor map (out) failed";
  if (make_map_2d(&tmW, Wpad, 64, 32, 64, 64, 32)) return "tensor map (W) failed";
  static bool attr = false;
  if (const char* e = set_attr_once(reinterpret_cast<const void*>(bypass_fwd_tc_kernel), &attr)) return e;
  const long long tiles = S / 128 * B;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  const uint32_t smem_bytes = 4096 + kStagesF * 2 * kTile + 1024;
  bypass_fwd_tc_kernel<<<grid, kThreadsB, smem_bytes > 120 * 1024 ? smem_bytes : 120 * 1024, stream>>>(tmH, tmS, tmO, tmW, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* bypass_bwd_tc(const void* dout, const void* dout_cl, int cl_pitch, void* pre_dpre, const void* h,
                          const void* WTpad, void* dhb, float* dW, int B, int C, long long S, int num_sms,
                          cudaStream_t stream) {
  if (C > 32 || S % 128) return "bypass_bwd_tc: need C <= 32 and S % 128 == 0";
  if (S > (1ll << 31) - 256) return "bypass_bwd_tc: slab too large";
  BypassParams p{};
  p.B = B; p.C = C; p.S = S; p.cl_pitch = cl_pitch;
  p.dout_cl = static_cast<const __nv_bfloat16*>(dout_cl);
  p.dW = dW;
  CUtensorMap tmP, tmG, tmH, tmD, tmWT;
  const uint64_t rows = static_cast<uint64_t>(B) * C;
  if (make_map_2d(&tmP, pre_dpre, S, rows, S, 64, C)) return "tensor map (pre) failed";
  if (make_map_2d(&tmG, dout ? dout : pre_dpre, S, rows, S, 64, C)) return "tensor map (dout) failed";
  if (make_map_2d(&tmH, h, S, rows, S, 64, C)) return "tensor map (h) failed";
  if (make_map_2d(&tmD, dhb, S, rows, S, 64, C)) return "tensor map (dhb) failed";
  if (make_map_2d(&tmWT, WTpad, 64, 32, 64, 64, 32)) return "tensor map (WT) failed";
  static bool attr = false;
  if (const char* e = set_attr_once(reinterpret_cast<const void*>(bypass_bwd_tc_kernel), &attr)) return e;
  const long long tiles = S / 128 * B;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  const uint32_t smem_bytes = 4096 + kStagesBw * 3 * kTile + 16384 + 1024;
  bypass_bwd_tc_kernel<<<grid, kThreadsB, smem_bytes > 120 * 1024 ? smem_bytes : 120 * 1024, stream>>>(tmP, tmG, tmH, tmD, tmWT, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

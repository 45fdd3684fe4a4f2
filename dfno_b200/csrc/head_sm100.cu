// head_sm100.cu -- projection head  out = W4 . gelu(W3 h + b3) + b4  (reference linear3 -> gelu -> linear4,
// dfno.py:348-351), forward and backward, straight from the engine's CHANNEL-MAJOR activation
// h[b*C + c][S positions] (no channels-last copy): a tile is 128 consecutive positions of all C channels,
// dropped into shared memory by TMA as two SWIZZLE_128B boxes of [C rows][64 positions] and used as an
// MN-major A operand (M = positions, K = channels).  Row C of every tile image is a constant row of ones and
// column C of the W3 operand holds b3, so the hidden bias rides through the MMA -- and, in the backward, the
// same ones row turns the tensor core into the reducer for db3 and dW4.  The 128-channel hidden layer never
// exists in memory.
//
//   forward   MMA   pre[pos, j] = sum_c h[c, pos] W3[j, c] + b3[j]
//             epi   out[pos]    = b4 + sum_j W4[j] gelu(pre[pos, j])      (packed fp16 GELU, HFMA2 dot)
//
//   backward  MMA1  pre (as above)
//             epi A P[pos, j]   = W4[j] gelu'(pre)   ACT[pos, j] = gelu(pre)      (fp16 tiles in smem)
//                   hs[c, pos]  = s dout[pos] h[c, pos],  hs[C, pos] = s dout[pos]   (fp16, s = 2^k keeps the
//                                 loss gradient inside the fp16 range; undone when the sums are flushed)
//             MMA2  dh0[pos, i]    = sum_j P[pos, j] W3[j, i]          epi B: g[i, pos] = dout[pos] dh0[pos, i]
//             MMA3  D3[j, i]      += sum_pos P[pos, j] hs[i, pos]      -> dW3 (i < C), db3 (i = C)
//             MMA4  D4[j, i]      += sum_pos ACT[pos, j] hs[i, pos]    -> dW4 (i = C)
#include "sm100_ptx.cuh"
#include "kernels.h"
#include "tma_host.h"

namespace dfno {
namespace {

constexpr int kHidH = 128;
constexpr uint32_t kColsHd = 512;

struct RowMap {                      // position row -> element offset in the public [B,1,X,Y,Z,T] layout
  int nrl;
  int R[4];
  long long SR[4];
  unsigned long long Rm[4];
  int Rs[4];
};

__device__ __forceinline__ long long row_to_offset(const RowMap& e, uint32_t r) {
  long long off = 0;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    if (l < e.nrl) {
      uint32_t d = r;
      if (l != e.nrl - 1) {
        const uint32_t q = static_cast<uint32_t>((static_cast<unsigned long long>(r) * e.Rm[l]) >> e.Rs[l]);
        d = r - q * static_cast<uint32_t>(e.R[l]);
        r = q;
      }
      off += static_cast<long long>(d) * e.SR[l];
    }
  }
  return off;
}

void fill_magic(RowMap* m) {
  for (int l = 0; l < 4; ++l) {
    const unsigned d = static_cast<unsigned>(m->R[l] > 0 ? m->R[l] : 1);
    int s = 0;
    while ((1ull << s) < d) ++s;
    m->Rm[l] = ((1ull << (31 + s)) / d) + 1;
    m->Rs[l] = 31 + s;
  }
}

// kind::f16 instruction descriptor with fp16 (not bf16) A and B
__device__ __forceinline__ uint32_t idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return umma_idesc_bf16_f32(M, N, a_mn, b_mn) & ~((7u << 7) | (7u << 10));
}

// ================================================================================ forward
constexpr int kStagesHF = 6;
constexpr int kGroupsHF = 4;
constexpr int kThreadsHF = 64 + 128 * kGroupsHF;

struct HeadFwdParams {
  int B, C, KR;
  long long S, tiles_per_b;
  const float* w4b4;          // [128 weights, 1 bias]
  float* out;
  RowMap map;
};

__global__ void __launch_bounds__(kThreadsHF, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW3,
                const HeadFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w3 = smem;                                    // [128 hid][64] K-major, column C = b3
  uint8_t* s_a = smem + 16384;                             // stages x 2 halves x [KR rows][64 pos]
  const uint32_t half_bytes = static_cast<uint32_t>(p.KR) * 128;
  const uint32_t stage_bytes = 2 * half_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + kStagesHF * stage_bytes);
  uint64_t* full = bars;              // [6]
  uint64_t* empty = bars + 6;         // [6]
  uint64_t* tfull = bars + 12;        // [4]
  uint64_t* tempty = bars + 16;       // [4]
  uint64_t* wfull = bars + 20;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 21);
  uint32_t* s_w4 = reinterpret_cast<uint32_t*>(bars + 24);   // [64] fp16x2 pairs of W4 (16-byte aligned)
  float* s_b4 = reinterpret_cast<float*>(s_w4 + 64);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const long long num_tiles = p.tiles_per_b * p.B;

  for (uint32_t i = threadIdx.x; i < kStagesHF * stage_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kStagesHF * 2 * 16; i += blockDim.x) {      // the ones row (row C) of every half
    const uint32_t hb = i >> 4, ch = i & 15;
    reinterpret_cast<uint2*>(s_a + hb * half_bytes + p.C * 128)[ch] = make_uint2(0x3F803F80u, 0x3F803F80u);
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
    s_w4[i] = h2_bits(h2_from_f32(p.w4b4[2 * i], p.w4b4[2 * i + 1]));
  if (threadIdx.x == 0) s_b4[0] = p.w4b4[kHidH];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmW3);
    for (int s = 0; s < kStagesHF; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 4; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kColsHd>(tmem_holder);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, 16384);
      tma_load_2d(s_w3, &tmW3, wfull, 0, 0);
      uint32_t s = 0, ph = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = static_cast<int>(tile / p.tiles_per_b);
        const int p0 = static_cast<int>((tile % p.tiles_per_b) * 128);
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], 2u * p.C * 128);
        uint8_t* st = s_a + s * stage_bytes;
        tma_load_2d(st, &tmH, &full[s], p0, b * p.C);
        tma_load_2d(st + half_bytes, &tmH, &full[s], p0 + 64, b * p.C);
        if (++s == kStagesHF) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, kHidH, /*A MN-major*/ 1, 0);
    const int ksteps = p.KR >> 4;
    mbar_wait(wfull, 0);
    const uint32_t a16 = (smem_u32(s_a) & 0x3FFFFu) >> 4, w_lo = umma_k128_lo(smem_u32(s_w3));
    const uint32_t a_lbo = ((half_bytes >> 4) & 0x3FFFu) << 16, mn_hi = umma_mn128_hi(1024);
    uint32_t s = 0, ph = 0;
    uint32_t n = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      const uint32_t a = n & 3;
      mbar_wait(&tempty[a], ((n >> 2) & 1) ^ 1);
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      {
        // whole converged warp, warp-uniform operands, one elected lane issues (sm100_ptx.cuh)
        const uint32_t abase = a16 + ((s * stage_bytes) >> 4);
        for (int ks = 0; ks < ksteps; ++ks)
          umma_f16_ss_lohi_warp(tmem_base + a * 128, (abase + ks * 128) | a_lbo, mn_hi, w_lo + ks * 2, kUmmaK128Hi, idesc,
                                ks > 0 ? 1u : 0u);
        umma_commit_warp(&empty[s]);
        umma_commit_warp(&tfull[a]);
      }
      __syncwarp();
      if (++s == kStagesHF) { s = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const float b4 = s_b4[0];
    long long n = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      if ((n & 3) != g) continue;
      const int a = static_cast<int>(n & 3);
      const int b = static_cast<int>(tile / p.tiles_per_b);
      const long long pos = (tile % p.tiles_per_b) * 128 + m;
      mbar_wait(&tfull[a], (n >> 2) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 128;
      float acc = b4;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + ch * 16, v);
        tmem_ld_wait();
        const uint4 wa = reinterpret_cast<const uint4*>(s_w4)[2 * ch], wb = reinterpret_cast<const uint4*>(s_w4)[2 * ch + 1];
        const uint32_t w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        __half2 part = __float2half2_rn(0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          part = __hfma2(gelu_h2(h2_from_f32(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]))),
                         h2_of_bits(w[i]), part);
        const float2 f = __half22float2(part);
        acc += f.x + f.y;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      if (pos < p.S) p.out[row_to_offset(p.map, static_cast<uint32_t>(b * p.S + pos))] = acc;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kColsHd>(tmem_base);
}

// ================================================================================ backward
constexpr int kMaxStagesHB = 6;                 // h tiles are 8 KB: deep prefetch keeps TMA latency off the MMA warp's in-order path
constexpr int kEpiHB = 4;                       // epilogue warps per TMEM lane quarter (one 32-column slice each)
constexpr int kThreadsHB = 64 + 128 * kEpiHB;
constexpr uint32_t kHD1 = 0;                    // 2 x 128 : pre-activations
constexpr uint32_t kHD2 = 256;                  // 2 x 48  : dh tiles
constexpr uint32_t kHD3 = 352;                  // 48      : dW3 / db3 accumulator [hid lanes, c]
constexpr uint32_t kHD4 = 400;                  // 48      : dW4 accumulator (column C)

struct HeadBwdParams {
  int B, C, KR, stages;
  long long S, tiles_per_b;
  const float* dout;          // fp32, public layout
  const float* amax;          // max |dout| (device scalar)
  const float* W4;
  __nv_bfloat16* g;           // [B*C, S]
  float* gW3; float* gb3; float* gW4; float* gb4;
  RowMap map;
};

__global__ void __launch_bounds__(kThreadsHB, 1)
head_bwd2_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW3,
                 const __grid_constant__ CUtensorMap tmW3T, const HeadBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t half_bytes = static_cast<uint32_t>(p.KR) * 128;
  const uint32_t tile_bytes = 2 * half_bytes;
  uint8_t* s_w3 = smem;                                   // 16 KB
  uint8_t* s_w3t = s_w3 + 16384;                          // 2 k-blocks x [KR c rows][64 hid] fp16
  uint8_t* s_p = s_w3t + tile_bytes;                      // 2 buffers x 2 x [128 pos][64 hid] fp16
  uint8_t* s_act = s_p + 65536;
  uint8_t* s_a = s_act + 65536;                           // stages x h tile
  uint8_t* s_hs = s_a + p.stages * tile_bytes;           // 2 x scaled fp16 copy of the h tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_hs + 2 * tile_bytes);
  uint64_t* a_full = bars;            // [8]
  uint64_t* a_empty = bars + 8;       // [8]
  uint64_t* w_full = bars + 16;
  uint64_t* d1_full = bars + 17;      // [2]
  uint64_t* d1_empty = bars + 19;     // [2]
  uint64_t* p_full = bars + 21;       // [2]
  uint64_t* d2_full = bars + 23;      // [2]
  uint64_t* all_done = bars + 25;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 26);
  float* s_gb4 = reinterpret_cast<float*>(bars + 28);
  uint32_t* s_w4h = reinterpret_cast<uint32_t*>(bars + 30);   // [64] fp16x2 pairs of W4 (16-byte aligned)

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const long long num_tiles = p.tiles_per_b * p.B;

  for (uint32_t i = threadIdx.x; i < (p.stages + 2) * tile_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(p.stages) * 2 * 16; i += blockDim.x) {
    const uint32_t hb = i >> 4, ch = i & 15;
    reinterpret_cast<uint2*>(s_a + hb * half_bytes + p.C * 128)[ch] = make_uint2(0x3F803F80u, 0x3F803F80u);
  }
  if (threadIdx.x == 0) s_gb4[0] = 0.f;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_w4h[i] = h2_bits(h2_from_f32(p.W4[2 * i], p.W4[2 * i + 1]));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmW3); tma_prefetch_desc(&tmW3T);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], 2 * kEpiHB);     // one group = 8 warps
      mbar_init(&p_full[i], 2 * kEpiHB); mbar_init(&d2_full[i], 1);
    }
    mbar_init(all_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kColsHd>(tmem_holder);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const float amax = *p.amax;
  const float scale = amax > 0.f ? exp2f(-ceilf(log2f(amax))) : 1.0f;      // |scale * dout| <= 1

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 16384 + tile_bytes);
      tma_load_2d(s_w3, &tmW3, w_full, 0, 0);
      tma_load_2d(s_w3t, &tmW3T, w_full, 0, 0);
      tma_load_2d(s_w3t + half_bytes, &tmW3T, w_full, 64, 0);
      uint32_t s = 0, ph = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = static_cast<int>(tile / p.tiles_per_b);
        const int p0 = static_cast<int>((tile % p.tiles_per_b) * 128);
        mbar_wait(&a_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&a_full[s], 2u * p.C * 128);
        uint8_t* st = s_a + s * tile_bytes;
        tma_load_2d(st, &tmH, &a_full[s], p0, b * p.C);
        tma_load_2d(st + half_bytes, &tmH, &a_full[s], p0 + 64, b * p.C);
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t KR = static_cast<uint32_t>(p.KR);
    const uint32_t idesc1 = umma_idesc_bf16_f32(128, kHidH, 1, 0);      // pre  = h^T-view . W3^T   (bf16)
    const uint32_t idesc2 = idesc_f16(128, KR, 0, 0);                   // dh0  = P . W3T^T         (fp16)
    const uint32_t idesc3 = idesc_f16(128, KR, 1, 0);                   // D3/4 += P^T-view . hs^T  (fp16)
    const int k1steps = p.KR >> 4;
    mbar_wait(w_full, 0);
    const long long nt = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // Two epilogue groups work on alternate tiles (buffer index = tile parity).  MMA1 runs one tile AHEAD of the
    // reductions: MMA1(n+1) only needs the accumulator its own group drained early in tile n-1, so it is issued
    // before part2(n) blocks on that tile's P / ACT / hs -- no group ever waits for its next pre-activations.
    // The whole converged warp runs these loops on warp-uniform operands and one elected lane issues each MMA
    // (sm100_ptx.cuh: umma_f16_ss_lohi_warp).  Inside `if (lane == 0)` every one of the 26 MMAs of a tile was wrapped
    // in an ELECT / R2UR waterfall loop (~120 cycles each): this warp, not the GELU math, set the pace of the kernel.
    const uint32_t a16 = (smem_u32(s_a) & 0x3FFFFu) >> 4, w3_lo = umma_k128_lo(smem_u32(s_w3));
    const uint32_t p16 = (smem_u32(s_p) & 0x3FFFFu) >> 4, act16 = (smem_u32(s_act) & 0x3FFFFu) >> 4;
    const uint32_t hs16 = (smem_u32(s_hs) & 0x3FFFFu) >> 4, w3t_lo = umma_k128_lo(smem_u32(s_w3t));
    const uint32_t a_lbo = ((half_bytes >> 4) & 0x3FFFu) << 16, p_lbo = ((16384u >> 4) & 0x3FFFu) << 16;
    const uint32_t mn_hi = umma_mn128_hi(1024), klbo = 1u << 16;
    const uint32_t tile16 = tile_bytes >> 4, half16 = half_bytes >> 4, n_st = static_cast<uint32_t>(p.stages);
    uint32_t k1 = 0, st1 = 0, ph1 = 0;          // MMA1 tile counter, its ring stage and phase
    uint32_t m2 = 0, st2 = 0;                   // part2 tile counter and ring stage
    auto mma1 = [&]() {
      const uint32_t buf = k1 & 1;
      mbar_wait(&d1_empty[buf], ((k1 >> 1) & 1) ^ 1);
      mbar_wait(&a_full[st1], ph1);
      tcgen05_fence_after();
      const uint32_t abase = a16 + st1 * tile16;
      for (int ks = 0; ks < k1steps; ++ks)
        umma_f16_ss_lohi_warp(tmem_base + kHD1 + buf * 128, (abase + ks * 128) | a_lbo, mn_hi, w3_lo + ks * 2, kUmmaK128Hi,
                              idesc1, ks > 0 ? 1u : 0u);
      umma_commit_warp(&d1_full[buf]);
      ++k1;
      if (++st1 == n_st) { st1 = 0; ph1 ^= 1; }
    };
    auto part2 = [&]() {
      const uint32_t pb = m2 & 1;
      mbar_wait(&p_full[pb], (m2 >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t pbase = p16 + pb * (32768u >> 4);
      const uint32_t abase = act16 + pb * (32768u >> 4);
      const uint32_t hsbase = hs16 + pb * tile16;
      const uint32_t acc0 = m2 > 0 ? 1u : 0u;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {            // K = hid
        const uint32_t kb = ks >> 2, kk = ks & 3;
        umma_f16_ss_lohi_warp(tmem_base + kHD2 + pb * 48, (pbase + kb * (16384u >> 4) + kk * 2) | klbo, kUmmaK128Hi,
                              w3t_lo + kb * half16 + kk * 2, kUmmaK128Hi, idesc2, ks > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {            // K = positions
        const uint32_t kb = ks >> 2, kk = ks & 3;
        const uint32_t b_lo = (hsbase + kb * half16 + kk * 2) | klbo;
        umma_f16_ss_lohi_warp(tmem_base + kHD3, (pbase + ks * 128) | p_lbo, mn_hi, b_lo, kUmmaK128Hi, idesc3,
                              (ks > 0) ? 1u : acc0);
        umma_f16_ss_lohi_warp(tmem_base + kHD4, (abase + ks * 128) | p_lbo, mn_hi, b_lo, kUmmaK128Hi, idesc3,
                              (ks > 0) ? 1u : acc0);
      }
      umma_commit_warp(&d2_full[pb]);
      umma_commit_warp(&a_empty[st2]);
      ++m2;
      if (++st2 == n_st) st2 = 0;
    };
    if (nt > 0) mma1();
    for (long long n = 0; n < nt; ++n) {
      if (n + 1 < nt) mma1();
      part2();
    }
    umma_commit_warp(all_done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int e = (warp - 2) >> 2;               // 0..3
    const int g = e >> 1;                        // epilogue group: tiles of parity g (buffer set g)
    const int sub = e & 1;                       // this warp's 64-column half of the hidden layer
    const int m = q * 32 + lane;                 // position inside the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float acc_gb4 = 0.f;
    // The 8 warps of a group own one tile at a time: warp (q, sub) handles positions 32q.. and hidden units
    // [64 sub, 64 sub + 64); the per-position side work is split between the two warps of a quarter (hs rows of
    // parity sub, dh channels [16 sub, 16 sub + 16)), each loading dout[m] itself (prefetched one tile ahead).
    // The other group works on the neighbouring tile meanwhile, so TMEM / mbarrier / fence latencies of one tile
    // overlap the GELU arithmetic of the other.
    const uint32_t step = 2 * gridDim.x;
    auto load_dout = [&](long long tile) -> float {
      const int b = static_cast<int>(tile / p.tiles_per_b);
      const long long pos = (tile % p.tiles_per_b) * 128 + m;
      return pos < p.S ? p.dout[row_to_offset(p.map, static_cast<uint32_t>(b * p.S + pos))] : 0.f;
    };
    auto drain_dh = [&](long long nn, long long tile, float dout) {
      const int b = static_cast<int>(tile / p.tiles_per_b);
      const long long pos = (tile % p.tiles_per_b) * 128 + m;
      if (16 * sub >= p.C) return;
      mbar_wait(&d2_full[g], (nn >> 1) & 1);
      tcgen05_fence_after();
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + kHD2 + g * 48 + 16 * sub, v);
      tmem_ld_wait();
      if (pos < p.S) {
        __nv_bfloat16* gp = p.g + (static_cast<long long>(b) * p.C + 16 * sub) * p.S + pos;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (16 * sub + i < p.C) gp[static_cast<long long>(i) * p.S] = __float2bfloat16(dout * __uint_as_float(v[i]));
      }
      tcgen05_fence_before();
    };
    const long long first = static_cast<long long>(blockIdx.x) + static_cast<long long>(g) * gridDim.x;
    float dout_next = first < num_tiles ? load_dout(first) : 0.f;
    float dout_prev = 0.f;
    long long tile_prev = 0, n_prev = -1;
    long long n = g;
    for (long long tile = first; tile < num_tiles; tile += step, n += 2) {
      const uint32_t s = static_cast<uint32_t>(n % p.stages);
      const float dout = dout_next;
      if (tile + step < num_tiles) dout_next = load_dout(tile + step);
      if (sub == 0) acc_gb4 += dout;
      mbar_wait(&d1_full[g], (n >> 1) & 1);
      tcgen05_fence_after();
      uint8_t* prow = s_p + g * 32768 + sub * 16384 + m * 128;
      uint8_t* arow = s_act + g * 32768 + sub * 16384 + m * 128;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {                         // two 32-column chunks
        uint32_t v0[16], v1[16];
        const uint32_t t1 = tmem_base + lane_addr + kHD1 + g * 128 + 64 * sub + 32 * hh;
        tmem_ld_32x32b_x16(t1, v0);
        tmem_ld_32x32b_x16(t1 + 16, v1);
        tmem_ld_wait();
        if (hh == 1) {                                         // accumulator is in registers: MMA1 of this group's
          tcgen05_fence_before();                              // next tile may overwrite it
          __syncwarp();
          if (lane == 0) mbar_arrive(&d1_empty[g]);
        }
        uint32_t pw[16], aw[16];
        const uint4* wq = reinterpret_cast<const uint4*>(s_w4h + 32 * sub + 16 * hh);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const uint4 w4 = wq[i4];
          const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = 4 * i4 + u;
            const float x0 = __uint_as_float(i < 8 ? v0[2 * i] : v1[2 * i - 16]);
            const float x1 = __uint_as_float(i < 8 ? v0[2 * i + 1] : v1[2 * i - 15]);
            const GeluH2 vg = gelu_vg_h2(h2_from_f32(x0, x1));
            pw[i] = h2_bits(__hmul2(vg.grad, h2_of_bits(ww[u])));
            aw[i] = h2_bits(vg.value);
          }
        }
        if (hh == 0) {
          // the group's P / ACT / hs buffers were last read by the MMAs of its previous tile (n - 2)
          if (n >= 2) mbar_wait(&d2_full[g], ((n >> 1) - 1) & 1);
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const uint32_t chunk = 4 * hh + 2 * h2;
          const uint32_t o0 = ((chunk ^ (m & 7)) << 4), o1 = (((chunk + 1) ^ (m & 7)) << 4);
          *reinterpret_cast<uint4*>(prow + o0) = make_uint4(pw[8 * h2 + 0], pw[8 * h2 + 1], pw[8 * h2 + 2], pw[8 * h2 + 3]);
          *reinterpret_cast<uint4*>(prow + o1) = make_uint4(pw[8 * h2 + 4], pw[8 * h2 + 5], pw[8 * h2 + 6], pw[8 * h2 + 7]);
          *reinterpret_cast<uint4*>(arow + o0) = make_uint4(aw[8 * h2 + 0], aw[8 * h2 + 1], aw[8 * h2 + 2], aw[8 * h2 + 3]);
          *reinterpret_cast<uint4*>(arow + o1) = make_uint4(aw[8 * h2 + 4], aw[8 * h2 + 5], aw[8 * h2 + 6], aw[8 * h2 + 7]);
        }
      }
      mbar_wait(&a_full[s], (n / p.stages) & 1);              // TMA image of the h tile visible to this thread
      {
        // ---- hs: scaled fp16 copy of this position's column, rows c = sub, sub+2, ...; row C = the gradient itself
        const float ds = dout * scale;
        const uint32_t colo = (m >> 6) * half_bytes + ((m & 7) << 1);
        const uint32_t ch = (m & 63) >> 3;
        const uint8_t* src = s_a + s * tile_bytes;
        uint8_t* dst = s_hs + g * tile_bytes;
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += 8) {
          uint16_t hv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = sub + 2 * (k0 + k);
            hv[k] = c < p.C ? *reinterpret_cast<const uint16_t*>(src + colo + c * 128 + ((ch ^ (c & 7)) << 4)) : 0;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = sub + 2 * (k0 + k);
            if (c < p.C)
              *reinterpret_cast<__half*>(dst + colo + c * 128 + ((ch ^ (c & 7)) << 4)) =
                  __float2half_rn(__uint_as_float(static_cast<uint32_t>(hv[k]) << 16) * ds);
          }
        }
        if (sub == (p.C & 1))
          *reinterpret_cast<__half*>(dst + colo + p.C * 128 + ((ch ^ (p.C & 7)) << 4)) = __float2half_rn(ds);
      }
      // this group's previous tile (its MMAs retired long ago).  Must precede the p_full arrival: the reductions of
      // THIS tile write the same D2 buffer as soon as all eight warps have arrived.
      if (n_prev >= 0) drain_dh(n_prev, tile_prev, dout_prev);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      dout_prev = dout;
      tile_prev = tile;
      n_prev = n;
    }
    if (n_prev >= 0) drain_dh(n_prev, tile_prev, dout_prev);
    n = n_prev + 1;                                            // > 0 iff this group processed a tile
    // ---- per-CTA flush of the weight gradients
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 16);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 8);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 4);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 2);
    acc_gb4 += __shfl_xor_sync(0xffffffffu, acc_gb4, 1);
    if (lane == 0 && sub == 0) atomicAdd(s_gb4, acc_gb4);
    asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiHB) : "memory");
    if (n > 0 && e == 0) {
      mbar_wait(all_done, 0);
      tcgen05_fence_after();
      const float inv = 1.0f / scale;
      const int j = m;                                   // TMEM lane = hidden unit
#pragma unroll
      for (int cb = 0; cb < 3; ++cb) {
        if (cb * 16 < p.KR) {
          uint32_t v3[16], v4[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + kHD3 + cb * 16, v3);
          tmem_ld_32x32b_x16(tmem_base + lane_addr + kHD4 + cb * 16, v4);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = cb * 16 + i;
            if (c < p.C) atomicAdd(p.gW3 + j * p.C + c, __uint_as_float(v3[i]) * inv);
            if (c == p.C) {
              atomicAdd(p.gb3 + j, __uint_as_float(v3[i]) * inv);
              atomicAdd(p.gW4 + j, __uint_as_float(v4[i]) * inv);
            }
          }
        }
      }
      if (j == 0) atomicAdd(p.gb4, s_gb4[0]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kColsHd>(tmem_base);
}

__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out) {
  float mx = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i * 4 < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i * 4 + 3 < n) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    } else {
      for (long long k = i * 4; k < n; ++k) mx = fmaxf(mx, fabsf(x[k]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out, __float_as_uint(mx));   // non-negative floats order like uints
}

int set_rowmap(RowMap* mp, int nrl, const int* R, const long long* SR) {
  if (nrl < 1 || nrl > 4) return -1;
  mp->nrl = nrl;
  for (int i = 0; i < 4; ++i) { mp->R[i] = i < nrl ? R[i] : 1; mp->SR[i] = i < nrl ? SR[i] : 0; }
  fill_magic(mp);
  return 0;
}

}  // namespace

// h: bf16 [B*C, S] channel-major; W3aug: bf16 [128, 64] with column C = b3; w4b4: fp32 [129]; out: fp32, addressed
// through the row digits (row = b*S + position).
const char* head_fwd(const void* h, const void* W3aug, const float* w4b4, float* out, int B, int C, long long S,
                     int nrl, const int* R, const long long* SR, int num_sms, cudaStream_t stream) {
  if (C < 1 || C > 47) return "head_fwd: 1 <= C <= 47";
  if (S % 8 || S > (1ll << 31) - 256 || static_cast<long long>(B) * S > (1ll << 31) - 256) return "head_fwd: bad slab size";
  HeadFwdParams p{};
  p.B = B; p.C = C; p.KR = (C + 1 + 15) / 16 * 16; p.S = S; p.tiles_per_b = (S + 127) / 128;
  p.w4b4 = w4b4; p.out = out;
  if (set_rowmap(&p.map, nrl, R, SR)) return "head_fwd: 1..4 row digits";
  CUtensorMap tmH, tmW3;
  if (make_map_2d(&tmH, h, S, static_cast<uint64_t>(B) * C, S, 64, C)) return "tensor map (h) failed";
  if (make_map_2d(&tmW3, W3aug, 64, 128, 64, 64, 128)) return "tensor map (W3) failed";
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "cudaFuncSetAttribute failed";
    attr = true;
  }
  const uint32_t smem_bytes = 16384 + kStagesHF * 2 * p.KR * 128 + 2048 + 1024;
  const long long tiles = p.tiles_per_b * B;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  head_fwd_kernel<<<grid, kThreadsHF, smem_bytes > 120 * 1024 ? smem_bytes : 120 * 1024, stream>>>(tmH, tmW3, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// W3T16: fp16 [KR, 128] (rows = input channel, zero padded); dout: fp32 public layout; amax_ws: one uint of scratch
// (receives max |dout|); g: bf16 [B*C, S]; gradients are accumulated with atomics.
const char* head_bwd2(const void* h, const void* W3aug, const void* W3T16, const float* W4, const float* dout,
                      long long n_dout, unsigned* amax_ws, void* g, float* gW3, float* gb3, float* gW4, float* gb4,
                      int B, int C, long long S, int nrl, const int* R, const long long* SR, int num_sms,
                      cudaStream_t stream) {
  if (C < 1 || C > 32) return "head_bwd: 1 <= C <= 32";
  if (S % 8 || S > (1ll << 31) - 256 || static_cast<long long>(B) * S > (1ll << 31) - 256) return "head_bwd: bad slab size";
  HeadBwdParams p{};
  p.B = B; p.C = C; p.KR = (C + 1 + 15) / 16 * 16; p.S = S; p.tiles_per_b = (S + 127) / 128;
  p.dout = dout; p.amax = reinterpret_cast<const float*>(amax_ws); p.W4 = W4;
  p.g = static_cast<__nv_bfloat16*>(g); p.gW3 = gW3; p.gb3 = gb3; p.gW4 = gW4; p.gb4 = gb4;
  if (set_rowmap(&p.map, nrl, R, SR)) return "head_bwd: 1..4 row digits";
  CUtensorMap tmH, tmW3, tmW3T;
  if (make_map_2d(&tmH, h, S, static_cast<uint64_t>(B) * C, S, 64, C)) return "tensor map (h) failed";
  if (make_map_2d(&tmW3, W3aug, 64, 128, 64, 64, 128)) return "tensor map (W3) failed";
  if (make_map_2d(&tmW3T, W3T16, 128, p.KR, 128, 64, p.KR)) return "tensor map (W3T) failed";
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(head_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "cudaFuncSetAttribute failed";
    attr = true;
  }
  if (cudaMemsetAsync(amax_ws, 0, 4, stream) != cudaSuccess) return "head_bwd: memset failed";
  absmax_kernel<<<num_sms * 4, 256, 0, stream>>>(dout, n_dout, amax_ws);
  const uint32_t tile_bytes = 2u * p.KR * 128;
  p.stages = kMaxStagesHB;
  while (p.stages > 2 && 16384 + tile_bytes + 131072 + (p.stages + 2) * tile_bytes + 2048 + 1024 > 227 * 1024) --p.stages;
  const uint32_t smem_bytes = 16384 + tile_bytes + 131072 + (p.stages + 2) * tile_bytes + 2048 + 1024;
  const long long tiles = p.tiles_per_b * B;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  head_bwd2_kernel<<<grid, kThreadsHB, smem_bytes, stream>>>(tmH, tmW3, tmW3T, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

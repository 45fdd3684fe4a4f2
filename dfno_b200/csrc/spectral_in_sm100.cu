// spectral_in_sm100.cu -- the FIRST two stages of a Fourier layer in one kernel (SURVEY.md K4 + K5 + K6,
// reference dfno.py:247-259: rfft over the last axis, fft over the next, then the pencil transpose R2).
//
//   Z1[(p, t), (kz, ri)]  = sum_z  h[(p, t), z] . F1[(kz, ri), z]          truncated real z-DFT       (G1a)
//   S1[(kz, p), (kt, ri)] = sum_(t, ri') Z1[(p, t), (kz, ri')] . F2[(kt, ri), (t, ri')]   truncated t-DFT (G1b)
//
// p = one (x, y) position of one (batch, channel) row, Rp positions per tile.  The round-2 chain ran the two
// GEMMs as separate launches with Z1 (0.63 GB per pass at 128^3 x 20) written to and re-read from HBM, and the
// second one -- K = 2T = 40, 80-byte operand rows -- running at 0.44 of copy bandwidth.  Here Z1 never leaves the
// SM:
//
//   MMA1   D1[m1 = p*T + t, n = 2 kz + ri]   A1 = the h tile (TMA, Rp*T lines of Z samples, K-major), B1 = F1
//   epi-1  D1 -> bf16 -> A2[m2 = kz*Rp + p, k = 2 t + ri]   (TMEM -> registers -> swizzled shared memory: the
//          transpose that turns MMA1's rows (t) into MMA2's reduction index)
//   MMA2   D2[m2, n2 = 2 kt + ri] = A2 . F2^T
//   epi-2  D2 -> bf16 pairs -> staging[kz][kt][y] for a run of Yc consecutive y of the same (b, c, x) row
//   flush  one 5-D TMA store per DESTINATION RANK: box (y-run, kt, kz-slab of that rank) straight into the
//          owner's S1 (or its staging block S1s) over NVLink -- the pencil transpose R2 rides on the store.
//
// Warp roles: 0 = TMA producer, 1 = MMA1 issuer, 3 = MMA2 issuer, 2 = store warp, 4.. = E epilogue groups of 4
// warps; group g owns tiles i = g (mod E) and the TMEM / A2 buffers of that slot.
#include "sm100_ptx.cuh"
#include "kernels.h"
#include "tma_host.h"
#include <cstdlib>

namespace dfno {
namespace {

// Probe switches (benchmarks/spin_probe.py, profiles/r2_spin_probe.txt): parts of the kernel can be turned off at run
// time to see what paces it.  Compiled in only with -DDFNO_SPIN_PROBE; the product build has none of the branches.
#ifdef DFNO_SPIN_PROBE
#define DFNO_SPIN_DBG(p) ((p).dbg)
#else
#define DFNO_SPIN_DBG(p) 0
#endif

constexpr int kMaxPeersIn = 8;
constexpr int kMaxE = 4;
constexpr int kMaxStagesIn = 6;

struct alignas(64) PeerMaps {
  CUtensorMap m[kMaxPeersIn];
};

struct SpecInParams {
  long long rows;            // (b, c, x) rows
  int X;
  int T, Rp, RT, RK;         // lines per position, positions per tile, Rp*T, Rp*KZ
  int KZ, mt, kzl, P;
  int Yc, tpc, ncy;          // positions per chunk, tiles per chunk, chunks per row
  int k1blocks, n1_pad;      // operator 1: 64-wide K blocks, padded rows (= MMA1 N)
  int k2blocks, n2_pad, k2steps;
  int stages, E, dbg;
  uint32_t blk1, stage_bytes, a2blk, a2_bytes, stg_bytes, peer_bytes;
};

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2, int32_t c3, int32_t c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

__global__ void __launch_bounds__(128 + 128 * kMaxE, 1)
spectral_in_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ PeerMaps pm,
                   const SpecInParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_b1 = smem;
  uint8_t* s_b2 = s_b1 + static_cast<uint32_t>(p.k1blocks) * p.n1_pad * 128;
  uint8_t* s_ring = s_b2 + static_cast<uint32_t>(p.k2blocks) * p.n2_pad * 128;   // both operator sizes are multiples of 1024
  uint8_t* s_a2 = s_ring + p.stages * p.stage_bytes;
  uint8_t* s_stg = s_a2 + p.E * p.a2_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stg + 2 * p.stg_bytes);
  uint64_t* full = bars;                       // [kMaxStagesIn] TMA -> MMA
  uint64_t* empty = full + kMaxStagesIn;       // [kMaxStagesIn] MMA -> TMA
  uint64_t* d1_full = empty + kMaxStagesIn;    // [kMaxE] MMA1 -> epilogue
  uint64_t* d1_empty = d1_full + kMaxE;        // [kMaxE] epilogue -> MMA1
  uint64_t* a2_full = d1_empty + kMaxE;        // [kMaxE] epilogue -> MMA2
  uint64_t* d2_full = a2_full + kMaxE;         // [kMaxE] MMA2 -> epilogue
  uint64_t* stg_done = d2_full + kMaxE;        // [2] epilogue -> store warp
  uint64_t* stg_free = stg_done + 2;           // [2] store warp -> epilogue
  uint64_t* bfull = stg_free + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bfull + 1);

  // broadcast through a shuffle so that the compiler knows the role index is warp-uniform (uniform datapath, no R2UR)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const long long n_chunks = p.rows * p.ncy;
  const long long my_chunks = blockIdx.x < n_chunks ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long my_tiles = my_chunks * p.tpc;

  // the K padding of A2 (columns 2T .. 16*k2steps) is never written by the epilogue: zero the buffers once
  {
    uint4* z0 = reinterpret_cast<uint4*>(s_a2);
    const uint32_t nz = p.E * p.a2_bytes / 16;
    for (uint32_t i = threadIdx.x; i < nz; i += blockDim.x) z0[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2);
    for (int j = 0; j < p.P; ++j) tma_prefetch_desc(&pm.m[j]);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int g = 0; g < p.E; ++g) {
      mbar_init(&d1_full[g], 1); mbar_init(&d1_empty[g], 1); mbar_init(&a2_full[g], 1); mbar_init(&d2_full[g], 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&stg_done[b], p.tpc); mbar_init(&stg_free[b], 1); }
    mbar_init(bfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t gcols = (DFNO_SPIN_DBG(p) & 128) ? 128u : static_cast<uint32_t>(p.n1_pad + p.n2_pad);   // TMEM columns per group: [D1 | D2]
  const uint32_t d2col = (DFNO_SPIN_DBG(p) & 128) ? 64u : static_cast<uint32_t>(p.n1_pad);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bfull, (static_cast<uint32_t>(p.k1blocks) * p.n1_pad + static_cast<uint32_t>(p.k2blocks) * p.n2_pad) * 128);
      for (int kb = 0; kb < p.k1blocks; ++kb) tma_load_2d(s_b1 + kb * p.n1_pad * 128, &tmB1, bfull, kb * 64, 0);
      for (int kb = 0; kb < p.k2blocks; ++kb) tma_load_2d(s_b2 + kb * p.n2_pad * 128, &tmB2, bfull, kb * 64, 0);
      uint32_t s = 0, ph = 0;
      for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int row = static_cast<int>(chunk / p.ncy);
        const int cy = static_cast<int>(chunk - static_cast<long long>(row) * p.ncy);
        for (int tt = 0; tt < p.tpc; ++tt) {
          const int line0 = (cy * p.Yc + tt * p.Rp) * p.T;
          mbar_wait(&empty[s], ph ^ 1);
          if (DFNO_SPIN_DBG(p) & 16) { mbar_arrive(&full[s]); if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1; } continue; }
          mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(p.k1blocks) * p.RT * 128);
          uint8_t* dst = s_ring + s * p.stage_bytes;
          for (int kb = 0; kb < p.k1blocks; ++kb) tma_load_3d(dst + kb * p.blk1, &tmH, &full[s], kb * 64, line0, row);
          if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA1 issuer =====================
    // The whole (converged) warp runs the issue loop with warp-uniform operands and one elected lane issues (see
    // umma_bf16_ss_k128_warp).  MMA2 has its own issuing warp, so neither chain waits behind the other's barriers.
    // (ONE warp for all MMA1 tiles: with two alternating warps a warp could reach its next use of a slot two
    // phases ahead of the barrier and pass a parity wait spuriously -- seen as a hang with tiny tiles.)
    mbar_wait(bfull, 0);
    const uint32_t idesc1 = umma_idesc_bf16_f32(128, static_cast<uint32_t>(p.n1_pad));
    const uint32_t stage16 = p.stage_bytes >> 4, blk1_16 = p.blk1 >> 4, b1blk16 = (p.n1_pad * 128) >> 4;
    const uint32_t dA1 = umma_k128_lo(smem_u32(s_ring)), dB1 = umma_k128_lo(smem_u32(s_b1));
    const uint32_t E = p.E, ST = p.stages, K1B = p.k1blocks;
    // tile j of this CTA: TMEM / A2 slot g1 = j mod E (phase u1), ring stage s1 = j mod ST (phase ph1)
    uint32_t g1 = 0, s1 = 0, u1 = 0, ph1 = 0;
    for (long long j = 0; j < my_tiles; ++j) {
      mbar_wait(&d1_empty[g1], u1 ^ 1);
      mbar_wait(&full[s1], ph1);
      tcgen05_fence_after();
      const uint32_t d = tmem_base + g1 * gcols;
      uint32_t da = dA1 + s1 * stage16, db = dB1;
      for (uint32_t kb = 0; kb < K1B; ++kb) {
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk) umma_bf16_ss_k128_warp(d, da + 2 * kk, db + 2 * kk, idesc1, (kb | kk) ? 1u : 0u);
        da += blk1_16;
        db += b1blk16;
      }
      umma_commit_warp(&empty[s1]);
      umma_commit_warp(&d1_full[g1]);
      if (++g1 == E) { g1 = 0; u1 ^= 1; }
      if (++s1 == ST) { s1 = 0; ph1 ^= 1; }
    }
  } else if (warp == 3) {
    // ===================== MMA2 issuer =====================
    mbar_wait(bfull, 0);
    const uint32_t idesc2 = umma_idesc_bf16_f32(128, static_cast<uint32_t>(p.n2_pad));
    const uint32_t a2b16 = p.a2_bytes >> 4, a2blk16 = p.a2blk >> 4, b2blk16 = (p.n2_pad * 128) >> 4;
    const uint32_t dA2 = umma_k128_lo(smem_u32(s_a2)), dB2 = umma_k128_lo(smem_u32(s_b2));
    const uint32_t E = p.E, K2S = p.k2steps;
    uint32_t g2 = 0, u2 = 0;
    for (long long i = 0; i < my_tiles; ++i) {
      mbar_wait(&a2_full[g2], u2);
      tcgen05_fence_after();
      const uint32_t d = tmem_base + g2 * gcols + d2col;
      const uint32_t da = dA2 + g2 * a2b16;
      for (uint32_t ks = 0; ks < K2S; ++ks) {
        const uint32_t kb = ks >> 2, kk = ks & 3;
        umma_bf16_ss_k128_warp(d, da + kb * a2blk16 + 2 * kk, dB2 + kb * b2blk16 + 2 * kk, idesc2, ks > 0 ? 1u : 0u);
      }
      umma_commit_warp(&d2_full[g2]);
      if (++g2 == E) { g2 = 0; u2 ^= 1; }
    }
  } else if (warp == 2) {
    // ===================== store warp: one TMA store per destination rank and chunk =====================
    if (lane == 0) {
      uint32_t cn = 0;
      for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++cn) {
        const int row = static_cast<int>(chunk / p.ncy);
        const int cy = static_cast<int>(chunk - static_cast<long long>(row) * p.ncy);
        const int bc = row / p.X, x = row - bc * p.X;
        const uint32_t b = cn & 1;
        mbar_wait(&stg_done[b], (cn >> 1) & 1);
        const uint8_t* src = s_stg + b * p.stg_bytes;
        if (!(DFNO_SPIN_DBG(p) & 4)) for (int j = 0; j < p.P; ++j) tma_store_5d(&pm.m[j], src + j * p.peer_bytes, cy * p.Yc * 2, x, 0, 0, bc);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&stg_free[b]);
      }
      tma_store_wait_all();
      __threadfence_system();
    }
  } else if (warp >= 4 && warp < 4 + 4 * p.E) {
    // ===================== epilogue groups =====================
    const int q = warp & 3, g = (warp - 4) >> 2;
    const int m = q * 32 + lane;                         // TMEM lane = row of D1 and of D2
    const bool elected = ((warp - 4) & 3) == 0 && lane == 0;
    const uint32_t barid = 1 + g;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * gcols;
    // epi-1: row m = p1*T + t  ->  A2[kz*Rp + p1, 2t .. 2t+1]
    const bool act1 = m < p.RT, warp1 = q * 32 < p.RT;
    const int p1 = m / p.T, t1 = m - p1 * p.T;
    const uint32_t k0 = 2 * t1;
    const uint32_t koff = (k0 >> 6) * p.a2blk, c16 = (k0 & 63) >> 3, wi = ((k0 & 7) >> 1) * 4;
    uint8_t* a2 = s_a2 + g * p.a2_bytes;
    // epi-2: row m = kz*Rp + p2  ->  staging[(kz*mt + kt)*Yc + tt*Rp + p2]
    const bool act2 = m < p.RK, warp2 = q * 32 < p.RK;
    const int kz2 = m / p.Rp, p2 = m - kz2 * p.Rp;
    const int n1 = 2 * p.KZ, n2 = 2 * p.mt;
    int gi = 0;                                           // group of the current tile (tiles rotate over the groups)
    uint32_t u = 0, cn = 0;                               // phase parity of this group's barriers; chunk counter
    // one lane polls, the warp follows: 128 threads spinning on try_wait slow every other mbarrier operation down
    auto wait_warp = [&](uint64_t* bar, uint32_t parity) {
      if (lane == 0) { if (DFNO_SPIN_DBG(p) & 256) mbar_spin(bar, parity); else mbar_wait(bar, parity); }
      __syncwarp();
    };
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++cn) {
      uint32_t* stg = reinterpret_cast<uint32_t*>(s_stg + (cn & 1) * p.stg_bytes);
      for (int tt = 0; tt < p.tpc; ++tt) {
        const bool mine = gi == g;
        if (++gi == p.E) gi = 0;
        if (!mine) continue;
        // ---------------- epi-1 ----------------
        wait_warp(&d1_full[g], u);
        tcgen05_fence_after();
        if (warp1 && !(DFNO_SPIN_DBG(p) & 1)) {
          for (int c0 = 0; c0 < n1; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + c0, v);
            tmem_ld_wait();
            if (act1) {
#pragma unroll
              for (int jz = 0; jz < 8; ++jz) {
                const int kz = (c0 >> 1) + jz;
                if (2 * kz < n1) {
                  const uint32_t r = static_cast<uint32_t>(kz * p.Rp + p1);
                  *reinterpret_cast<uint32_t*>(a2 + koff + r * 128 + (((c16 ^ (r & 7)) << 4) | wi)) =
                      pack_bf16x2(__uint_as_float(v[2 * jz]), __uint_as_float(v[2 * jz + 1]));
                }
              }
            }
          }
        }
        tcgen05_fence_before();
        if (!(DFNO_SPIN_DBG(p) & 8)) fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
        if (elected) { mbar_arrive(&d1_empty[g]); mbar_arrive(&a2_full[g]); }
        // ---------------- epi-2 ----------------
        wait_warp(&d2_full[g], u);
        tcgen05_fence_after();
        wait_warp(&stg_free[cn & 1], ((cn >> 1) & 1) ^ 1);
        u ^= 1;
        if (warp2 && !(DFNO_SPIN_DBG(p) & 2)) {
          for (int c0 = 0; c0 < n2; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + d2col + c0, v);
            tmem_ld_wait();
            if (act2) {
#pragma unroll
              for (int jt = 0; jt < 8; ++jt) {
                const int kt = (c0 >> 1) + jt;
                if (kt < p.mt)
                  stg[(kz2 * p.mt + kt) * p.Yc + tt * p.Rp + p2] =
                      pack_bf16x2(__uint_as_float(v[2 * jt]), __uint_as_float(v[2 * jt + 1]));
              }
            }
          }
        }
        tcgen05_fence_before();
        if (!(DFNO_SPIN_DBG(p) & 8)) fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
        if (elected) mbar_arrive(&stg_done[cn & 1]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

inline uint32_t align_up_u32(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// Tile configuration for one problem shape (shared by the launcher and by the eligibility query).
const char* plan_spectral_in(SpecInParams& p, int n1_pad, int k1_pad, int n2_pad, int k2_pad, int P, long long dst_off,
                             const long long* dstr, int BC, int X, int Yl, int T, int Z, int KZ, int mt) {
  if (P < 1 || P > kMaxPeersIn || KZ % P) return "spectral_in: 1..8 destination ranks, KZ divisible by their number";
  if (Z % 8 || Z > 256 || k1_pad % 64 || k1_pad < Z || k1_pad > 256) return "spectral_in: need Z % 8 == 0, Z <= 256";
  if (T < 1 || T > 64 || k2_pad % 64 || k2_pad < 2 * T) return "spectral_in: need T <= 64";
  if (n1_pad % 16 || n1_pad < 2 * KZ || n1_pad > 128 || n2_pad % 16 || n2_pad < 2 * mt || n2_pad > 128)
    return "spectral_in: operator padding";
  if (KZ > 128 || mt < 1) return "spectral_in: mode counts";
  if (Yl % 4) return "spectral_in: the local y extent must be a multiple of 4 (stores are clipped in 16-byte units)";
  if (dst_off % 8 || dstr[0] % 8 || dstr[1] % 8 || dstr[2] % 8 || dstr[3] % 8)
    return "spectral_in: destination offset / strides must be multiples of 8 elements (16-byte TMA alignment)";
  p = SpecInParams{};
  p.rows = static_cast<long long>(BC) * X;
  if (p.rows > (1ll << 30)) return "spectral_in: tensor too large";
  p.X = X; p.T = T; p.KZ = KZ; p.mt = mt; p.P = P; p.kzl = KZ / P;
  p.k1blocks = k1_pad / 64; p.n1_pad = n1_pad;
  p.k2steps = (2 * T + 15) / 16; p.k2blocks = (p.k2steps + 3) / 4; p.n2_pad = n2_pad;
  if (p.k2blocks * 64 > k2_pad) return "spectral_in: operator 2 is narrower than its reduction";
  const uint32_t ops_bytes = static_cast<uint32_t>(p.k1blocks) * n1_pad * 128 + static_cast<uint32_t>(p.k2blocks) * n2_pad * 128;
  if ((static_cast<uint32_t>(p.k1blocks) * n1_pad * 128) % 1024 || ops_bytes % 1024) return "spectral_in: operator rows must be a multiple of 8";
  const uint32_t budget = 227 * 1024 - 1024 /*align*/ - 512 /*barriers*/;
  const int rmax = (128 / T) < (128 / KZ) ? (128 / T) : (128 / KZ);
  bool ok = false;
  for (int min_st = 3; min_st >= 2 && !ok; --min_st)       // prefer a deep TMA ring and two epilogue groups
  for (int Rp = 4; Rp >= 1 && !ok; Rp >>= 1) {
    if (Rp > rmax) continue;
    int yc = 32;
    while (yc > 4 && yc / 2 >= Yl) yc >>= 1;                   // the smallest of {4, 8, 16, 32} covering Yl, 32 beyond
    for (; yc >= 4 && !ok; yc >>= 1) {
      if (yc % Rp) continue;
      const uint32_t peer_bytes = static_cast<uint32_t>(p.kzl) * mt * yc * 4;
      if (peer_bytes % 128) continue;
      const uint32_t stg_bytes = align_up_u32(peer_bytes * P, 1024);
      const uint32_t blk1 = align_up_u32(static_cast<uint32_t>(Rp) * T * 128, 1024);
      const uint32_t a2blk = align_up_u32(static_cast<uint32_t>(Rp) * KZ * 128, 1024);
      const uint32_t n_g = static_cast<uint32_t>(n1_pad + n2_pad);
      static const int tryE[] = {4, 3, 2, 1};
      for (int ie = 0; ie < 4 && !ok; ++ie) {
        const int E = tryE[ie];
        if (E * n_g > 512) continue;
        for (int st = 5; st >= 2; --st) {
          const uint32_t need = ops_bytes + st * p.k1blocks * blk1 + E * p.k2blocks * a2blk + 2 * stg_bytes;
          if (need <= budget && st >= min_st && (E >= 2 || min_st == 2)) {
            p.Rp = Rp; p.Yc = yc; p.E = E; p.stages = st; p.blk1 = blk1; p.a2blk = a2blk;
            p.stage_bytes = p.k1blocks * blk1; p.a2_bytes = p.k2blocks * a2blk; p.stg_bytes = stg_bytes;
            p.peer_bytes = peer_bytes;
            ok = true;
            break;
          }
        }
      }
    }
  }
  if (!ok) return "spectral_in: no tile configuration fits shared memory";
  p.RT = p.Rp * T; p.RK = p.Rp * KZ;
  p.tpc = p.Yc / p.Rp; p.ncy = (Yl + p.Yc - 1) / p.Yc;
  return nullptr;
}

}  // namespace

// nullptr when spectral_in supports the shape (no launch): the engine asks before it drops G1a + G1b from its chain.
const char* spectral_in_check(int n1_pad, int k1_pad, int n2_pad, int k2_pad, int P, long long dst_off, const long long* dstr,
                              int BC, int X, int Yl, int T, int Z, int KZ, int mt, int* cfg) {
  SpecInParams p;
  const char* e = plan_spectral_in(p, n1_pad, k1_pad, n2_pad, k2_pad, P, dst_off, dstr, BC, X, Yl, T, Z, KZ, mt);
  if (!e && cfg) { cfg[0] = p.Rp; cfg[1] = p.Yc; cfg[2] = p.E; cfg[3] = p.stages; }
  return e;
}

// h: bf16 [rows = B*C*X, Yl, T, Z] (the engine layout of one activation).  op1: padded bf16 [n1_pad >= 2 KZ, k1_pad >= Z],
// op2: padded bf16 [n2_pad >= 2 mt, k2_pad >= 2 T] (reduction index 2 t + ri).  Destination: for rank j the bf16
// tensor at dst_ptrs[j] + dst_off viewed as [B*C, kzl, mt, X, Yl*2] with element strides dstr = {x, kt, kz, bc}
// (the y / (re, im) run is contiguous); rank j receives the modes kz in [j*kzl, (j+1)*kzl).
const char* spectral_in(const void* h, const void* op1, int n1_pad, int k1_pad, const void* op2, int n2_pad, int k2_pad,
                        const long long* dst_ptrs, int P, long long dst_off, const long long* dstr, int BC, int X,
                        int Yl, int T, int Z, int KZ, int mt, int num_sms, cudaStream_t stream) {
  SpecInParams p;
  if (const char* err = plan_spectral_in(p, n1_pad, k1_pad, n2_pad, k2_pad, P, dst_off, dstr, BC, X, Yl, T, Z, KZ, mt)) return err;
#ifdef DFNO_SPIN_PROBE
  if (const char* e = getenv("DFNO_SPIN_DBG")) p.dbg = atoi(e);
  if (const char* e = getenv("DFNO_SPIN_E")) { const int v = atoi(e); if (v >= 1 && v <= 3 && v * (n1_pad + n2_pad) <= 512) p.E = v; }
  if (const char* e = getenv("DFNO_SPIN_ST")) { const int v = atoi(e); if (v >= 2 && v <= kMaxStagesIn) p.stages = v; }
#endif
  const uint32_t ops_bytes = static_cast<uint32_t>(p.k1blocks) * n1_pad * 128 + static_cast<uint32_t>(p.k2blocks) * n2_pad * 128;
  CUtensorMap tmH, tmB1, tmB2;
  PeerMaps pm;
  if (make_map_3d(&tmH, h, Z, static_cast<uint64_t>(Yl) * T, p.rows, Z, static_cast<uint64_t>(Yl) * T * Z, 64, p.RT, 1))
    return "tensor map (h) failed";
  if (make_map_2d(&tmB1, op1, k1_pad, n1_pad, k1_pad, 64, n1_pad)) return "tensor map (operator 1) failed";
  if (make_map_2d(&tmB2, op2, k2_pad, n2_pad, k2_pad, 64, n2_pad)) return "tensor map (operator 2) failed";
  for (int j = 0; j < kMaxPeersIn; ++j) {
    const int jj = j < P ? j : 0;
    const uint64_t dims[5] = {static_cast<uint64_t>(Yl) * 2, static_cast<uint64_t>(X), static_cast<uint64_t>(mt),
                              static_cast<uint64_t>(p.kzl), static_cast<uint64_t>(BC)};
    const uint64_t str[4] = {static_cast<uint64_t>(dstr[0]), static_cast<uint64_t>(dstr[1]), static_cast<uint64_t>(dstr[2]),
                             static_cast<uint64_t>(dstr[3])};
    const uint32_t box[5] = {static_cast<uint32_t>(p.Yc) * 2, 1, static_cast<uint32_t>(mt), static_cast<uint32_t>(p.kzl), 1};
    const void* base = reinterpret_cast<const void*>(static_cast<uintptr_t>(dst_ptrs[jj]) + static_cast<uintptr_t>(dst_off) * 2);
    if (make_map_nd_plain(&pm.m[j], base, 5, dims, str, box)) return "tensor map (destination) failed";
  }
  const uint32_t smem_bytes = ops_bytes + p.stages * p.stage_bytes + p.E * p.a2_bytes + 2 * p.stg_bytes + 512 + 1024;
  if (smem_bytes > 227 * 1024) return "spectral_in: probe override does not fit shared memory";
  const long long chunks = p.rows * p.ncy;
  const int grid = static_cast<int>(chunks < num_sms ? chunks : num_sms);
  const int threads = 128 + 128 * p.E;
  const uint32_t dyn = smem_bytes > 120 * 1024 ? smem_bytes : 120 * 1024;      // one CTA per SM (512 TMEM columns each)
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(spectral_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "cudaFuncSetAttribute failed";
    attr = true;
  }
  spectral_in_kernel<<<grid, threads, dyn, stream>>>(tmH, tmB1, tmB2, pm, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

// dft_gemm_sm100.cu -- "skinny" GEMM with a resident operator matrix on tcgen05 / TMEM / TMA.
//
//     C[M, N] = A[M, K] * B[N, K]^T        A, B bf16 (K-major), fp32 accumulation in TMEM
//
// This is the workhorse of the Fourier layers: every truncated (inverse) DFT stage of
// SURVEY.md §2.5 (K4, K5, K7, K10, K12, K13 and their adjoints) is such a product with
// M = millions of "lines" of the field, K = the transformed axis (x interleaved re/im),
// N = the retained modes (x re/im) -- i.e. a tiny operator B applied to a huge streamed A.
// The design follows from that shape, not from a square-GEMM template:
//
//   * B (<= 256 x 256 bf16) is loaded ONCE per CTA by TMA and stays resident in shared memory.
//   * the kernel is persistent (one CTA per SM); A tiles of 128 rows x K stream through a
//     multi-stage TMA/mbarrier ring (SWIZZLE_128B, 64-element K blocks);
//   * one thread issues tcgen05.mma (M=128, N=N_pad, K=16) into a double-buffered TMEM
//     accumulator, so the epilogue of tile i overlaps the loads+MMA of tile i+1;
//   * the epilogue (4 warps = the 4 TMEM lane quarters) reads the accumulator with
//     tcgen05.ld and writes either a row-major tile (optionally adding a bf16 tensor) or a
//     *scattered, transposed* layout addressed through a small mixed-radix table.  The
//     scatter target may live on another GPU: the base pointer is selected per peer from a
//     table of NVLink-mapped symmetric buffers, which is how the pencil-transpose
//     all-to-all (Repartition R2/R3) is fused into the producing GEMM:
//     the tile goes straight from TMEM to its owner's memory while the tensor core works on
//     the next tile.  No NCCL call, no pack/unpack pass.
//
// Work is memory bound by construction (AI ~ N flop/byte); the tensor core only has to
// stay off the critical path, which is why one CTA per SM with M=128 UMMA is sufficient.
#include "sm100_ptx.cuh"
#include "dft_gemm.h"
#include "tma_host.h"

namespace dfno {

static constexpr int kTileM = 128;
static constexpr int kBlockK = 64;                 // bf16 elements per 128-byte swizzle row
static constexpr int kMaxThreads = 64 + 128 * 4;   // warp0 TMA, warp1 MMA, then E groups of 4 epilogue warps
static constexpr uint32_t kTmemCols = 512;
static constexpr int kMaxStages = 8;                // A-tile ring depth: bytes in flight per SM must cover
                                                    // HBM latency x bandwidth share (~45 KB), small tiles need more stages
static constexpr int kMaxAcc = 8;                  // TMEM accumulator stages

struct SmemLayout {
  uint32_t b_bytes;       // kblocks * n_pad * 128
  uint32_t a_tile_bytes;  // bytes of one ring stage = kbs * 16384
  uint32_t kbs;           // 64-wide K blocks per ring stage (= kblocks unless the whole-K tile is too large)
  uint32_t stages;
  uint32_t nacc;          // TMEM accumulator stages (multiple of the number of epilogue groups)
  uint32_t stage_off;     // byte offset of the epilogue staging area
  uint32_t stage_pitch;   // bytes per staged fp32 row (+16 B pad); 0 = direct stores
};

// floor(n / d) for n < 2^31 with a host-computed magic number
__device__ __forceinline__ uint32_t fast_div(uint32_t n, unsigned long long magic, int shift) {
  return static_cast<uint32_t>((static_cast<unsigned long long>(n) * magic) >> shift);
}

// mixed-radix row address (shared by the scatter and head epilogues)
__device__ __forceinline__ long long row_offset(const EpiParams& e, uint32_t r, int& peer) {
  long long off = e.base_off;
  peer = 0;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    if (l < e.nrl) {
      uint32_t d = r;
      if (l != e.nrl - 1) {
        const uint32_t q = fast_div(r, e.Rm[l], e.Rs[l]);
        d = r - q * static_cast<uint32_t>(e.R[l]);
        r = q;
      }
      if (e.peer_sel == PEER_BY_ROW && l == e.peer_lvl) {
        const uint32_t pq = fast_div(d, e.Pm, e.Ps);
        peer = static_cast<int>(pq);
        d -= pq * static_cast<uint32_t>(e.peer_div);
      }
      off += static_cast<long long>(d) * e.SR[l];
    }
  }
  return off;
}

__global__ void __launch_bounds__(kMaxThreads, 1)
dft_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmParams p, const SmemLayout L) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + L.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + L.stages * L.a_tile_bytes);
  uint64_t* full = bars;                      // [8]   TMA -> MMA
  uint64_t* empty = bars + 8;                 // [8]   MMA -> TMA
  uint64_t* tfull = bars + 16;                // [8]   MMA -> epilogue
  uint64_t* tempty = bars + 24;               // [8]   epilogue -> MMA
  uint64_t* bfull = bars + 32;                // [1]   B resident
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 33);
  long long* s_coloff = reinterpret_cast<long long*>(bars + 40);        // [128] pair -> element offset
  float* s_vec = reinterpret_cast<float*>(s_coloff + 128);              // [512] EPI_HEAD vectors
  uint8_t* s_colpeer = reinterpret_cast<uint8_t*>(s_vec + 512);         // [128] pair -> peer

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int kblocks = p.k_pad / kBlockK;
  const int num_tiles = (p.M + kTileM - 1) / kTileM;
  const int E = (static_cast<int>(blockDim.x) - 64) >> 7;      // epilogue groups (4 warps each)
  const int nacc = static_cast<int>(L.nacc);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (uint32_t s = 0; s < L.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < nacc; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);               // the four warps of the group that drains this stage
    }
    mbar_init(bfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_holder);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bfull, L.b_bytes);
      for (int kb = 0; kb < kblocks; ++kb)
        tma_load_2d(smem_b + kb * p.n_pad * 128, &tmB, bfull, kb * kBlockK, 0);
      uint32_t s = 0, ph = 0;
      const int kbs = static_cast<int>(L.kbs);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb0 = 0; kb0 < kblocks; kb0 += kbs) {          // one ring stage per K chunk (usually the whole K)
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], L.a_tile_bytes);
          uint8_t* dst = smem_a + s * L.a_tile_bytes;
          for (int kb = 0; kb < kbs; ++kb)
            tma_load_2d(dst + kb * (kTileM * 128), &tmA, &full[s], (kb0 + kb) * kBlockK, tile * kTileM);
          if (++s == L.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // With the skinny stages (K = 20..48: 3 MMAs per 128-row tile, 60 k tiles) this warp's work per tile -- not
    // the tensor core, not HBM -- set the pace of the kernel (profiles/r2_ncu_G1b.json: TMA and epilogue both
    // waiting on it), so the loop carries no integer division (stage / phase counters are incremental).
    // The WHOLE converged warp runs the loop on warp-uniform operands and one elected lane issues each instruction
    // (umma_bf16_ss_k128_warp): with the loop inside `if (lane == 0)` the compiler kept the descriptors in vector
    // registers and wrapped every UTCHMMA in an ELECT / R2UR waterfall loop, ~120 cycles per MMA on this warp.
    mbar_wait(bfull, 0);
    {
      // a_format lives in bits [7,10): 1 = BF16, 0 = F16
      const uint32_t idesc = umma_idesc_bf16_f32(kTileM, p.n_pad) & ~(p.a_f16 ? (7u << 7) : 0u);
      const int ksteps = (p.K + 15) / 16;       // K=16 per instruction; zero tail needs no MMA
      const int kbs = static_cast<int>(L.kbs);
      const uint32_t bdesc0 = umma_k128_lo(smem_u32(smem_b));
      const uint32_t adesc0 = umma_k128_lo(smem_u32(smem_a));
      const uint32_t a_stage16 = L.a_tile_bytes >> 4;              // descriptor address units are 16 bytes
      const uint32_t b_kb16 = static_cast<uint32_t>(p.n_pad) * 128 >> 4;
      const uint32_t n_stages = L.stages;
      uint32_t s = 0, ph = 0, a = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[a], aph ^ 1);
        const uint32_t d_tmem = tmem_base + a * p.n_pad;
        for (int kb0 = 0; kb0 < kblocks; kb0 += kbs) {
          mbar_wait(&full[s], ph);
          tcgen05_fence_after();
          const uint32_t adesc_s = adesc0 + s * a_stage16;
          const int ks_end = min(ksteps, (kb0 + kbs) * 4);
          for (int ks = kb0 * 4; ks < ks_end; ++ks) {
            const uint32_t kb = ks >> 2, kk = ks & 3;
            umma_bf16_ss_k128_warp(d_tmem, adesc_s + ((kb - kb0) * (kTileM * 128 >> 4) + kk * 2),
                                   bdesc0 + (kb * b_kb16 + kk * 2), idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit_warp(&empty[s]);        // smem stage may be refilled once the MMAs retire
          if (kb0 + kbs >= kblocks) umma_commit_warp(&tfull[a]);   // accumulator ready for the epilogue
          if (++s == n_stages) { s = 0; ph ^= 1; }
        }
        if (++a == static_cast<uint32_t>(nacc)) { a = 0; aph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> global / peer memory ==========
    // E groups of 4 warps (one per TMEM lane quarter; hardware: warp w may only touch lanes
    // 32*(w%4)..+31).  Group g drains the tiles n = g, g+E, ... of this CTA, so E tile
    // epilogues are in flight and their latency chains overlap.
    const int q = warp & 3;                                      // TMEM lane quarter of this warp
    const int g = (warp - 2) >> 2;                               // epilogue group
    const int r_in_tile = q * 32 + lane;
    const int npairs = p.N >> 1;
    {
      // per-CTA lookup tables (all epilogue warps; named barrier 1)
      const int et = threadIdx.x - 64;
      const int nthr = 128 * E;
      if (p.epi.mode == EPI_PAIR_SCATTER) {
        for (int j = et; j < npairs && j < 128; j += nthr) {
          int jj = j, peer = 0;
          if (p.epi.peer_sel == PEER_BY_COL) { peer = jj / p.epi.peer_div; jj -= peer * p.epi.peer_div; }
          const int j0 = jj % p.epi.J[0], j1 = jj / p.epi.J[0];
          s_coloff[j] = j0 * p.epi.SJ[0] + j1 * p.epi.SJ[1];
          s_colpeer[j] = static_cast<uint8_t>(peer);
        }
      } else if (p.epi.mode == EPI_HEAD) {
        for (int j = et; j < p.N; j += nthr) { s_vec[j] = p.epi.v0[j]; s_vec[256 + j] = p.epi.v1[j]; }
        if (et == 0) s_vec[511] = p.epi.v1[p.N];          // output bias stored right after the weights
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
    }
    // this group drains the CTA's tiles g, g+E, ...: accumulator stage and phase advance without divisions
    // (nacc is a multiple of E)
    int a = g;
    uint32_t aph = 0;
    for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles;
         tile += E * gridDim.x, a += E, aph ^= (a >= nacc ? 1u : 0u), a -= (a >= nacc ? nacc : 0)) {
      const long long row = static_cast<long long>(tile) * kTileM + r_in_tile;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull[a], aph);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * p.n_pad;

      if (p.epi.mode == EPI_ROWMAJOR && L.stage_pitch != 0) {
        // ---- coalesced row-major store: TMEM -> fp32 staging rows in smem (a private 32-row
        // slab per warp) -> groups of lanes write whole rows contiguously
        uint8_t* slab = smem + L.stage_off + ((g * 4 + q) * 32) * L.stage_pitch;
        uint8_t* myrow = slab + lane * L.stage_pitch;
        const bool st16 = !p.epi.out_fp32;                       // bf16 output: stage packed bf16
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + c0, v);
          tmem_ld_wait();
          if (st16) {
            uint4 u0, u1;
            u0.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
            u0.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
            u0.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
            u0.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
            u1.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));
            u1.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
            u1.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13]));
            u1.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
            reinterpret_cast<uint4*>(myrow + c0 * 2)[0] = u0;
            reinterpret_cast<uint4*>(myrow + c0 * 2)[1] = u1;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<uint4*>(myrow + c0 * 4)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        // the accumulator is drained into smem: release the TMEM stage before the global stores
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[a]);
        const int vec_per_row = p.N >> 3;                       // 8 outputs per lane
        const int rows_per_it = 32 / vec_per_row;               // N = 128 -> 16 lanes per row, 2 rows / instr
        const int lr = lane / vec_per_row, lc = lane % vec_per_row;
        const long long row0 = static_cast<long long>(tile) * kTileM + q * 32;
        for (int rb = lr; rb < 32; rb += 4 * rows_per_it) {
          uint4 addv[4];
          bool okv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {                          // issue all global loads first
            const int rr = rb + u * rows_per_it;
            okv[u] = rr < 32 && row0 + rr < p.M;
            addv[u] = make_uint4(0, 0, 0, 0);
            if (okv[u] && p.epi.add_src != nullptr)
              addv[u] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.epi.add_src) +
                                                        (row0 + rr) * p.epi.ld_add + lc * 8);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!okv[u]) continue;
            const int rr = rb + u * rows_per_it;
            const long long grow = row0 + rr;
            const uint8_t* srow = slab + rr * L.stage_pitch;
            if (st16) {
              uint4 sv = *reinterpret_cast<const uint4*>(srow + lc * 16);
              if (p.epi.add_src != nullptr) {
                float2 a, b;
                a = unpack_bf16x2(sv.x); b = unpack_bf16x2(addv[u].x); sv.x = pack_bf16x2(a.x + b.x, a.y + b.y);
                a = unpack_bf16x2(sv.y); b = unpack_bf16x2(addv[u].y); sv.y = pack_bf16x2(a.x + b.x, a.y + b.y);
                a = unpack_bf16x2(sv.z); b = unpack_bf16x2(addv[u].z); sv.z = pack_bf16x2(a.x + b.x, a.y + b.y);
                a = unpack_bf16x2(sv.w); b = unpack_bf16x2(addv[u].w); sv.w = pack_bf16x2(a.x + b.x, a.y + b.y);
              }
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.epi.peers[0]) + grow * p.epi.ldc + lc * 8) = sv;
            } else {
              float4 f0 = reinterpret_cast<const float4*>(srow + lc * 32)[0];
              float4 f1 = reinterpret_cast<const float4*>(srow + lc * 32)[1];
              float2 t;
              t = unpack_bf16x2(addv[u].x); f0.x += t.x; f0.y += t.y;
              t = unpack_bf16x2(addv[u].y); f0.z += t.x; f0.w += t.y;
              t = unpack_bf16x2(addv[u].z); f1.x += t.x; f1.y += t.y;
              t = unpack_bf16x2(addv[u].w); f1.z += t.x; f1.w += t.y;
              float* o = reinterpret_cast<float*>(p.epi.peers[0]) + grow * p.epi.ldc + lc * 8;
              reinterpret_cast<float4*>(o)[0] = f0;
              reinterpret_cast<float4*>(o)[1] = f1;
            }
          }
        }
        __syncwarp();                                            // slab reusable by this warp's next tile
        continue;
      } else if (p.epi.mode == EPI_ROWMAJOR) {
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + c0, v);
          tmem_ld_wait();
          if (!row_ok) continue;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          const int ncol = min(16, p.N - c0);
          const bool vec = p.epi.vec_ok && ncol == 16;
          if (p.epi.add_src != nullptr) {
            const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(p.epi.add_src) + row * p.epi.ld_add + c0;
            for (int i = 0; i < ncol; ++i) f[i] += __bfloat162float(ap[i]);
          }
          if (p.epi.out_fp32) {
            float* o = reinterpret_cast<float*>(p.epi.peers[0]) + row * p.epi.ldc + c0;
            if (vec) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            } else {
              for (int i = 0; i < ncol; ++i) o[i] = f[i];
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.epi.peers[0]) + row * p.epi.ldc + c0;
            if (vec) {
              uint4 u0, u1;
              u0.x = pack_bf16x2(f[0], f[1]);   u0.y = pack_bf16x2(f[2], f[3]);
              u0.z = pack_bf16x2(f[4], f[5]);   u0.w = pack_bf16x2(f[6], f[7]);
              u1.x = pack_bf16x2(f[8], f[9]);   u1.y = pack_bf16x2(f[10], f[11]);
              u1.z = pack_bf16x2(f[12], f[13]); u1.w = pack_bf16x2(f[14], f[15]);
              reinterpret_cast<uint4*>(o)[0] = u0;
              reinterpret_cast<uint4*>(o)[1] = u1;
            } else {
              for (int i = 0; i < ncol; ++i) o[i] = __float2bfloat16(f[i]);
            }
          }
        }
      } else if (p.epi.mode == EPI_PAIR_SCATTER) {
        // ---- pair scatter: (re, im) pairs to a mixed-radix address, possibly on a peer GPU
        int rpeer;
        const long long roff = row_offset(p.epi, static_cast<uint32_t>(row_ok ? row : 0), rpeer);
        __nv_bfloat16* const rbase = reinterpret_cast<__nv_bfloat16*>(p.epi.peers[rpeer]) + roff;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + c0, v);
          tmem_ld_wait();
          if (!row_ok) continue;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = (c0 >> 1) + i;
            if (j < npairs) {
              __nv_bfloat16* base = rbase;
              if (p.epi.peer_sel == PEER_BY_COL)
                base = reinterpret_cast<__nv_bfloat16*>(p.epi.peers[s_colpeer[j]]) + roff;
              *reinterpret_cast<uint32_t*>(base + s_coloff[j]) =
                  pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            }
          }
        }
      } else {
        // ---- projection head: out = b4 + sum_j W4[j] * gelu(acc[j] + b3[j])
        float part = s_vec[511];
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (c0 + i < p.N) {
              const float pre = __uint_as_float(v[i]) + s_vec[c0 + i];
              part = fmaf(s_vec[256 + c0 + i], gelu_erf(pre), part);
            }
          }
        }
        if (row_ok) {
          int unused;
          reinterpret_cast<float*>(p.epi.peers[0])[row_offset(p.epi, static_cast<uint32_t>(row), unused)] = part;
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
    if (p.epi.peer_sel != PEER_NONE) __threadfence_system();   // publish peer stores
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static void magic_for(unsigned d, unsigned long long* magic, int* shift) {
  int s = 0;
  while ((1ull << s) < d) ++s;
  *magic = ((1ull << (31 + s)) / d) + 1;
  *shift = 31 + s;
}

const char* dft_gemm_launch(const void* A, long long lda, const void* Bmat, GemmParams p, int num_sms,
                            cudaStream_t stream) {
  if (p.M <= 0) return nullptr;
  if (p.n_pad % 16 || p.n_pad < 16 || p.n_pad > 256) return "n_pad must be a multiple of 16 in [16,256]";
  if (p.k_pad % kBlockK || p.k_pad < kBlockK || p.k_pad > 512) return "k_pad must be a multiple of 64 in [64,512]";
  if (p.K > p.k_pad || p.N > p.n_pad) return "K/N exceed padded operator";
  if ((lda * 2) % 16) return "A row pitch must be a multiple of 16 bytes";
  if (reinterpret_cast<uintptr_t>(A) % 16 || reinterpret_cast<uintptr_t>(Bmat) % 16) return "A/B base must be 16B aligned";
  if (p.M > (1ll << 31) - 256) return "M too large for one launch";

  if (p.epi.mode == EPI_ROWMAJOR) {
    const long long esz = p.epi.out_fp32 ? 4 : 2;
    bool ok = (p.epi.ldc * esz) % 16 == 0 && reinterpret_cast<uintptr_t>(p.epi.peers[0]) % 16 == 0;
    if (p.epi.add_src)
      ok = ok && (p.epi.ld_add * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(p.epi.add_src) % 16 == 0;
    p.epi.vec_ok = ok ? 1 : 0;
  }
  SmemLayout L;
  const int kblocks = p.k_pad / kBlockK;
  L.b_bytes = static_cast<uint32_t>(kblocks) * p.n_pad * 128;
  const uint32_t budget = 227 * 1024 - 1024 /*align slack*/ - 4096 /*barriers + tables*/;
  // ring stage = the whole-K A tile when two of them fit next to the operator; for long K (up to 512)
  // the largest divisor of the K blocks that leaves room for >= 3 stages
  L.kbs = static_cast<uint32_t>(kblocks);
  if (L.b_bytes + 2u * kblocks * kTileM * 128 > budget) {
    uint32_t best = 0;
    for (uint32_t dv = 1; dv < static_cast<uint32_t>(kblocks); ++dv)
      if (kblocks % dv == 0 && L.b_bytes + 3u * dv * kTileM * 128 <= budget) best = dv;
    if (!best) return "operator too large for shared memory";
    L.kbs = best;
  }
  L.a_tile_bytes = L.kbs * kTileM * 128;
  // coalesced row-major epilogue: needs N to be a multiple of 8 dividing 256 and aligned rows
  L.stage_off = 0; L.stage_pitch = 0;
  uint32_t stage_total = 0;
  if (p.epi.mode == EPI_ROWMAJOR && p.epi.vec_ok && p.N % 8 == 0 && (p.N == 8 || p.N == 16 || p.N == 32 ||
      p.N == 64 || p.N == 128 || p.N == 256)) {
    const uint32_t pitch = ((p.N + 15) / 16 * 16) * (p.epi.out_fp32 ? 4 : 2) + 16;   // whole 16-column chunks are staged
    L.stage_pitch = pitch;
  }
  // epilogue groups and TMEM accumulator stages
  int E = 4;
  if (512 / p.n_pad < E) E = 512 / p.n_pad;                              // n_pad 256 -> 2 groups
  if (L.stage_pitch) {                                                   // one private slab per epilogue warp
    while (E > 1 && L.b_bytes + 2 * L.a_tile_bytes + 4u * E * 32 * L.stage_pitch > budget) E >>= 1;
    if (L.b_bytes + 2 * L.a_tile_bytes + 4u * E * 32 * L.stage_pitch > budget) { L.stage_pitch = 0; stage_total = 0; }
    else stage_total = 4u * E * 32 * L.stage_pitch;
  }
  int nacc = 512 / p.n_pad;
  if (nacc > kMaxAcc) nacc = kMaxAcc;
  nacc = nacc / E * E;
  L.nacc = static_cast<uint32_t>(nacc);
  for (int l = 0; l < 4; ++l) magic_for(static_cast<unsigned>(p.epi.R[l] > 0 ? p.epi.R[l] : 1), &p.epi.Rm[l], &p.epi.Rs[l]);
  magic_for(static_cast<unsigned>(p.epi.peer_div > 0 ? p.epi.peer_div : 1), &p.epi.Pm, &p.epi.Ps);
  L.stages = (budget - L.b_bytes - stage_total) / L.a_tile_bytes;
  if (L.stages > kMaxStages) L.stages = kMaxStages;
  if (stage_total) L.stage_off = L.b_bytes + L.stages * L.a_tile_bytes + 4096;
  uint32_t smem_bytes = L.b_bytes + L.stages * L.a_tile_bytes + 4096 + stage_total;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;     // force one CTA per SM (TMEM: 512 cols)

  CUtensorMap tmA, tmB;
  // A: the K tail beyond p.K (up to k_pad) and the M tail are zero-filled by TMA
  if (make_map_2d(&tmA, A, static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M), static_cast<uint64_t>(lda),
                  kBlockK, kTileM))
    return "cuTensorMapEncodeTiled(A) failed";
  if (make_map_2d(&tmB, Bmat, static_cast<uint64_t>(p.k_pad), static_cast<uint64_t>(p.n_pad),
                  static_cast<uint64_t>(p.k_pad), kBlockK, static_cast<uint32_t>(p.n_pad)))
    return "cuTensorMapEncodeTiled(B) failed";

  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(dft_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "cudaFuncSetAttribute(max dynamic smem) failed";
    attr_set = true;
  }
  const int num_tiles = static_cast<int>((p.M + kTileM - 1) / kTileM);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  dft_gemm_kernel<<<grid, 64 + 128 * E, smem_bytes, stream>>>(tmA, tmB, p, L);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

// pointwise.cu -- channel/time mixing kernels around the Fourier layers (sm_100a, CUDA cores).
//
// Internal activation layout of the fused engine: h[bc = b*C + c][x][y_local][t][z], bf16,
// z contiguous (so that every DFT stage is a K-major GEMM, see dft_gemm_sm100.cu).  The
// public tensors keep the reference layout [B, C, X, Y, Z, T] (t contiguous); the lift and
// the projection head are where the two layouts meet, so no transpose pass ever runs.
//
//   lift_fwd        : x[B,Cin,X,Y,Z,Tin] -> h = gelu(W2 ._c gelu(W1 ._t x + b1) + b2)
//                     (reference: linear1 -> gelu -> linear2 -> gelu, dfno.py:333-338; K15+K16)
//   lift_bwd        : dh -> dW1, db1, dW2, db2 (recomputes the tiny activations)
//   bypass_gelu_fwd : pre = spec + W ._c h ; out = gelu(pre)         (K2 + K14, dfno.py:244,291)
//   bypass_gelu_bwd : dpre = dout * gelu'(pre) ; dhb = W^T ._c dpre  (weight grad: kreduce GEMM)
//   to_channels_last / from_channels_last : layout bridges for the projection head
#include "sm100_ptx.cuh"
#include "kernels.h"

namespace dfno {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// ------------------------------------------------------------------------------------------
// lift
// ------------------------------------------------------------------------------------------
constexpr int kLiftMaxIn = 32;     // Cin * Tin values per position kept in registers
constexpr int kLiftMaxW = 4096;    // floats of shared memory for W1,b1,W2,b2

template <typename TIn>
__global__ void __launch_bounds__(128)
lift_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ b1,
                const float* __restrict__ W2, const float* __restrict__ b2, __nv_bfloat16* __restrict__ h,
                LiftDims d) {
  __shared__ float sw[kLiftMaxW];
  float* sW1 = sw;                         // [T][Tin]
  float* sb1 = sW1 + d.T * d.Tin;          // [T]
  float* sW2 = sb1 + d.T;                  // [C][Cin]
  float* sb2 = sW2 + d.C * d.Cin;          // [C]
  for (int i = threadIdx.x; i < d.T * d.Tin; i += blockDim.x) sW1[i] = W1[i];
  for (int i = threadIdx.x; i < d.T; i += blockDim.x) sb1[i] = b1[i];
  for (int i = threadIdx.x; i < d.C * d.Cin; i += blockDim.x) sW2[i] = W2[i];
  for (int i = threadIdx.x; i < d.C; i += blockDim.x) sb2[i] = b2[i];
  __syncthreads();

  const int zp = d.Z >> 1;                              // z pairs
  const long long npos = static_cast<long long>(d.B) * d.X * d.Y * zp;
  const long long plane = static_cast<long long>(d.X) * d.Y;   // (x,y) positions
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < npos;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int z2 = static_cast<int>(idx % zp);
    const long long xy = (idx / zp) % plane;
    const int b = static_cast<int>(idx / (zp * plane));
    float xin[2][kLiftMaxIn];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      for (int ci = 0; ci < d.Cin; ++ci) {
        const TIn* src = x + ((((static_cast<long long>(b) * d.Cin + ci) * plane + xy) * d.Z) + (2 * z2 + q)) * d.Tin;
        for (int ti = 0; ti < d.Tin; ++ti) xin[q][ci * d.Tin + ti] = ldf(src + ti);
      }
    }
    for (int t = 0; t < d.T; ++t) {
      float a1[2][8];
      for (int ci = 0; ci < d.Cin; ++ci) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float v = sb1[t];
          for (int ti = 0; ti < d.Tin; ++ti) v = fmaf(sW1[t * d.Tin + ti], xin[q][ci * d.Tin + ti], v);
          a1[q][ci] = gelu_erf(v);
        }
      }
      for (int c = 0; c < d.C; ++c) {
        float v0 = sb2[c], v1 = sb2[c];
        for (int ci = 0; ci < d.Cin; ++ci) {
          v0 = fmaf(sW2[c * d.Cin + ci], a1[0][ci], v0);
          v1 = fmaf(sW2[c * d.Cin + ci], a1[1][ci], v1);
        }
        __nv_bfloat16* dst = h + ((((static_cast<long long>(b) * d.C + c) * plane + xy) * d.T + t) * d.Z) + 2 * z2;
        *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(gelu_erf(v0), gelu_erf(v1));
      }
    }
  }
}

// dW1[T][Tin], db1[T], dW2[C][Cin], db2[C] accumulated with atomics into fp32 buffers.
// Channel-indexed sums (db2, dW2) stay in registers over the whole grid-stride loop and are
// reduced once per thread; time-indexed sums (db1, dW1) are warp-reduced once per t.
template <typename TIn, int C, int CIN>
__global__ void __launch_bounds__(128)
lift_bwd_kernel(const TIn* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ b1,
                const float* __restrict__ W2, const float* __restrict__ b2,
                const __nv_bfloat16* __restrict__ dh, float* __restrict__ gW1, float* __restrict__ gb1,
                float* __restrict__ gW2, float* __restrict__ gb2, LiftDims d) {
  __shared__ float sw[kLiftMaxW];
  __shared__ float sg[kLiftMaxW];          // block-local gradient accumulators, same layout
  float* sW1 = sw;
  float* sb1 = sW1 + d.T * d.Tin;
  float* sW2 = sb1 + d.T;
  float* sb2 = sW2 + C * CIN;
  const int nW = d.T * d.Tin + d.T + C * CIN + C;
  float* gsW1 = sg;
  float* gsb1 = gsW1 + d.T * d.Tin;
  float* gsW2 = gsb1 + d.T;
  float* gsb2 = gsW2 + C * CIN;
  for (int i = threadIdx.x; i < d.T * d.Tin; i += blockDim.x) sW1[i] = W1[i];
  for (int i = threadIdx.x; i < d.T; i += blockDim.x) sb1[i] = b1[i];
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) sW2[i] = W2[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb2[i] = b2[i];
  for (int i = threadIdx.x; i < nW; i += blockDim.x) sg[i] = 0.f;
  __syncthreads();

  float accb2[C], accW2[C][CIN];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    accb2[c] = 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) accW2[c][ci] = 0.f;
  }

  const int lane = threadIdx.x & 31;
  const int zp = d.Z >> 1;
  const long long npos = static_cast<long long>(d.B) * d.X * d.Y * zp;
  const long long plane = static_cast<long long>(d.X) * d.Y;
  const long long nloop = (npos + static_cast<long long>(gridDim.x) * blockDim.x - 1) /
                          (static_cast<long long>(gridDim.x) * blockDim.x);
  for (long long it = 0; it < nloop; ++it) {
    const long long idx = (it * gridDim.x + blockIdx.x) * static_cast<long long>(blockDim.x) + threadIdx.x;
    const bool ok = idx < npos;                          // whole warps stay in the loop (shuffles)
    const long long id = ok ? idx : 0;
    const int z2 = static_cast<int>(id % zp);
    const long long xy = (id / zp) % plane;
    const int b = static_cast<int>(id / (zp * plane));
    float xin[2][kLiftMaxIn];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const TIn* src = x + ((((static_cast<long long>(b) * CIN + ci) * plane + xy) * d.Z) + (2 * z2 + q)) * d.Tin;
        for (int ti = 0; ti < d.Tin; ++ti) xin[q][ci * d.Tin + ti] = ok ? ldf(src + ti) : 0.f;
      }
    for (int t = 0; t < d.T; ++t) {
      float a1[2][CIN], g1p[2][CIN], da1[2][CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float v = sb1[t];
          for (int ti = 0; ti < d.Tin; ++ti) v = fmaf(sW1[t * d.Tin + ti], xin[q][ci * d.Tin + ti], v);
          a1[q][ci] = gelu_erf(v);
          g1p[q][ci] = gelu_erf_grad(v);
          da1[q][ci] = 0.f;
        }
      }
      const __nv_bfloat16* src = dh + (((static_cast<long long>(b) * C * plane + xy) * d.T + t) * d.Z) + 2 * z2;
      const long long cstride = plane * d.T * d.Z;
      uint32_t gv[C];                                     // all channel loads of this t in flight together
#pragma unroll
      for (int c = 0; c < C; ++c) gv[c] = ok ? *reinterpret_cast<const uint32_t*>(src + c * cstride) : 0u;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v0 = sb2[c], v1 = sb2[c];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          v0 = fmaf(sW2[c * CIN + ci], a1[0][ci], v0);
          v1 = fmaf(sW2[c * CIN + ci], a1[1][ci], v1);
        }
        const float2 g = unpack_bf16x2(gv[c]);
        const float g0 = g.x * gelu_erf_grad(v0);
        const float g1 = g.y * gelu_erf_grad(v1);
        accb2[c] += g0 + g1;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          accW2[c][ci] += g0 * a1[0][ci] + g1 * a1[1][ci];
          da1[0][ci] = fmaf(sW2[c * CIN + ci], g0, da1[0][ci]);
          da1[1][ci] = fmaf(sW2[c * CIN + ci], g1, da1[1][ci]);
        }
      }
      float sb1v = 0.f;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float e0 = da1[0][ci] * g1p[0][ci], e1 = da1[1][ci] * g1p[1][ci];
        sb1v += e0 + e1;
        for (int ti = 0; ti < d.Tin; ++ti) {
          float sres = warp_sum(e0 * xin[0][ci * d.Tin + ti] + e1 * xin[1][ci * d.Tin + ti]);
          if (lane == 0) atomicAdd(&gsW1[t * d.Tin + ti], sres);
        }
      }
      sb1v = warp_sum(sb1v);
      if (lane == 0) atomicAdd(&gsb1[t], sb1v);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float sres = warp_sum(accb2[c]);
    if (lane == 0) atomicAdd(&gsb2[c], sres);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float sw2 = warp_sum(accW2[c][ci]);
      if (lane == 0) atomicAdd(&gsW2[c * CIN + ci], sw2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.T * d.Tin; i += blockDim.x) atomicAdd(&gW1[i], gsW1[i]);
  for (int i = threadIdx.x; i < d.T; i += blockDim.x) atomicAdd(&gb1[i], gsb1[i]);
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) atomicAdd(&gW2[i], gsW2[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&gb2[i], gsb2[i]);
}

// ------------------------------------------------------------------------------------------
// bypass conv + GELU
// ------------------------------------------------------------------------------------------
// Each thread owns 2 consecutive z of one (b, x, y, t) and all C channels.
// spec_pre: in = spectral branch, out (in place) = pre-activation (kept for the backward).
template <int C, bool kCL>
__global__ void __launch_bounds__(128)
bypass_gelu_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ spec_pre,
                       const float* __restrict__ W, __nv_bfloat16* __restrict__ out,
                       __nv_bfloat16* __restrict__ out_cl, int cl_pitch, int B, long long S, int save_pre) {
  __shared__ __align__(16) float sW[C * C];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sW[i] = W[i];
  __syncthreads();
  const long long S2 = S >> 1;                       // position pairs per channel
  const long long total = static_cast<long long>(B) * S2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = idx / S2, p2 = idx % S2;
    const long long base = b * C * S + 2 * p2;
    float h0[C], h1[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(h + base + i * S));
      h0[i] = v.x; h1[i] = v.y;
    }
    // channels-last rows (for the projection head) are assembled in registers and written as
    // 8-byte vectors: two rows of cl_pitch bf16 per thread
    uint32_t cl0[kCL ? C / 2 + 1 : 1], cl1[kCL ? C / 2 + 1 : 1];
    // all spectral-branch values are fetched up front: with few resident warps the loads
    // must overlap each other, not the dependent FMA chains
    uint32_t sp[C];
#pragma unroll
    for (int o = 0; o < C; ++o) sp[o] = *reinterpret_cast<const uint32_t*>(spec_pre + base + o * S);
#pragma unroll
    for (int o = 0; o < C; ++o) {
      const float2 s = unpack_bf16x2(sp[o]);
      float a0 = s.x, a1 = s.y;
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float w = sW[o * C + i];
        a0 = fmaf(w, h0[i], a0);
        a1 = fmaf(w, h1[i], a1);
      }
      if (save_pre) *reinterpret_cast<uint32_t*>(spec_pre + base + o * S) = pack_bf16x2(a0, a1);
      const float y0 = gelu_erf(a0), y1 = gelu_erf(a1);
      if (out) *reinterpret_cast<uint32_t*>(out + base + o * S) = pack_bf16x2(y0, y1);
      if (kCL) {
        // even channel: low half, odd channel: high half of the packed word
        if ((o & 1) == 0) {
          cl0[o >> 1] = __bfloat16_as_ushort(__float2bfloat16(y0));
          cl1[o >> 1] = __bfloat16_as_ushort(__float2bfloat16(y1));
        } else {
          cl0[o >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16(y0))) << 16;
          cl1[o >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16(y1))) << 16;
        }
      }
    }
    if (kCL) {
      static_assert(C % 4 == 0, "channels-last rows are written as 8-byte vectors");
      __nv_bfloat16* r0 = out_cl + (b * S + 2 * p2) * cl_pitch;
      __nv_bfloat16* r1 = r0 + cl_pitch;
#pragma unroll
      for (int w2 = 0; w2 < C / 4; ++w2) {
        reinterpret_cast<uint2*>(r0)[w2] = make_uint2(cl0[2 * w2], cl0[2 * w2 + 1]);
        reinterpret_cast<uint2*>(r1)[w2] = make_uint2(cl1[2 * w2], cl1[2 * w2 + 1]);
      }
    }
  }
}

// dpre = dout * gelu'(pre);  dhb = W^T dpre.  dout may come channel-major (internal layout)
// or channels-last (from the projection head).
template <int C>
__global__ void __launch_bounds__(256)
bypass_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ dout_cl,
                       int cl_pitch, const __nv_bfloat16* __restrict__ pre, const float* __restrict__ W,
                       __nv_bfloat16* __restrict__ dpre, __nv_bfloat16* __restrict__ dhb, int B, long long S) {
  __shared__ __align__(16) float sW[C * C];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sW[i] = W[i];
  __syncthreads();
  const long long S2 = S >> 1;
  const long long total = static_cast<long long>(B) * S2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = idx / S2, p2 = idx % S2;
    const long long base = b * C * S + 2 * p2;
    float g0[C], g1[C];
#pragma unroll
    for (int o = 0; o < C; ++o) {
      float2 g;
      if (dout_cl) {
        const long long r = (b * S + 2 * p2) * cl_pitch + o;
        g.x = __bfloat162float(dout_cl[r]);
        g.y = __bfloat162float(dout_cl[r + cl_pitch]);
      } else {
        g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + base + o * S));
      }
      const float2 p = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pre + base + o * S));
      g0[o] = g.x * gelu_erf_grad(p.x);
      g1[o] = g.y * gelu_erf_grad(p.y);
      *reinterpret_cast<uint32_t*>(dpre + base + o * S) = pack_bf16x2(g0[o], g1[o]);
    }
#pragma unroll 4
    for (int i = 0; i < C; ++i) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int o = 0; o < C; ++o) {
        const float w = sW[o * C + i];
        a0 = fmaf(w, g0[o], a0);
        a1 = fmaf(w, g1[o], a1);
      }
      *reinterpret_cast<uint32_t*>(dhb + base + i * S) = pack_bf16x2(a0, a1);
    }
  }
}

__global__ void gelu_probe_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ dy, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) { y[i] = gelu_erf(x[i]); dy[i] = gelu_erf_grad(x[i]); }
}

// Generic strided permutation of 32-bit words (one (re, im) bf16 pair each): dst is walked in its
// own mixed-radix order (innermost digit first), the same digits address src through src_strides.
// Used on the receiving side of the fused pencil transposes: peers deposit their contribution as
// one long contiguous run per source rank (NVLink-friendly), this kernel interleaves the runs into
// the K-major layout the next GEMM stage reads.  The tensors are the *truncated* spectra, a few MB.
struct PermuteDesc {
  int nd;
  unsigned size[6];
  unsigned long long magic[6];
  int shift[6];
  long long sstr[6];
  long long dstr[6];
};

// V = words per thread access (the innermost digit is contiguous on both sides and a multiple of V);
// digits by magic-number division (n < 2^31); kUnroll independent loads in flight per thread.
template <typename Vec, int kUnroll>
__global__ void __launch_bounds__(256)
permute_vec_kernel(const Vec* __restrict__ src, Vec* __restrict__ dst, unsigned total, PermuteDesc d) {
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * kUnroll) {
    Vec v[kUnroll];
    long long dof[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      unsigned r = i0 + u * stride;
      long long so = 0;
      dof[u] = -1;
      if (r < total) {
        dof[u] = 0;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          if (l < d.nd) {
            unsigned dig = r;
            if (l != d.nd - 1) {
              const unsigned q = static_cast<unsigned>((static_cast<unsigned long long>(r) * d.magic[l]) >> d.shift[l]);
              dig = r - q * d.size[l];
              r = q;
            }
            so += static_cast<long long>(dig) * d.sstr[l];
            dof[u] += static_cast<long long>(dig) * d.dstr[l];
          }
        }
        v[u] = src[so];
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dof[u] >= 0) dst[dof[u]] = v[u];
  }
}

int grid_for(long long work_items, int threads, int num_sms, int per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms) * per_sm;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
const char* permute_u32(const void* src, void* dst, int nd, const int* size, const long long* sstr,
                        const long long* dstr, int num_sms, cudaStream_t s) {
  if (nd < 1 || nd > 6) return "permute_u32: 1..6 digits";
  // widest vector the innermost (contiguous) digit allows
  int V = 1;
  if (sstr[0] == 1 && dstr[0] == 1) {
    for (int cand = 4; cand > 1 && V == 1; cand >>= 1) {
      bool ok = size[0] % cand == 0 && reinterpret_cast<uintptr_t>(src) % (4 * cand) == 0 &&
                reinterpret_cast<uintptr_t>(dst) % (4 * cand) == 0;
      for (int i = 1; i < nd; ++i) ok = ok && sstr[i] % cand == 0 && dstr[i] % cand == 0;
      if (ok) V = cand;
    }
  }
  PermuteDesc d;
  d.nd = nd;
  long long total = 1;
  for (int i = 0; i < 6; ++i) {
    long long sz = i < nd ? size[i] : 1;
    if (sz <= 0) return nullptr;
    if (i == 0) sz /= V;
    d.size[i] = static_cast<unsigned>(sz);
    d.sstr[i] = i < nd ? (i == 0 ? sstr[i] : sstr[i] / V) : 0;
    d.dstr[i] = i < nd ? (i == 0 ? dstr[i] : dstr[i] / V) : 0;
    int sh = 0;
    while ((1ull << sh) < static_cast<unsigned long long>(sz)) ++sh;
    d.magic[i] = ((1ull << (31 + sh)) / static_cast<unsigned long long>(sz)) + 1;
    d.shift[i] = 31 + sh;
    total *= sz;
  }
  if (total >= (1ll << 31)) return "permute_u32: tensor too large for one launch";
  constexpr int kUnroll = 4;
  long long blocks = (total + 256 * kUnroll - 1) / (256 * kUnroll);
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  const unsigned tot = static_cast<unsigned>(total);
  const int g = static_cast<int>(blocks);
  if (V == 4)
    permute_vec_kernel<uint4, kUnroll><<<g, 256, 0, s>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), tot, d);
  else if (V == 2)
    permute_vec_kernel<uint2, kUnroll><<<g, 256, 0, s>>>(static_cast<const uint2*>(src), static_cast<uint2*>(dst), tot, d);
  else
    permute_vec_kernel<uint32_t, kUnroll><<<g, 256, 0, s>>>(static_cast<const uint32_t*>(src),
                                                            static_cast<uint32_t*>(dst), tot, d);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* gelu_probe(const float* x, float* y, float* dy, long long n, cudaStream_t s) {
  gelu_probe_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(x, y, dy, n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* lift_fwd(const void* x, int x_is_bf16, const float* W1, const float* b1, const float* W2,
                     const float* b2, void* h, LiftDims d, int num_sms, cudaStream_t s) {
  if (d.Z % 2) return "Z must be even";
  if (d.Cin * d.Tin > kLiftMaxIn || d.Cin > 8) return "lift: Cin*Tin too large for the fused kernel";
  if (d.T * d.Tin + d.T + d.C * d.Cin + d.C > kLiftMaxW) return "lift: weights exceed shared memory budget";
  const long long npos = static_cast<long long>(d.B) * d.X * d.Y * (d.Z / 2);
  const int grid = grid_for(npos, 128, num_sms, 8);
  if (x_is_bf16)
    lift_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), W1, b1, W2, b2,
                                                        static_cast<__nv_bfloat16*>(h), d);
  else
    lift_fwd_kernel<float><<<grid, 128, 0, s>>>(static_cast<const float*>(x), W1, b1, W2, b2,
                                                static_cast<__nv_bfloat16*>(h), d);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

#define DFNO_DISPATCH_C(C_, BODY)                 \
  switch (C_) {                                   \
    case 4:  { constexpr int kC = 4;  BODY; } break;  \
    case 8:  { constexpr int kC = 8;  BODY; } break;  \
    case 12: { constexpr int kC = 12; BODY; } break;  \
    case 16: { constexpr int kC = 16; BODY; } break;  \
    case 20: { constexpr int kC = 20; BODY; } break;  \
    case 24: { constexpr int kC = 24; BODY; } break;  \
    case 32: { constexpr int kC = 32; BODY; } break;  \
    default: return "unsupported channel width (supported: 4,8,12,16,20,24,32)"; \
  }

template <int C>
static const char* lift_bwd_cin(const void* x, int x_is_bf16, const float* W1, const float* b1, const float* W2,
                                const float* b2, const void* dh, float* gW1, float* gb1, float* gW2, float* gb2,
                                LiftDims d, int grid, cudaStream_t s) {
#define DFNO_LIFT_BWD(CIN_)                                                                                       \
  if (x_is_bf16)                                                                                                  \
    lift_bwd_kernel<__nv_bfloat16, C, CIN_><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), W1, b1, W2, \
                                                                 b2, static_cast<const __nv_bfloat16*>(dh), gW1,  \
                                                                 gb1, gW2, gb2, d);                               \
  else                                                                                                            \
    lift_bwd_kernel<float, C, CIN_><<<grid, 128, 0, s>>>(static_cast<const float*>(x), W1, b1, W2, b2,            \
                                                         static_cast<const __nv_bfloat16*>(dh), gW1, gb1, gW2, gb2, d);
  switch (d.Cin) {
    case 1: DFNO_LIFT_BWD(1); break;
    case 2: DFNO_LIFT_BWD(2); break;
    case 3: DFNO_LIFT_BWD(3); break;
    case 4: DFNO_LIFT_BWD(4); break;
    default: return "lift_bwd: supported input channel counts are 1..4";
  }
#undef DFNO_LIFT_BWD
  return nullptr;
}

const char* lift_bwd(const void* x, int x_is_bf16, const float* W1, const float* b1, const float* W2,
                     const float* b2, const void* dh, float* gW1, float* gb1, float* gW2, float* gb2,
                     LiftDims d, int num_sms, cudaStream_t s) {
  if (d.Z % 2) return "Z must be even";
  if (d.Cin * d.Tin > kLiftMaxIn || d.Cin > 4) return "lift: Cin*Tin too large for the fused kernel";
  if (d.T * d.Tin + d.T + d.C * d.Cin + d.C > kLiftMaxW) return "lift: weights exceed shared memory budget";
  const long long npos = static_cast<long long>(d.B) * d.X * d.Y * (d.Z / 2);
  const int grid = grid_for(npos, 128, num_sms, 4);
  const char* err = nullptr;
  DFNO_DISPATCH_C(d.C, (err = lift_bwd_cin<kC>(x, x_is_bf16, W1, b1, W2, b2, dh, gW1, gb1, gW2, gb2, d, grid, s)));
  if (err) return err;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* bypass_gelu_fwd(const void* h, void* spec_pre, const float* W, void* out, void* out_cl, int cl_pitch,
                            int B, int C, long long S, int save_pre, int num_sms, cudaStream_t s) {
  if (S % 2) return "spatial size per channel must be even";
  const int grid = grid_for(static_cast<long long>(B) * (S / 2), 128, num_sms, 16);
  if (out_cl) {
    DFNO_DISPATCH_C(C, (bypass_gelu_fwd_kernel<kC, true><<<grid, 128, 0, s>>>(
                           static_cast<const __nv_bfloat16*>(h), static_cast<__nv_bfloat16*>(spec_pre), W,
                           static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(out_cl), cl_pitch, B, S,
                           save_pre)));
  } else {
    DFNO_DISPATCH_C(C, (bypass_gelu_fwd_kernel<kC, false><<<grid, 128, 0, s>>>(
                           static_cast<const __nv_bfloat16*>(h), static_cast<__nv_bfloat16*>(spec_pre), W,
                           static_cast<__nv_bfloat16*>(out), nullptr, cl_pitch, B, S, save_pre)));
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* bypass_gelu_bwd(const void* dout, const void* dout_cl, int cl_pitch, const void* pre, const float* W,
                            void* dpre, void* dhb, int B, int C, long long S, int num_sms, cudaStream_t s) {
  if (S % 2) return "spatial size per channel must be even";
  const int grid = grid_for(static_cast<long long>(B) * (S / 2), 256, num_sms, 4);
  DFNO_DISPATCH_C(C, (bypass_gelu_bwd_kernel<kC><<<grid, 256, 0, s>>>(
                         static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(dout_cl),
                         cl_pitch, static_cast<const __nv_bfloat16*>(pre), W, static_cast<__nv_bfloat16*>(dpre),
                         static_cast<__nv_bfloat16*>(dhb), B, S)));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

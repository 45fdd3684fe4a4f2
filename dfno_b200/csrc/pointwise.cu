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
// One thread owns 8 consecutive z of one (b, x, y) -- a 16-byte bf16 vector of every output row it produces,
// so a warp writes whole 256-byte z-lines -- and walks t and c.  Both GELUs run in packed fp16 (sm100_ptx.cuh):
// the outer one is evaluated B*C*X*Y*Z*T times per step.  The few input values a thread needs stay in
// registers when Tin == 1 (the benchmark / two-phase case) and are re-read through L1 otherwise (Cin <= 4, Tin <= 64).
constexpr int kLiftMaxTin = 64;
constexpr int kLiftMaxW = 4096;    // floats of shared memory for W1, b1 (+ packed W2, b2)

template <typename TIn>
__device__ __forceinline__ void lift_load8(const TIn* __restrict__ src, int Tin, int ti, float (&v)[8]) {
#pragma unroll
  for (int z = 0; z < 8; ++z) v[z] = ldf(src + z * Tin + ti);
}
template <>
__device__ __forceinline__ void lift_load8<float>(const float* __restrict__ src, int Tin, int ti, float (&v)[8]) {
  if (Tin == 1) {
    const float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int z = 0; z < 8; ++z) v[z] = src[z * Tin + ti];
  }
}

// v1[z] = b1[t] + sum_ti W1[t, ti] x[ci, z, ti]
template <typename TIn, bool kRegs>
__device__ __forceinline__ void lift_inner(const TIn* __restrict__ xci, const float (*xr)[8], int ci, int Tin,
                                           const float* sW1t, float b1t, float (&v)[8]) {
#pragma unroll
  for (int z = 0; z < 8; ++z) v[z] = b1t;
  for (int ti = 0; ti < Tin; ++ti) {
    float xv[8];
    if (kRegs) {                        // Tin == 1
#pragma unroll
      for (int z = 0; z < 8; ++z) xv[z] = xr[ci][z];
    } else {
      lift_load8<TIn>(xci, Tin, ti, xv);
    }
    const float w = sW1t[ti];
#pragma unroll
    for (int z = 0; z < 8; ++z) v[z] = fmaf(w, xv[z], v[z]);
  }
}

template <typename TIn, int CIN, bool kRegs>
__global__ void __launch_bounds__(256)
lift_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ b1,
                const float* __restrict__ W2, const float* __restrict__ b2, __nv_bfloat16* __restrict__ h,
                LiftDims d) {
  __shared__ float sw[kLiftMaxW];
  float* sW1 = sw;                                          // [T][Tin]
  float* sb1 = sW1 + d.T * d.Tin;                           // [T]
  __half2* sW2 = reinterpret_cast<__half2*>(sb1 + d.T);     // [C][CIN] (value duplicated in both halves)
  __half2* sb2 = sW2 + d.C * CIN;                           // [C]
  for (int i = threadIdx.x; i < d.T * d.Tin; i += blockDim.x) sW1[i] = W1[i];
  for (int i = threadIdx.x; i < d.T; i += blockDim.x) sb1[i] = b1[i];
  for (int i = threadIdx.x; i < d.C * CIN; i += blockDim.x) sW2[i] = __float2half2_rn(W2[i]);
  for (int i = threadIdx.x; i < d.C; i += blockDim.x) sb2[i] = __float2half2_rn(b2[i]);
  __syncthreads();

  const int zv = d.Z >> 3;
  const long long plane = static_cast<long long>(d.X) * d.Y;
  const long long nitems = static_cast<long long>(d.B) * plane * zv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < nitems;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int z0 = static_cast<int>(idx % zv) * 8;
    const long long xy = (idx / zv) % plane;
    const int b = static_cast<int>(idx / (zv * plane));
    const TIn* xb = x + (((static_cast<long long>(b) * CIN) * plane + xy) * d.Z + z0) * d.Tin;
    const long long xci_stride = plane * d.Z * d.Tin;
    float xr[kRegs ? CIN : 1][8];
    if (kRegs) {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) lift_load8<TIn>(xb + ci * xci_stride, 1, 0, xr[ci]);
    }
    __nv_bfloat16* hb = h + ((static_cast<long long>(b) * d.C * plane + xy) * d.T) * d.Z + z0;
    const long long hc_stride = plane * d.T * d.Z;
    for (int t = 0; t < d.T; ++t) {
      __half2 a1[CIN][4];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float v[8];
        lift_inner<TIn, kRegs>(xb + ci * xci_stride, xr, ci, d.Tin, sW1 + t * d.Tin, sb1[t], v);
#pragma unroll
        for (int k = 0; k < 4; ++k) a1[ci][k] = gelu_h2(h2_from_f32(v[2 * k], v[2 * k + 1]));
      }
      for (int c = 0; c < d.C; ++c) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __half2 acc = sb2[c];
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) acc = __hfma2(sW2[c * CIN + ci], a1[ci][k], acc);
          o[k] = h2_to_bf16x2(gelu_h2(acc));
        }
        *reinterpret_cast<uint4*>(hb + c * hc_stride + static_cast<long long>(t) * d.Z) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// dW1[T][Tin], db1[T], dW2[C][Cin], db2[C] accumulated with atomics into fp32 buffers.  Four adjacent lanes share
// one item (8 consecutive z of one (b, x, y)) and split the C channels between them: a thread keeps C/4 channel
// sums in registers over the whole grid-stride loop (with all C channels per thread the kernel needed 241
// registers and ran at 12 % occupancy, 0.21 of copy bandwidth: profiles/r2_ncu_kernels.json), the input-gradient
// partials of the four quarters meet in two shuffles per value, and time-indexed sums are warp-reduced once per
// t.  The loss gradient can be far below the fp16 range, so everything it multiplies is fp32; only the GELU'
// evaluations are packed fp16.
template <typename TIn, int C, int CIN, bool kRegs>
__global__ void __launch_bounds__(128, (CIN == 1 ? 4 : 2))      // several input channels: more live values, no spills
lift_bwd_kernel(const TIn* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ b1,
                const float* __restrict__ W2, const float* __restrict__ b2,
                const __nv_bfloat16* __restrict__ dh, float* __restrict__ gW1, float* __restrict__ gb1,
                float* __restrict__ gW2, float* __restrict__ gb2, LiftDims d) {
  static_assert(C % 4 == 0, "the channels are split over four lanes");
  constexpr int CG = C / 4;
  __shared__ float sw[kLiftMaxW];
  __shared__ float sg[kLiftMaxW];
  float* sW1 = sw;
  float* sb1 = sW1 + d.T * d.Tin;
  float* sW2f = sb1 + d.T;                                   // [C][CIN] fp32 (input-gradient path)
  __half2* sW2 = reinterpret_cast<__half2*>(sW2f + C * CIN); // [C][CIN] packed fp16 (recomputation)
  __half2* sb2 = sW2 + C * CIN;
  float* gsW1 = sg;
  float* gsb1 = gsW1 + d.T * d.Tin;
  for (int i = threadIdx.x; i < d.T * d.Tin; i += blockDim.x) { sW1[i] = W1[i]; gsW1[i] = 0.f; }
  for (int i = threadIdx.x; i < d.T; i += blockDim.x) { sb1[i] = b1[i]; gsb1[i] = 0.f; }
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) { sW2f[i] = W2[i]; sW2[i] = __float2half2_rn(W2[i]); }
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb2[i] = __float2half2_rn(b2[i]);
  __syncthreads();

  float accb2[CG], accW2[CG][CIN];
#pragma unroll
  for (int u = 0; u < CG; ++u) {
    accb2[u] = 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) accW2[u][ci] = 0.f;
  }
  const int lane = threadIdx.x & 31;
  const int cg = lane & 3, c0 = cg * CG;                     // this lane's channels: [c0, c0 + CG)
  const int zv = d.Z >> 3;
  const long long plane = static_cast<long long>(d.X) * d.Y;
  const long long nitems = static_cast<long long>(d.B) * plane * zv;
  const long long per_it = (static_cast<long long>(gridDim.x) * blockDim.x) >> 2;
  const long long nloop = (nitems + per_it - 1) / per_it;
  const long long item0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 2;
  for (long long it = 0; it < nloop; ++it) {
    const long long idx = it * per_it + item0;
    const bool ok = idx < nitems;                           // whole warps stay in the loop (shuffles)
    const long long id = ok ? idx : 0;
    const int z0 = static_cast<int>(id % zv) * 8;
    const long long xy = (id / zv) % plane;
    const int b = static_cast<int>(id / (zv * plane));
    const TIn* xb = x + (((static_cast<long long>(b) * CIN) * plane + xy) * d.Z + z0) * d.Tin;
    const long long xci_stride = plane * d.Z * d.Tin;
    float xr[kRegs ? CIN : 1][8];
    if (kRegs) {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) lift_load8<TIn>(xb + ci * xci_stride, 1, 0, xr[ci]);
    }
    const long long hc_stride = plane * d.T * d.Z;
    const __nv_bfloat16* gb = dh + ((static_cast<long long>(b) * C * plane + xy) * d.T) * d.Z + z0 + c0 * hc_stride;
    for (int t = 0; t < d.T; ++t) {
      __half2 a1[CIN][4];
      float a1f[CIN][8], g1f[CIN][8], da1[CIN][8];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float v[8];
        lift_inner<TIn, kRegs>(xb + ci * xci_stride, xr, ci, d.Tin, sW1 + t * d.Tin, sb1[t], v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const GeluH2 vg = gelu_vg_h2(h2_from_f32(v[2 * k], v[2 * k + 1]));
          a1[ci][k] = vg.value;
          const float2 av = __half22float2(vg.value), gv = __half22float2(vg.grad);
          a1f[ci][2 * k] = av.x; a1f[ci][2 * k + 1] = av.y;
          g1f[ci][2 * k] = gv.x; g1f[ci][2 * k + 1] = gv.y;
          da1[ci][2 * k] = 0.f; da1[ci][2 * k + 1] = 0.f;
        }
      }
      uint4 gv[CG];                                         // this lane's channel loads, all in flight together
#pragma unroll
      for (int u = 0; u < CG; ++u)
        gv[u] = ok ? *reinterpret_cast<const uint4*>(gb + u * hc_stride + static_cast<long long>(t) * d.Z)
                   : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < CG; ++u) {
        const int c = c0 + u;
        const uint32_t gw[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
        float dv[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __half2 acc = sb2[c];
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) acc = __hfma2(sW2[c * CIN + ci], a1[ci][k], acc);
          const float2 gr = __half22float2(gelu_vg_h2(acc).grad);
          const float2 g = unpack_bf16x2(gw[k]);
          dv[2 * k] = g.x * gr.x; dv[2 * k + 1] = g.y * gr.y;
        }
        float sb = 0.f;
#pragma unroll
        for (int z = 0; z < 8; ++z) sb += dv[z];
        accb2[u] += sb;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float w = sW2f[c * CIN + ci];
          float sacc = 0.f;
#pragma unroll
          for (int z = 0; z < 8; ++z) {
            sacc = fmaf(dv[z], a1f[ci][z], sacc);
            da1[ci][z] = fmaf(w, dv[z], da1[ci][z]);
          }
          accW2[u][ci] += sacc;
        }
      }
      // input-gradient partials of the four channel quarters -> every lane of the quad holds the full sum
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int z = 0; z < 8; ++z) {
          float v = da1[ci][z];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          da1[ci][z] = v;
        }
      }
      const bool own = ok && cg == 0;                       // one lane of the quad feeds the time-indexed sums
      float sb1v = 0.f;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float e[8];
#pragma unroll
        for (int z = 0; z < 8; ++z) { e[z] = da1[ci][z] * g1f[ci][z]; sb1v += e[z]; }
        for (int ti = 0; ti < d.Tin; ++ti) {
          float xv[8];
          if (kRegs) {
#pragma unroll
            for (int z = 0; z < 8; ++z) xv[z] = xr[ci][z];
          } else {
            lift_load8<TIn>(xb + ci * xci_stride, d.Tin, ti, xv);
          }
          float sres = 0.f;
#pragma unroll
          for (int z = 0; z < 8; ++z) sres = fmaf(e[z], xv[z], sres);
          sres = warp_sum(own ? sres : 0.f);
          if (lane == 0) atomicAdd(&gsW1[t * d.Tin + ti], sres);
        }
      }
      sb1v = warp_sum(own ? sb1v : 0.f);
      if (lane == 0) atomicAdd(&gsb1[t], sb1v);
    }
  }
  __shared__ float gsW2[64 * 4 + 64];
  for (int i = threadIdx.x; i < C * CIN + C; i += blockDim.x) gsW2[i] = 0.f;
  __syncthreads();
  // channel sums: lanes with the same quarter (lane & 3) hold partials of the same channels
  auto quarter_sum = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
  };
#pragma unroll
  for (int u = 0; u < CG; ++u) {
    const float sres = quarter_sum(accb2[u]);
    if (lane < 4) atomicAdd(&gsW2[C * CIN + c0 + u], sres);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float sw2 = quarter_sum(accW2[u][ci]);
      if (lane < 4) atomicAdd(&gsW2[(c0 + u) * CIN + ci], sw2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.T * d.Tin; i += blockDim.x) atomicAdd(&gW1[i], gsW1[i]);
  for (int i = threadIdx.x; i < d.T; i += blockDim.x) atomicAdd(&gb1[i], gsb1[i]);
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) atomicAdd(&gW2[i], gsW2[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&gb2[i], gsW2[C * CIN + i]);
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

// the packed fp16 GELU the fused kernels use (pairs of adjacent elements)
__global__ void gelu_probe_h2_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ dy, long long n) {
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 2;
  if (i + 1 < n) {
    const GeluH2 r = gelu_vg_h2(h2_from_f32(x[i], x[i + 1]));
    const float2 v = __half22float2(r.value), g = __half22float2(r.grad);
    const float2 v2 = __half22float2(gelu_h2(h2_from_f32(x[i], x[i + 1])));
    y[i] = v.x; y[i + 1] = v2.y; dy[i] = g.x; dy[i + 1] = g.y;
  }
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

const char* gelu_probe_h2(const float* x, float* y, float* dy, long long n, cudaStream_t s) {
  gelu_probe_h2_kernel<<<static_cast<int>((n / 2 + 255) / 256), 256, 0, s>>>(x, y, dy, n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* gelu_probe(const float* x, float* y, float* dy, long long n, cudaStream_t s) {
  gelu_probe_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(x, y, dy, n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

static const char* lift_check(const LiftDims& d) {
  if (d.Z % 8) return "lift: Z must be a multiple of 8";
  if (d.Cin < 1 || d.Cin > 4) return "lift: supported input channel counts are 1..4";
  if (d.Tin < 1 || d.Tin > kLiftMaxTin) return "lift: 1 <= Tin <= 64";
  if (d.C > 64) return "lift: C <= 64";
  if (d.T * d.Tin + d.T + 2 * (d.C * d.Cin + d.C) > kLiftMaxW) return "lift: weights exceed shared memory budget";
  return nullptr;
}

const char* lift_fwd(const void* x, int x_is_bf16, const float* W1, const float* b1, const float* W2,
                     const float* b2, void* h, LiftDims d, int num_sms, cudaStream_t s) {
  if (const char* e = lift_check(d)) return e;
  const long long nitems = static_cast<long long>(d.B) * d.X * d.Y * (d.Z / 8);
  const int grid = grid_for(nitems, 256, num_sms, 4);
  const bool regs = d.Tin == 1;
#define DFNO_LIFT_FWD(T_, CIN_, R_) \
  lift_fwd_kernel<T_, CIN_, R_><<<grid, 256, 0, s>>>(static_cast<const T_*>(x), W1, b1, W2, b2, static_cast<__nv_bfloat16*>(h), d)
#define DFNO_LIFT_FWD_T(T_)                                                                         \
  switch (d.Cin) {                                                                                  \
    case 1: if (regs) DFNO_LIFT_FWD(T_, 1, true); else DFNO_LIFT_FWD(T_, 1, false); break;          \
    case 2: if (regs) DFNO_LIFT_FWD(T_, 2, true); else DFNO_LIFT_FWD(T_, 2, false); break;          \
    case 3: if (regs) DFNO_LIFT_FWD(T_, 3, true); else DFNO_LIFT_FWD(T_, 3, false); break;          \
    default: if (regs) DFNO_LIFT_FWD(T_, 4, true); else DFNO_LIFT_FWD(T_, 4, false); break;         \
  }
  if (x_is_bf16) { DFNO_LIFT_FWD_T(__nv_bfloat16) } else { DFNO_LIFT_FWD_T(float) }
#undef DFNO_LIFT_FWD_T
#undef DFNO_LIFT_FWD
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
                                LiftDims d, int grid, bool regs, cudaStream_t s) {
#define DFNO_LIFT_BWD2(T_, CIN_, R_)                                                                              \
  lift_bwd_kernel<T_, C, CIN_, R_><<<grid, 128, 0, s>>>(static_cast<const T_*>(x), W1, b1, W2, b2,                \
                                                        static_cast<const __nv_bfloat16*>(dh), gW1, gb1, gW2, gb2, d)
#define DFNO_LIFT_BWD(CIN_)                                                                    \
  if (x_is_bf16) { if (regs) DFNO_LIFT_BWD2(__nv_bfloat16, CIN_, true); else DFNO_LIFT_BWD2(__nv_bfloat16, CIN_, false); } \
  else { if (regs) DFNO_LIFT_BWD2(float, CIN_, true); else DFNO_LIFT_BWD2(float, CIN_, false); }
  switch (d.Cin) {
    case 1: DFNO_LIFT_BWD(1); break;
    case 2: DFNO_LIFT_BWD(2); break;
    case 3: DFNO_LIFT_BWD(3); break;
    case 4: DFNO_LIFT_BWD(4); break;
    default: return "lift_bwd: supported input channel counts are 1..4";
  }
#undef DFNO_LIFT_BWD
#undef DFNO_LIFT_BWD2
  return nullptr;
}

const char* lift_bwd(const void* x, int x_is_bf16, const float* W1, const float* b1, const float* W2,
                     const float* b2, const void* dh, float* gW1, float* gb1, float* gW2, float* gb2,
                     LiftDims d, int num_sms, cudaStream_t s) {
  if (const char* e = lift_check(d)) return e;
  const long long nitems = static_cast<long long>(d.B) * d.X * d.Y * (d.Z / 8);
  const int grid = grid_for(4 * nitems, 128, num_sms, 4);          // four lanes per item (channel quarters)
  const bool regs = d.Tin == 1;
  const char* err = nullptr;
  DFNO_DISPATCH_C(d.C, (err = lift_bwd_cin<kC>(x, x_is_bf16, W1, b1, W2, b2, dh, gW1, gb1, gW2, gb2, d, grid, regs, s)));
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

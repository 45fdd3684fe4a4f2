// loss.cu -- the elementwise part of the relative-L2 / MSE losses (reference loss.py:8-35)
// in two passes instead of the eight ATen kernels of the autograd graph:
//
//   forward    part[b]     += sum (y_hat - y)^2          part[B + b] += sum y^2        one read of both fields
//   backward   grad[b, :]   = (y_hat[b, :] - y[b, :]) * scale[b]                       one read, one write
//
// The cross-rank sum of `part` (2B floats) stays with the caller (peer-memory all-reduce of the fused engine, or
// NCCL); `scale` is a device vector so the step remains CUDA-graph capturable.
#include <cuda_runtime.h>
#include "kernels.h"

namespace dfno {
namespace {

constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads)
sq_partials_kernel(const float* __restrict__ yh, const float* __restrict__ y, float* __restrict__ part,
                   long long n_per_b, int B) {
  const int b = blockIdx.y;
  const float* a = yh + static_cast<long long>(b) * n_per_b;
  const float* r = y + static_cast<long long>(b) * n_per_b;
  float sd = 0.f, sr = 0.f;
  const long long n4 = n_per_b >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(r)) & 15) == 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (vec) {
    for (long long i = t0; i < n4; i += stride) {
      const float4 u = reinterpret_cast<const float4*>(a)[i], v = reinterpret_cast<const float4*>(r)[i];
      const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
      sd += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      sr += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (long long i = (n4 << 2) + t0; i < n_per_b; i += stride) { const float d = a[i] - r[i]; sd += d * d; sr += r[i] * r[i]; }
  } else {
    for (long long i = t0; i < n_per_b; i += stride) { const float d = a[i] - r[i]; sd += d * d; sr += r[i] * r[i]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
  }
  __shared__ float s_d[kLossThreads / 32], s_r[kLossThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_d[warp] = sd; s_r[warp] = sr; }
  __syncthreads();
  if (warp == 0) {
    sd = lane < kLossThreads / 32 ? s_d[lane] : 0.f;
    sr = lane < kLossThreads / 32 ? s_r[lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
    }
    if (lane == 0) { atomicAdd(&part[b], sd); atomicAdd(&part[B + b], sr); }
  }
}

__global__ void __launch_bounds__(kLossThreads)
scaled_diff_kernel(const float* __restrict__ yh, const float* __restrict__ y, const float* __restrict__ scale,
                   float* __restrict__ grad, long long n_per_b, int scale_per_b) {
  const int b = blockIdx.y;
  const float s = scale[scale_per_b ? b : 0];
  const long long off = static_cast<long long>(b) * n_per_b;
  const float* a = yh + off;
  const float* r = y + off;
  float* g = grad + off;
  const long long n4 = n_per_b >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (vec) {
    for (long long i = t0; i < n4; i += stride) {
      const float4 u = reinterpret_cast<const float4*>(a)[i], v = reinterpret_cast<const float4*>(r)[i];
      reinterpret_cast<float4*>(g)[i] = make_float4((u.x - v.x) * s, (u.y - v.y) * s, (u.z - v.z) * s, (u.w - v.w) * s);
    }
    for (long long i = (n4 << 2) + t0; i < n_per_b; i += stride) g[i] = (a[i] - r[i]) * s;
  } else {
    for (long long i = t0; i < n_per_b; i += stride) g[i] = (a[i] - r[i]) * s;
  }
}

int loss_grid(long long n_per_b, int B, int num_sms) {
  long long want = (n_per_b / 4 + kLossThreads - 1) / kLossThreads;
  const long long cap = (static_cast<long long>(num_sms) * 8 + B - 1) / B;
  if (want > cap) want = cap;
  return static_cast<int>(want < 1 ? 1 : want);
}

}  // namespace

// part (2B floats, zeroed by the caller) += per-sample sums of (y_hat - y)^2 and y^2
const char* sq_partials(const float* yh, const float* y, float* part, long long n_per_b, int B, int num_sms,
                        cudaStream_t s) {
  if (B < 1 || B > 65535 || n_per_b < 1) return "sq_partials: bad shape";
  sq_partials_kernel<<<dim3(loss_grid(n_per_b, B, num_sms), B), kLossThreads, 0, s>>>(yh, y, part, n_per_b, B);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// grad[b, :] = (y_hat[b, :] - y[b, :]) * scale[b]   (scale_per_b = 0: one scale for the whole tensor)
const char* scaled_diff(const float* yh, const float* y, const float* scale, float* grad, long long n_per_b, int B,
                        int scale_per_b, int num_sms, cudaStream_t s) {
  if (B < 1 || B > 65535 || n_per_b < 1) return "scaled_diff: bad shape";
  scaled_diff_kernel<<<dim3(loss_grid(n_per_b, B, num_sms), B), kLossThreads, 0, s>>>(yh, y, scale, grad, n_per_b,
                                                                                     scale_per_b);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

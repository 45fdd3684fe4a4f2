// TMA semantics probe: load one N-D tiled box (SWIZZLE_128B, bf16) into shared memory with
// cp.async.bulk.tensor.{2d,4d} and dump the raw shared-memory bytes.
//
// Purpose: the staged peer layout S1s[a, r_src, x, y_loc] could be consumed by the next GEMM without the
// local permutation if a 4-D box (y_loc*2, r_src, x, a) = (32, 2, 128, 1) lands in shared memory as the
// dense [128 rows][64 elements] K-major tile with the 128-byte swizzle applied to the *dense* offsets
// (DESIGN.md section 7, item 3).  This kernel lets a test check exactly that on hardware before the
// production GEMM is touched.  Not on any hot path.
#include <cuda_bf16.h>
#include <stdint.h>

#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace dfno {
namespace {

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(128)
tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, uint32_t bytes,
                 uint8_t* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_4d(smem, &tm, &bar, c0, c1, c2, c3);
  }
  mbar_wait(&bar, 0);
  for (uint32_t i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(smem + i);
}

}  // namespace

const char* tma_probe_4d(const void* src, const long long* dims, const long long* strides_elems, const int* box,
                         const int* coords, void* out, cudaStream_t s) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled unavailable";
  cuuint64_t d[4], st[3];
  cuuint32_t b[4], es[4] = {1, 1, 1, 1};
  uint64_t vol = 1;
  for (int i = 0; i < 4; ++i) {
    d[i] = static_cast<cuuint64_t>(dims[i]);
    b[i] = static_cast<cuuint32_t>(box[i]);
    vol *= b[i];
  }
  for (int i = 0; i < 3; ++i) st[i] = static_cast<cuuint64_t>(strides_elems[i]) * 2;   // bytes, dims 1..3
  if (b[0] * 2 > 128) return "inner box dimension exceeds the 128-byte swizzle span";
  const uint32_t bytes = static_cast<uint32_t>(vol * 2);
  if (bytes > 64 * 1024 || bytes % 16) return "box must be a multiple of 16 bytes and at most 64 KB";
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), d, st, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled(4-D) failed";
  cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  tma_probe_kernel<<<1, 128, bytes + 1024, s>>>(tm, coords[0], coords[1], coords[2], coords[3], bytes,
                                                static_cast<uint8_t*>(out));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace dfno

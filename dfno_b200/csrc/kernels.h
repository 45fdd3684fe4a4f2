// Launchers of the torch-free CUDA translation units (pointwise / spectral / optimizer / p2p).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dfno {
}  // namespace dfno

// Kernel launchers of libvrt (implemented in the .cu files, called by capi.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/vrt.h"

namespace vrt {

// K1: LSVO<D>::castRay over a ray buffer, reference node layout (lsvo_kernels.cu)
cudaError_t launch_lsvo_cast_ref(const uint2* nodes, int depth, int guard, const float* d_origin, const float* d_dir,
                                 float coef, float bias, uint64_t n, vrt_hit* d_out, unsigned long long* d_complexity,
                                 cudaStream_t stream);

}  // namespace vrt

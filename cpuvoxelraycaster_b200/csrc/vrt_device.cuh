// Device-side common definitions for libvrt (sm_100a).
//
// Numerical contract: every translation unit is compiled with --fmad=false, IEEE division and
// square root (nvcc defaults -prec-div=true -prec-sqrt=true, -ftz=false), because the reference's
// canonical semantics are "no FMA contraction" (CMakeLists.txt:19-22,28 — plain Release flags,
// no -march).  Hit records are then bit-identical to the CPU reference.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/vrt.h"

namespace vrt {

constexpr int kSvoMaxDepth = 23;                        // lsvo.hpp:37
constexpr float kEps = 1.0f / float(1 << kSvoMaxDepth);  // lsvo.hpp:40

struct float3x { float x, y, z; };

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return (ax * bx + ay * by) + az * bz;                // glm::dot: (x + y) + z
}
// glm::normalize = v * (1 / sqrt(dot(v, v)))
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
    const float inv = 1.0f / sqrtf(dot3(x, y, z, x, y, z));
    x *= inv; y *= inv; z *= inv;
}
// frac (utils.cpp:60-64) = modf fractional part
__device__ __forceinline__ float fracf(float f) { return f - truncf(f); }

// ---- Philox4x32-10 on getRand's 100-level lattice (utils.cpp:77-81) -------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// getRand(lo, hi) = lo + (hi - lo) * (float(r % 100) / 100.0f)
__device__ __forceinline__ float lattice(uint32_t word, float lo, float hi) {
    const float rv = float(word % 100u) / 100.0f;
    return lo + (hi - lo) * rv;
}

}  // namespace vrt

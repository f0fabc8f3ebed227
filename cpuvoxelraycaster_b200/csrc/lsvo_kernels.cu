// Batched LSVO<D>::castRay (reference include/lsvo.hpp:33-172) — kernel K1.
#include "lsvo_step.cuh"
#include "kernels.h"

namespace vrt {

__device__ __forceinline__ void store_hit(vrt_hit* out, const LsvoResult& r, const LsvoHit& h, int depth) {
    float4* q = reinterpret_cast<float4*>(out);
    if (!r.hit) {
        q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        q[1] = make_float4(0.f, 0.f, 0.f, __uint_as_float(r.complexity));
        q[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float S = float(1 << depth);
    q[0] = make_float4(h.pos[0], h.pos[1], h.pos[2], h.distance);
    q[1] = make_float4(h.normal[0], h.normal[1], h.normal[2], __uint_as_float(r.complexity));
    q[2] = make_float4(h.uv[0], h.uv[1], __uint_as_float(VRT_HIT_FLAG_HIT), __int_as_float(r.scale));
    q[3] = make_float4(__int_as_float(int((h.corner[0] - 1.0f) * S)), __int_as_float(int((h.corner[1] - 1.0f) * S)),
                       __int_as_float(int((h.corner[2] - 1.0f) * S)), __uint_as_float(r.face));
}

// v0: one thread per ray, traversal stack in shared memory.
template <typename Nodes>
__global__ void __launch_bounds__(128) lsvo_cast_kernel(Nodes nodes, int depth, int guard, const float* __restrict__ origin,
                                                        const float* __restrict__ dir, float coef, float bias, uint64_t n,
                                                        vrt_hit* __restrict__ out, unsigned long long* __restrict__ total_complexity) {
    extern __shared__ uint2 smem[];
    Stack64<128> stack{smem + threadIdx.x};
    nodes.slots = pin(nodes.slots);
    guard = pin(guard);
    const int depth_offset = pin(kSvoMaxDepth - depth);

    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t iters = 0u;
    if (i < n) {
        const float ox = origin[3 * i], oy = origin[3 * i + 1], oz = origin[3 * i + 2];
        const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        LsvoResult r;
        lsvo_cast_ray(nodes, stack, depth_offset, guard, ox, oy, oz, dx, dy, dz, coef, bias, r);
        LsvoHit h;
        if (r.hit) lsvo_finish(r, ox, oy, oz, depth, h);
        store_hit(out + i, r, h, depth);
        iters = r.complexity;
    }
    // Σ complexity: warp reduce, one atomic per warp
    for (int o = 16; o > 0; o >>= 1) iters += __shfl_xor_sync(0xffffffffu, iters, o);
    if ((threadIdx.x & 31) == 0 && iters) atomicAdd(total_complexity, (unsigned long long)iters);
}

// K1b: the same kernel on Trav2 (bookkeeping off the ALU pipe, cone test compiled out for coef = bias = 0)
template <typename Nodes, bool kCone>
__global__ void __launch_bounds__(128) lsvo_cast2_kernel(Nodes nodes, int depth, int guard, const float* __restrict__ origin,
                                                         const float* __restrict__ dir, float coef, float bias, uint64_t n,
                                                         vrt_hit* __restrict__ out, unsigned long long* __restrict__ total_complexity) {
    extern __shared__ uint2 smem[];
    nodes.slots = pin(nodes.slots);
    const float guard_sf = keep_in_register(guard_scale_f(guard), smem + threadIdx.x);
    guard = keep_in_register(guard, smem + threadIdx.x);
    Stack64s<128> stack = Stack64s<128>::make(smem + threadIdx.x, kSvoMaxDepth - depth);
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t iters = 0u;
    if (i < n) {
        const float ox = origin[3 * i], oy = origin[3 * i + 1], oz = origin[3 * i + 2];
        const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        LsvoResult r;
        lsvo_cast_ray2<kCone>(nodes, stack, guard, guard_sf, ox, oy, oz, dx, dy, dz, coef, bias, r);
        LsvoHit h;
        if (r.hit) lsvo_finish(r, ox, oy, oz, depth, h);
        store_hit(out + i, r, h, depth);
        iters = r.complexity;
    }
    for (int o = 16; o > 0; o >>= 1) iters += __shfl_xor_sync(0xffffffffu, iters, o);
    if ((threadIdx.x & 31) == 0 && iters) atomicAdd(total_complexity, (unsigned long long)iters);
}

cudaError_t launch_lsvo_cast2(const uint2* nodes, int depth, int guard, const float* d_origin, const float* d_dir, float coef, float bias,
                              uint64_t n, vrt_hit* d_out, unsigned long long* d_complexity, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int block = 128;
    const size_t smem = size_t(depth + 1) * block * 8;
    const uint64_t grid = (n + block - 1) / block;
    if (coef == 0.0f && bias == 0.0f)
        lsvo_cast2_kernel<RefNodes, false><<<unsigned(grid), block, smem, stream>>>(RefNodes{nodes}, depth, guard, d_origin, d_dir, coef, bias, n, d_out, d_complexity);
    else
        lsvo_cast2_kernel<RefNodes, true><<<unsigned(grid), block, smem, stream>>>(RefNodes{nodes}, depth, guard, d_origin, d_dir, coef, bias, n, d_out, d_complexity);
    return cudaGetLastError();
}

cudaError_t launch_lsvo_cast_ref(const uint2* nodes, bool compact, int depth, int guard, const float* d_origin, const float* d_dir,
                                 float coef, float bias, uint64_t n, vrt_hit* d_out, unsigned long long* d_complexity,
                                 cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int block = 128;
    const size_t smem = size_t(depth + 1) * block * 8;
    const uint64_t grid = (n + block - 1) / block;
    if (compact)
        lsvo_cast_kernel<CompactNodes><<<unsigned(grid), block, smem, stream>>>(CompactNodes{nodes}, depth, guard, d_origin, d_dir, coef,
                                                                               bias, n, d_out, d_complexity);
    else
        lsvo_cast_kernel<RefNodes><<<unsigned(grid), block, smem, stream>>>(RefNodes{nodes}, depth, guard, d_origin, d_dir, coef, bias, n,
                                                                           d_out, d_complexity);
    return cudaGetLastError();
}

}  // namespace vrt

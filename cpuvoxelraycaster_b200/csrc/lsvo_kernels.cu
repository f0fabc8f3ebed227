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

// K1b: the same kernel on Trav2 (bookkeeping off the ALU pipe, cone test compiled out for coef = bias = 0).  Grid-stride, so that a
// gated launch (below) can use a bounded grid; with one CTA per 128 rays the loop body runs once.
// gate != nullptr: the kernel runs only if *gate == want (the automatic choice between K1b and K1p, made on the device).
// kGuard = false: the loop guard cannot bind for this scene and is compiled out (lsvo_step.cuh, Trav2).
template <typename Nodes, bool kCone, bool kGuard>
__global__ void __launch_bounds__(128) lsvo_cast2_kernel(Nodes nodes, int depth, int guard, const float* __restrict__ origin,
                                                         const float* __restrict__ dir, float coef, float bias, uint64_t n,
                                                         vrt_hit* __restrict__ out, unsigned long long* __restrict__ total_complexity,
                                                         const unsigned long long* __restrict__ gate, unsigned long long want) {
    extern __shared__ uint2 smem[];
    if (gate && *gate != want) return;
    nodes.slots = pin(nodes.slots);
    const float guard_sf = keep_in_register(guard_scale_f(guard), smem + threadIdx.x);
    guard = keep_in_register(guard, smem + threadIdx.x);
    Stack64s<128> stack = Stack64s<128>::make(smem + threadIdx.x, kSvoMaxDepth - depth);
    uint32_t iters = 0u;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const float ox = origin[3 * i], oy = origin[3 * i + 1], oz = origin[3 * i + 2];
        const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        LsvoResult r;
        lsvo_cast_ray2<kCone, false, kGuard>(nodes, stack, guard, guard_sf, ox, oy, oz, dx, dy, dz, coef, bias, r);
        LsvoHit h;
        if (r.hit) lsvo_finish(r, ox, oy, oz, depth, h);
        store_hit(out + i, r, h, depth);
        iters += r.complexity;
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) iters += __shfl_xor_sync(0xffffffffu, iters, o);
    if ((threadIdx.x & 31) == 0 && iters) atomicAdd(total_complexity, (unsigned long long)iters);
}

// Are the rays of a batch coherent the way the reference's are (32 consecutive rays = neighbouring pixels of one camera)?  32 groups
// of 32 consecutive rays, spread over the batch, are looked at: a group is coherent when its directions (normalised) and its origins
// stay close to those of the group's first ray.  *gate = 1 when at least 3/4 of the groups are: one thread per ray (K1b) then beats
// the regenerating persistent kernel (8.6 vs 6.5 Grays/s on 1080p camera rays; 4.5 vs 6.5 on random rays — profiles/r02_probe_bounds.txt).
// The choice only decides which kernel runs; both return the same bytes.
__global__ void __launch_bounds__(1024) classify_rays_kernel(const float* __restrict__ origin, const float* __restrict__ dir, uint64_t n,
                                                            unsigned long long* __restrict__ gate) {
    __shared__ int votes;
    if (threadIdx.x == 0) votes = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, group = threadIdx.x >> 5;
    const uint64_t groups = n / 32;                        // n >= 1024 (the launcher's rule)
    const uint64_t i = (groups * uint64_t(group) / 32) * 32 + uint64_t(lane);
    float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
    const float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-30f));
    dx *= inv; dy *= inv; dz *= inv;
    const float ox = origin[3 * i], oy = origin[3 * i + 1], oz = origin[3 * i + 2];
    float spread = fmaxf(fmaxf(fabsf(dx - __shfl_sync(0xffffffffu, dx, 0)), fabsf(dy - __shfl_sync(0xffffffffu, dy, 0))),
                         fabsf(dz - __shfl_sync(0xffffffffu, dz, 0)));
    float apart = fmaxf(fmaxf(fabsf(ox - __shfl_sync(0xffffffffu, ox, 0)), fabsf(oy - __shfl_sync(0xffffffffu, oy, 0))),
                        fabsf(oz - __shfl_sync(0xffffffffu, oz, 0)));
    if (!(spread == spread) || !(apart == apart)) spread = 1e9f;          // NaN rays: not coherent
    for (int o = 16; o > 0; o >>= 1) {
        spread = fmaxf(spread, __shfl_xor_sync(0xffffffffu, spread, o));
        apart = fmaxf(apart, __shfl_xor_sync(0xffffffffu, apart, o));
    }
    if (lane == 0 && spread < 0.25f && apart < 0.05f) atomicAdd(&votes, 1);
    __syncthreads();
    if (threadIdx.x == 0) *gate = votes >= 24 ? 1ull : 0ull;
}

cudaError_t launch_classify_rays(const float* d_origin, const float* d_dir, uint64_t n, unsigned long long* d_gate, cudaStream_t stream) {
    classify_rays_kernel<<<1, 1024, 0, stream>>>(d_origin, d_dir, n, d_gate);
    return cudaGetLastError();
}

cudaError_t launch_lsvo_cast2(const uint2* nodes, int depth, int guard, const float* d_origin, const float* d_dir, float coef, float bias,
                              uint64_t n, vrt_hit* d_out, unsigned long long* d_complexity, cudaStream_t stream,
                              const unsigned long long* d_gate, unsigned long long want) {
    if (n == 0) return cudaSuccess;
    const int block = 128;
    const size_t smem = size_t(depth + 1) * block * 8;
    uint64_t grid = (n + block - 1) / block;
    if (d_gate) {                                          // bounded grid: a launch that turns out not to be wanted must cost nothing
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const uint64_t cap = uint64_t(sms) * 16 * 8;       // 8 rounds of 16 CTAs per SM: dynamic enough to stay balanced
        if (grid > cap) grid = cap;
    }
    auto launch = [&](auto kernel) {
        kernel<<<unsigned(grid), block, smem, stream>>>(RefNodes{nodes}, depth, guard, d_origin, d_dir, coef, bias, n, d_out, d_complexity, d_gate, want);
    };
    const bool cone = !(coef == 0.0f && bias == 0.0f), guarded = guard >= kSvoMaxDepth - depth;   // can the loop guard ever bind?
    if (cone) { if (guarded) launch(lsvo_cast2_kernel<RefNodes, true, true>); else launch(lsvo_cast2_kernel<RefNodes, true, false>); }
    else { if (guarded) launch(lsvo_cast2_kernel<RefNodes, false, true>); else launch(lsvo_cast2_kernel<RefNodes, false, false>); }
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

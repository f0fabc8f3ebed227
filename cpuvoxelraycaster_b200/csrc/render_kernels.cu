// Frame rendering — kernels K0 (ray generation) + K4 (shading and secondary rays), fused; and the resolve.
//
// Replaces the swarm lambda src/main.cpp:139-154, Camera::getRay (camera_controller.hpp:34-54),
// RayCaster::renderRay/castRay/getGlobalIllumination and the texture lookup (raycaster.hpp:67-240),
// samples_to_image (raycaster.hpp:94-103).  The per-sample arithmetic is in render_chain.cuh.
//
// Structure: each lane owns one pixel (8x4 pixel tile per warp) and walks a run of its samples; a sample is
// a chain of up to six rays.  The chain is a small state machine around ONE inlined copy of the traversal loop,
// so lanes at different chain stages still execute the traversal converged.  This one-lane-per-pixel kernel is
// the default: measured against persistent/regenerating and shared-memory-state variants in profiles/.
#include "lsvo_step.cuh"
#include "render_chain.cuh"

namespace vrt {

// 4 CTAs per SM (128 registers): more resident warps at fewer registers measured slower (profiles/r01_summary.md)
#ifndef VRT_K4_MIN_CTAS
#define VRT_K4_MIN_CTAS 4
#endif
// kLive: the interactive-loop extras (checkerboard pixel mapping, focal length read from the autofocus kernel's output).
// Compiled out of the plain instantiation so that they cost the many-sample frames nothing (register allocation of the
// traversal loop is sensitive to every extra live value: 73.9 vs 75.6 ms on cfg 4).
template <typename Nodes, bool kLive>
__global__ void __launch_bounds__(128, VRT_K4_MIN_CTAS) render_accumulate_kernel(Nodes nodes, RenderLaunch L, uint32_t* __restrict__ accum,
                                                                unsigned long long* __restrict__ counters) {
    extern __shared__ uint2 smem[];
    Stack64<128> stack{smem + threadIdx.x};
    nodes.slots = pin(nodes.slots);
    const int guard = pin(L.guard);
    const int depth_offset = pin(kSvoMaxDepth - L.depth);

    // 8x4 pixel tile per warp, 4 tiles side by side per block: coherent primary rays share nodes
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // checkerboard frames (main.cpp:137,143) render every other pixel of a row: the tile's columns are then the
    // indices of the rendered pixels, so that no lane idles
    const int checker = kLive ? L.checker : 0;
    const int columns = checker ? (L.width + 1) / 2 : L.width;
    const int tiles_x = (columns + 31) / 32;
    // blockIdx.x = chunk * tiles + tile: the pixel's samples are cut into L.spp_chunks runs handled by different CTAs
    // (more, shorter CTAs: keeps the tail short when a GPU owns only a slice of the frame; sums stay exact — integers)
    const int tiles_y = int(gridDim.x) / (tiles_x * L.spp_chunks);
    const int chunk = int(blockIdx.x) / (tiles_x * tiles_y), tile = int(blockIdx.x) - chunk * tiles_x * tiles_y;
    const int bx = tile % tiles_x, by = tile / tiles_x;
    const int s_begin = (chunk * L.spp) / L.spp_chunks, s_end = ((chunk + 1) * L.spp) / L.spp_chunks;
    // Lane → (pixel, sample) mapping inside the warp's 8x4 pixel tile.  Q = L.samples_per_warp lanes share a pixel and
    // take consecutive samples; the tile's 32 pixels are walked in Q groups of P = 32/Q pixels.  Q = 1 is the plain
    // one-lane-per-pixel mapping.  With many samples per pixel, lanes of one pixel cast nearly identical primary and
    // shadow rays (they differ by the lens jitter only), which keeps the warp converged far better than 32 neighbouring
    // pixels do.  Sums are integers, so the mapping cannot change the frame.
    const int Q = L.samples_per_warp, P = 32 / Q;
    const int sub = lane & (Q - 1);                        // which of the pixel's Q concurrent samples
    const int runs = (s_end - s_begin) / Q;                // launcher guarantees divisibility

    // ray statistics per class: private shared-memory counters (12 registers less across the traversal loop)
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + (L.depth + 1) * 128) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 12; ++k) cnt[k * 128] = 0u;
    const float SCALE = 1.0f / float(1 << L.depth);                           // raycaster.hpp:123-124 / main.cpp:82
    const float n_norm = SCALE * 0.0078125f * 2.0f;                           // raycaster.hpp:171-172
    const float aspect = float(L.width) / float(L.height);                    // main.cpp:133
    const float focal_length = (kLive && L.focal) ? __ldg(L.focal) : L.cam.focal_length;   // main.cpp:114-121 on the device

    for (int g = 0; g < Q; ++g) {
        const int j = g * P + lane / Q;                    // pixel index inside the 8x4 tile
        // 4-row tiles are dealt round-robin to tile_step owners (multi-GPU row partition, balanced sky/terrain)
        const int y = L.row_begin + (by * L.tile_step + L.tile_index) * 4 + (j >> 3);
        int x = bx * 32 + warp * 8 + (j & 7);
        if (checker) x = 2 * x + checker_x_parity(checker, L.checker_area_height, y);
        const bool active = x < L.width && y < L.row_end;
        const uint32_t pixel = uint32_t(y) * uint32_t(L.width) + uint32_t(x);
        uint32_t sum_r = 0, sum_g = 0, sum_b = 0;
        if (active) {
            const float lens_x = float(x) / float(L.height) - aspect * 0.5f;  // main.cpp:145
            const float lens_y = float(y) / float(L.height) - 0.5f;           // main.cpp:146
            for (int k = 0; k < runs; ++k) {
                const uint32_t sample = uint32_t(L.sample_offset + s_begin + k * Q + sub);
                ChainState c;
                NextRay nr;
                chain_begin(L, c, pixel, sample, lens_x, lens_y, SCALE, focal_length, nr);
                int stage = kPrimary;
                while (stage != kDone) {
                    LsvoResult r;
                    lsvo_cast_ray(nodes, stack, depth_offset, guard, nr.ox, nr.oy, nr.oz, nr.dx, nr.dy, nr.dz, nr.coef, 0.0f, r);
                    cnt[stage * 128] += 1u;
                    cnt[(6 + stage) * 128] += r.complexity;
                    LsvoHit h;
                    if (r.hit) lsvo_finish(r, nr.ox, nr.oy, nr.oz, L.depth, h);
                    stage = chain_advance(L, c, stage, r, h, pixel, sample, SCALE, n_norm, nr);
                }
                chain_colour(L, c, sum_r, sum_g, sum_b);
            }
        }
        // the Q lanes of a pixel add up their sums; lane sub == 0 commits them
        for (int o = Q >> 1; o > 0; o >>= 1) {
            sum_r += __shfl_xor_sync(0xffffffffu, sum_r, o);
            sum_g += __shfl_xor_sync(0xffffffffu, sum_g, o);
            sum_b += __shfl_xor_sync(0xffffffffu, sum_b, o);
        }
        if (active && sub == 0) {
            uint4* a = reinterpret_cast<uint4*>(accum) + pixel;               // Sample, raycaster.hpp:18-24,87-90
            if (L.spp_chunks == 1) {
                uint4 v = *a;
                v.x += sum_r; v.y += sum_g; v.z += sum_b; v.w += uint32_t(L.spp);
                *a = v;
            } else {
                uint32_t* w = reinterpret_cast<uint32_t*>(a);
                atomicAdd(w, sum_r); atomicAdd(w + 1, sum_g); atomicAdd(w + 2, sum_b); atomicAdd(w + 3, uint32_t(s_end - s_begin));
            }
        }
    }

    // statistics: rays and Σ complexity per ray class (warp reduce, one atomic per warp and class)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        uint32_t a = cnt[k * 128], b = cnt[(6 + k) * 128];
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0 && a) {
            atomicAdd(counters + k, (unsigned long long)a);
            atomicAdd(counters + 6 + k, (unsigned long long)b);
        }
    }
}

// Camera::getClosestPoint (camera_controller.hpp:56-60) and the focal-length rule of main.cpp:114-121, one thread.
template <typename Nodes>
__global__ void __launch_bounds__(128) autofocus_kernel(Nodes nodes, int depth, int guard, vrt_camera cam, float* __restrict__ focal) {
    extern __shared__ uint2 smem[];
    if (threadIdx.x != 0) return;
    Stack64<128> stack{smem};
    const float scale = 1.0f / float(1 << depth);
    const float ox = cam.position[0] * scale + 1.0f, oy = cam.position[1] * scale + 1.0f, oz = cam.position[2] * scale + 1.0f;
    float dx, dy, dz;
    view_to_world(cam.rot_mat, 0.0f, 0.0f, 1.0f, dx, dy, dz);                 // camera_vec, camera_controller.hpp:31
    LsvoResult r;
    lsvo_cast_ray(nodes, stack, kSvoMaxDepth - depth, guard, ox, oy, oz, dx, dy, dz, 0.0f, 0.0f, r);
    *focal = r.hit ? r.t_min * float(1 << depth) : 100.0f;
}

// samples_to_image (raycaster.hpp:94-103) or the 0.4/0.6 temporal blend of renderRay (:79-85).
__global__ void resolve_kernel(const uint32_t* __restrict__ accum, uint8_t* __restrict__ rgba, int width, int row_begin,
                               int row_end, int use_samples, int tile_step, int tile_index) {
    const uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;       // index inside the owned rows
    const int x = int(j % uint64_t(width));
    const int row = int(j / uint64_t(width));                                 // owned-row ordinal
    const int y = row_begin + ((row >> 2) * tile_step + tile_index) * 4 + (row & 3);
    if (y >= row_end) return;
    const uint64_t i = uint64_t(y) * width + x;
    const uint4 a = reinterpret_cast<const uint4*>(accum)[i];
    uchar4* out = reinterpret_cast<uchar4*>(rgba) + i;
    if (use_samples) {
        // double division then uint8 conversion == integer division for sums < 2^32
        const uint32_t n = a.w ? a.w : 1u;
        *out = make_uchar4(uint8_t(a.x / n), uint8_t(a.y / n), uint8_t(a.z / n), 255);
    } else {
        if (a.w == 0u) return;                                                // not rendered this frame (checkerboard)
        const uchar4 old = *out;
        // the accumulator holds exactly one sample in this mode
        const uint8_t nr = mul_u8(uint8_t(a.x), 1.0f - 0.4f), ng = mul_u8(uint8_t(a.y), 1.0f - 0.4f), nb = mul_u8(uint8_t(a.z), 1.0f - 0.4f);
        const int r = min(255, int(mul_u8(old.x, 0.4f)) + int(nr));          // add(sf::Color&, const sf::Color&), utils.cpp:35-40
        const int g = min(255, int(mul_u8(old.y, 0.4f)) + int(ng));
        const int b = min(255, int(mul_u8(old.z, 0.4f)) + int(nb));
        *out = make_uchar4(uint8_t(r), uint8_t(g), uint8_t(b), 255);
    }
}

cudaError_t launch_render_accumulate_ref(const uint2* nodes, bool compact, const RenderLaunch& L, uint32_t* d_accum,
                                         unsigned long long* d_counters, cudaStream_t stream) {
    const int rows = L.row_end - L.row_begin;
    if (rows <= 0 || L.width <= 0 || L.spp <= 0) return cudaSuccess;
    const int block = 128;
    const int columns = L.checker ? (L.width + 1) / 2 : L.width;
    const int tiles_x = (columns + 31) / 32, tiles_y = ((rows + 3) / 4 + L.tile_step - 1 - L.tile_index) / L.tile_step;
    if (tiles_y <= 0) return cudaSuccess;
    const size_t smem = size_t(L.depth + 1) * block * 8 + 12 * block * sizeof(uint32_t);   // stacks + statistics
    // Sample runs: a power of two, enough for >= 28 waves of CTAs (4 CTAs x 148 SMs resident) so the last wave is a
    // small part of the launch even when a GPU owns 1/8 of the frame, but runs of >= 8 samples so that 8+ lanes can
    // share a pixel.  tools/probe_slice.py: whole frame 74.4 ms at (2 runs, 32 lanes/pixel) vs 77.7 ms at (4, 1);
    // 1/8 slice 9.60 ms at (8, 8) vs 9.92 ms at (29, 1).
    RenderLaunch Lc = L;
    const long tiles = long(tiles_x) * tiles_y;
    long chunks = 1;
    while (chunks * tiles < 28L * 4 * 148 && chunks * 2 * 8 <= L.spp) chunks *= 2;
    if (L.spp_chunks > 0) chunks = L.spp_chunks < L.spp ? L.spp_chunks : L.spp;      // explicit override
    Lc.spp_chunks = int(chunks);
    // lanes per pixel: the largest power of two (<= 32) that divides every chunk's sample count
    int q = 32;
    for (long c = 0; c < chunks; ++c) {
        const long ns = ((c + 1) * L.spp) / chunks - (c * L.spp) / chunks;
        while (q > 1 && ns % q) q >>= 1;
    }
    if (L.samples_per_warp > 0 && L.samples_per_warp < q) q = L.samples_per_warp;       // explicit override (power of two)
    Lc.samples_per_warp = q;
    const unsigned grid = unsigned(tiles * chunks);
    const bool live = L.checker != 0 || L.focal != nullptr;
    if (compact && live) render_accumulate_kernel<CompactNodes, true><<<grid, block, smem, stream>>>(CompactNodes{nodes}, Lc, d_accum, d_counters);
    else if (compact) render_accumulate_kernel<CompactNodes, false><<<grid, block, smem, stream>>>(CompactNodes{nodes}, Lc, d_accum, d_counters);
    else if (live) render_accumulate_kernel<RefNodes, true><<<grid, block, smem, stream>>>(RefNodes{nodes}, Lc, d_accum, d_counters);
    else render_accumulate_kernel<RefNodes, false><<<grid, block, smem, stream>>>(RefNodes{nodes}, Lc, d_accum, d_counters);
    return cudaGetLastError();
}

cudaError_t launch_autofocus(const uint2* nodes, bool compact, int depth, int guard, const vrt_camera& cam, float* d_focal,
                             cudaStream_t stream) {
    const size_t smem = size_t(depth + 1) * 128 * 8;
    if (compact) autofocus_kernel<CompactNodes><<<1, 128, smem, stream>>>(CompactNodes{nodes}, depth, guard, cam, d_focal);
    else autofocus_kernel<RefNodes><<<1, 128, smem, stream>>>(RefNodes{nodes}, depth, guard, cam, d_focal);
    return cudaGetLastError();
}

cudaError_t launch_resolve(const uint32_t* d_accum, uint8_t* d_rgba, int width, int row_begin, int row_end, int use_samples,
                           int tile_step, int tile_index, cudaStream_t stream) {
    const int tiles = ((row_end - row_begin + 3) / 4 + tile_step - 1 - tile_index) / tile_step;
    const uint64_t n = uint64_t(tiles > 0 ? tiles : 0) * 4 * width;
    if (!n) return cudaSuccess;
    resolve_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(d_accum, d_rgba, width, row_begin, row_end, use_samples, tile_step,
                                                                 tile_index);
    return cudaGetLastError();
}

}  // namespace vrt

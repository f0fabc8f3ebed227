// Frame rendering — kernels K0 (ray generation) + K4 (shading and secondary rays), fused.
//
// Replaces the swarm lambda src/main.cpp:139-154, Camera::getRay (camera_controller.hpp:34-54),
// RayCaster::renderRay/castRay/getGlobalIllumination and the texture lookup (raycaster.hpp:67-240),
// samples_to_image (raycaster.hpp:94-103).  Arithmetic follows those lines op for op (fp32, no
// contraction) so the 8-bit colour of every sample equals the CPU restatement's; the racy global
// xorshf96 (utils.cpp:11-25) is replaced by Philox4x32-10 keyed by (pixel, sample, dimension) on the
// same 100-level lattice, which makes stochastic frames reproducible and row/spp partitions exact.
//
// Structure: each lane owns one pixel and walks its samples; a sample is a chain of up to six rays
// (primary, sun shadow, GI, GI shadow, second bounce, its shadow).  The chain is a small state
// machine around ONE inlined copy of the traversal loop, so lanes at different chain stages still
// execute the traversal converged.
#include "lsvo_step.cuh"
#include "kernels.h"

namespace vrt {

enum Stage : int { kPrimary = 0, kShadow = 1, kGi0 = 2, kGi0Shadow = 3, kGi1 = 4, kGi1Shadow = 5, kDone = 6 };

__device__ __forceinline__ uint8_t mul_u8(uint8_t c, float f) {       // mult(sf::Color&, float), utils.cpp:43-48
    return uint8_t(fminf(255.0f, float(c) * f));
}

// v * rot_mat (camera_controller.hpp:51-54); m is column major
__device__ __forceinline__ void view_to_world(const float* m, float vx, float vy, float vz, float& x, float& y, float& z) {
    x = (m[0] * vx + m[1] * vy) + m[2] * vz;
    y = (m[3] * vx + m[4] * vy) + m[5] * vz;
    z = (m[6] * vx + m[7] * vy) + m[8] * vz;
}

template <typename Nodes>
__global__ void __launch_bounds__(128) render_accumulate_kernel(Nodes nodes, RenderLaunch L, uint32_t* __restrict__ accum,
                                                                unsigned long long* __restrict__ counters) {
    extern __shared__ uint2 smem[];
    Stack64<128> stack{smem + threadIdx.x};
    nodes.slots = pin(nodes.slots);
    const int guard = pin(L.guard);
    const int depth_offset = pin(kSvoMaxDepth - L.depth);

    // 8x4 pixel tile per warp, 4 tiles side by side per block: coherent primary rays share nodes
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_x = (L.width + 31) / 32;
    // blockIdx.x = chunk * tiles + tile: the pixel's samples are cut into L.spp_chunks runs handled by different CTAs
    // (more, shorter CTAs: keeps the tail short when a GPU owns only a slice of the frame; sums stay exact — integers)
    const int tiles_y = int(gridDim.x) / (tiles_x * L.spp_chunks);
    const int chunk = int(blockIdx.x) / (tiles_x * tiles_y), tile = int(blockIdx.x) - chunk * tiles_x * tiles_y;
    const int bx = tile % tiles_x, by = tile / tiles_x;
    const int s_begin = (chunk * L.spp) / L.spp_chunks, s_end = ((chunk + 1) * L.spp) / L.spp_chunks;
    const int x = bx * 32 + warp * 8 + (lane & 7);
    // 4-row tiles are dealt round-robin to tile_step owners (multi-GPU row partition, balanced sky/terrain)
    const int y = L.row_begin + (by * L.tile_step + L.tile_index) * 4 + (lane >> 3);
    const bool active = x < L.width && y < L.row_end;

    uint32_t n_rays[6] = {0, 0, 0, 0, 0, 0};
    uint32_t n_iter[6] = {0, 0, 0, 0, 0, 0};

    if (active) {
        const float SCALE = 1.0f / float(1 << L.depth);                       // raycaster.hpp:123-124 / main.cpp:82
        const float n_norm = SCALE * 0.0078125f * 2.0f;                       // raycaster.hpp:171-172
        const float aspect = float(L.width) / float(L.height);                // main.cpp:133
        const float lens_x = float(x) / float(L.height) - aspect * 0.5f;      // main.cpp:145
        const float lens_y = float(y) / float(L.height) - 0.5f;               // main.cpp:146
        const uint32_t pixel = uint32_t(y) * uint32_t(L.width) + uint32_t(x);
        uint32_t sum_r = 0, sum_g = 0, sum_b = 0;

        for (int s = s_begin; s < s_end; ++s) {
            const uint32_t sample = uint32_t(L.sample_offset + s);
            const uint4 rnd0 = philox4x32_10(pixel, sample, 0u, 0u, L.seed_lo, L.seed_hi);

            // chain state
            float ox, oy, oz, dx, dy, dz, coef;
            float nx = 0.f, ny = 0.f, nz = 0.f;          // primary normal
            float light = 0.f;
            float dot_gi0 = 0.f, dot_gi1 = 0.f, irr0 = 0.f, irr1 = 0.f;
            float gnx = 0.f, gny = 0.f, gnz = 0.f;       // normal of the current GI hit
            float gpx = 0.f, gpy = 0.f, gpz = 0.f;       // position of the current GI hit
            float tlx = 0.f, tly = 0.f, tlz = 0.f;       // unit vector to the light of the pending shadow ray
            uint8_t tex_r = 0, tex_g = 0, tex_b = 0;
            bool have_hit = false, gi0_hit = false, gi1_hit = false;

            {   // Camera::getRay, camera_controller.hpp:34-49
                const float u0 = lattice(rnd0.x, -0.5f, 0.5f), u1 = lattice(rnd0.y, -0.5f, 0.5f);
                float fx = lens_x, fy = lens_y, fz = L.cam.fov;
                normalize3(fx, fy, fz);
                fx *= L.cam.focal_length; fy *= L.cam.focal_length; fz *= L.cam.focal_length;
                const float rx = L.cam.aperture * u0, ry = L.cam.aperture * u1, rz = L.cam.aperture * 0.0f;
                float qx = fx - rx, qy = fy - ry, qz = fz - rz;
                normalize3(qx, qy, qz);
                float wx, wy, wz;
                view_to_world(L.cam.rot_mat, qx, qy, qz, dx, dy, dz);
                view_to_world(L.cam.rot_mat, rx, ry, rz, wx, wy, wz);
                ox = (L.cam.position[0] + wx) * SCALE + 1.0f;                  // main.cpp:149
                oy = (L.cam.position[1] + wy) * SCALE + 1.0f;
                oz = (L.cam.position[2] + wz) * SCALE + 1.0f;
                coef = 0.0f;
            }

            int stage = kPrimary;
            while (stage != kDone) {
                LsvoResult r;
                lsvo_cast_ray(nodes, stack, depth_offset, guard, ox, oy, oz, dx, dy, dz, coef, 0.0f, r);
#pragma unroll
                for (int k = 0; k < 6; ++k) {                                  // predicated: keeps the counters in registers
                    n_rays[k] += (stage == k) ? 1u : 0u;
                    n_iter[k] += (stage == k) ? r.complexity : 0u;
                }
                LsvoHit h;
                if (r.hit) lsvo_finish(r, ox, oy, oz, L.depth, h);

                int next = kDone;
                switch (stage) {
                    case kPrimary: {                                           // raycaster.hpp:131-145
                        if (!r.hit) break;
                        have_hit = true;
                        nx = h.normal[0]; ny = h.normal[1]; nz = h.normal[2];
                        const uint8_t* tex = (ny != 0.0f) ? L.tex_top : L.tex_side;    // :211-215
                        const float u = fminf(fmaxf(h.uv[0], 0.0f), 1.0f), v = fminf(fmaxf(h.uv[1], 0.0f), 1.0f);   // :237-238
                        const uint32_t tx = uint32_t(16.0f * u), ty = uint32_t(16.0f * v);                          // :239
                        const uint8_t* texel = tex + 3u * (ty * 16u + tx);
                        tex_r = __ldg(texel); tex_g = __ldg(texel + 1); tex_b = __ldg(texel + 2);
                        // keep the hit for the GI stage
                        gpx = h.pos[0]; gpy = h.pos[1]; gpz = h.pos[2];
                        // sun shadow ray, :139,:152-153
                        ox = h.pos[0] + nx * SCALE * 0.001f; oy = h.pos[1] + ny * SCALE * 0.001f; oz = h.pos[2] + nz * SCALE * 0.001f;
                        tlx = L.light[0] - ox; tly = L.light[1] - oy; tlz = L.light[2] - oz;
                        normalize3(tlx, tly, tlz);
                        dx = tlx; dy = tly; dz = tlz; coef = 0.0f;
                        next = kShadow;
                        break;
                    }
                    case kShadow: {                                            // :155-157
                        if (!r.hit) light = fmaxf(0.0f, dot3(tlx, tly, tlz, nx, ny, nz));
                        if (!L.use_gi) break;
                        // getGlobalIllumination level 0, :169-194 — from the primary hit (gp*, n*)
                        const float c1 = lattice(rnd0.z, -1000.0f, 1000.0f), c2 = lattice(rnd0.w, -1000.0f, 1000.0f);
                        float ax, ay, az;
                        if (nx != 0.0f) { ax = 0.0f; ay = c1; az = c2; }
                        else if (ny != 0.0f) { ax = c1; ay = 0.0f; az = c2; }
                        else if (nz != 0.0f) { ax = c1; ay = c2; az = 0.0f; }
                        else break;                                            // start inside a solid: no estimate
                        ox = gpx + nx * n_norm; oy = gpy + ny * n_norm; oz = gpz + nz * n_norm;     // :174
                        dx = (nx + ax) * n_norm; dy = (ny + ay) * n_norm; dz = (nz + az) * n_norm;  // :192
                        normalize3(dx, dy, dz);
                        dot_gi0 = dot3(dx, dy, dz, nx, ny, nz);               // :193
                        coef = 0.5f;
                        next = kGi0;
                        break;
                    }
                    case kGi0:
                    case kGi1: {                                               // :194-198
                        if (!r.hit) break;
                        if (stage == kGi0) gi0_hit = true; else gi1_hit = true;
                        gnx = h.normal[0]; gny = h.normal[1]; gnz = h.normal[2];
                        gpx = h.pos[0]; gpy = h.pos[1]; gpz = h.pos[2];
                        ox = gpx + gnx * n_norm; oy = gpy + gny * n_norm; oz = gpz + gnz * n_norm;   // :196
                        tlx = L.light[0] - ox; tly = L.light[1] - oy; tlz = L.light[2] - oz;         // :197
                        normalize3(tlx, tly, tlz);
                        dx = tlx; dy = tly; dz = tlz; coef = 0.5f;
                        next = stage + 1;
                        break;
                    }
                    case kGi0Shadow: {                                         // :199-200
                        if (!r.hit) irr0 = fmaxf(0.0f, dot3(gnx, gny, gnz, tlx, tly, tlz));
                        if (L.gi_bounces < 2) break;
                        // second bounce (extension): the same estimator from the GI hit, dimensions 4,5
                        const uint4 rnd1 = philox4x32_10(pixel, sample, 1u, 0u, L.seed_lo, L.seed_hi);
                        const float c1 = lattice(rnd1.x, -1000.0f, 1000.0f), c2 = lattice(rnd1.y, -1000.0f, 1000.0f);
                        float ax, ay, az;
                        if (gnx != 0.0f) { ax = 0.0f; ay = c1; az = c2; }
                        else if (gny != 0.0f) { ax = c1; ay = 0.0f; az = c2; }
                        else if (gnz != 0.0f) { ax = c1; ay = c2; az = 0.0f; }
                        else break;
                        ox = gpx + gnx * n_norm; oy = gpy + gny * n_norm; oz = gpz + gnz * n_norm;
                        dx = (gnx + ax) * n_norm; dy = (gny + ay) * n_norm; dz = (gnz + az) * n_norm;
                        normalize3(dx, dy, dz);
                        dot_gi1 = dot3(dx, dy, dz, gnx, gny, gnz);
                        coef = 0.5f;
                        next = kGi1;
                        break;
                    }
                    case kGi1Shadow: {
                        if (!r.hit) irr1 = fmaxf(0.0f, dot3(gnx, gny, gnz, tlx, tly, tlz));
                        break;
                    }
                    default: break;
                }
                stage = next;
            }

            if (have_hit) {
                float gi = 0.0f;
                if (L.use_gi && gi0_hit) {
                    float irr = irr0;
                    if (L.gi_bounces >= 2) irr = irr + (gi1_hit ? fminf(0.5f, irr1 * dot_gi1) : 0.0f);
                    const float e0 = fminf(0.5f, irr * dot_gi0);              // :201
                    gi = fmaxf(0.0f, 1000000.0f * e0 / 1.0f);                  // :201,:206
                }
                const float f = fminf(1.0f, fmaxf(0.0f, light + gi));         // :163
                sum_r += mul_u8(tex_r, f); sum_g += mul_u8(tex_g, f); sum_b += mul_u8(tex_b, f);
            }
        }
        uint4* a = reinterpret_cast<uint4*>(accum) + pixel;                   // Sample, raycaster.hpp:18-24,87-90
        if (L.spp_chunks == 1) {
            uint4 v = *a;
            v.x += sum_r; v.y += sum_g; v.z += sum_b; v.w += uint32_t(L.spp);
            *a = v;
        } else {
            uint32_t* w = reinterpret_cast<uint32_t*>(a);
            atomicAdd(w, sum_r); atomicAdd(w + 1, sum_g); atomicAdd(w + 2, sum_b); atomicAdd(w + 3, uint32_t(s_end - s_begin));
        }
    }

    // statistics: rays and Σ complexity per ray class (warp reduce, one atomic per warp and class)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        uint32_t a = n_rays[k], b = n_iter[k];
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

// samples_to_image (raycaster.hpp:94-103) or the 0.4/0.6 temporal blend of renderRay (:79-85).
__global__ void resolve_kernel(const uint32_t* __restrict__ accum, uint8_t* __restrict__ rgba, int width, int row_begin,
                               int row_end, int use_samples, int tile_step, int tile_index) {
    const uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;       // index inside the owned rows
    const int x = int(j % uint64_t(width));
    const int r = int(j / uint64_t(width));                                   // owned-row ordinal
    const int y = row_begin + ((r >> 2) * tile_step + tile_index) * 4 + (r & 3);
    if (y >= row_end) return;
    const uint64_t i = uint64_t(y) * width + x;
    const uint4 a = reinterpret_cast<const uint4*>(accum)[i];
    uchar4* out = reinterpret_cast<uchar4*>(rgba) + i;
    if (use_samples) {
        // double division then uint8 conversion == integer division for sums < 2^32
        const uint32_t n = a.w ? a.w : 1u;
        *out = make_uchar4(uint8_t(a.x / n), uint8_t(a.y / n), uint8_t(a.z / n), 255);
    } else {
        const uchar4 old = *out;
        // the accumulator holds exactly one sample in this mode
        const uint8_t nr = mul_u8(uint8_t(a.x), 1.0f - 0.4f), ng = mul_u8(uint8_t(a.y), 1.0f - 0.4f), nb = mul_u8(uint8_t(a.z), 1.0f - 0.4f);
        const int r = min(255, int(mul_u8(old.x, 0.4f)) + int(nr));          // add(sf::Color&, const sf::Color&), utils.cpp:35-40
        const int g = min(255, int(mul_u8(old.y, 0.4f)) + int(ng));
        const int b = min(255, int(mul_u8(old.z, 0.4f)) + int(nb));
        *out = make_uchar4(uint8_t(r), uint8_t(g), uint8_t(b), 255);
    }
}

cudaError_t launch_render_accumulate_ref(const uint2* nodes, const RenderLaunch& L, uint32_t* d_accum,
                                         unsigned long long* d_counters, cudaStream_t stream) {
    const int rows = L.row_end - L.row_begin;
    if (rows <= 0 || L.width <= 0 || L.spp <= 0) return cudaSuccess;
    const int block = 128;
    const int tiles_x = (L.width + 31) / 32, tiles_y = ((rows + 3) / 4 + L.tile_step - 1 - L.tile_index) / L.tile_step;
    if (tiles_y <= 0) return cudaSuccess;
    const size_t smem = size_t(L.depth + 1) * block * 8;
    RefNodes nv{nodes};
    // aim for >= ~96 waves of CTAs (4 CTAs x 148 SMs resident): measured 76.1 ms (4 chunks) vs 77.4 ms (1 chunk) on one GPU,
    // and the last wave stays a small fraction of the launch when a GPU owns 1/8 of the frame
    RenderLaunch Lc = L;
    const long tiles = long(tiles_x) * tiles_y;
    long chunks = (96L * 4 * 148 + tiles - 1) / tiles;
    if (chunks > L.spp) chunks = L.spp;
    if (chunks < 1) chunks = 1;
    if (L.spp_chunks > 0) chunks = L.spp_chunks < L.spp ? L.spp_chunks : L.spp;      // explicit override
    Lc.spp_chunks = int(chunks);
    render_accumulate_kernel<RefNodes><<<unsigned(tiles * chunks), block, smem, stream>>>(nv, Lc, d_accum, d_counters);
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

// K4s — the frame kernel with the per-sample chain state parked in shared memory.
//
// Same pixel → lane mapping, same arithmetic and same results as render_accumulate_kernel
// (render_kernels.cu; reference raycaster.hpp:118-207, camera_controller.hpp:34-49, main.cpp:139-154).
// What changes is where the state lives.  ncu on K4 (profiles/r01_summary.md): 124-128 registers per thread →
// 4 CTAs = 16 warps per SM, 1.6 eligible warps per scheduler, 72 % issue-slot utilisation, while the traversal
// loop itself needs < 50 registers (K1: 48, 32 warps per SM, 81 %).  Everything a sample carries ACROSS a
// traversal — normals, hit points, light vector, texel, partial irradiances, pixel sums — is therefore kept in
// a per-thread column of shared memory (volatile: one LDS/STS per use, never cached in a register) and only
// touched in the short code between rays.  The hot loop then runs with the traversal state alone.
#include "lsvo_step.cuh"
#include "kernels.h"

namespace vrt {

namespace {

enum StageS : int { sPrimary = 0, sShadow = 1, sGi0 = 2, sGi0Shadow = 3, sGi1 = 4, sGi1Shadow = 5, sDone = 6 };

// per-thread shared-memory fields (word index; thread t's field f lives at chain[f * 128 + t])
enum Field : int {
    fNx, fNy, fNz, fLight, fDotGi0, fDotGi1, fIrr0, fIrr1, fGnx, fGny, fGnz, fGpx, fGpy, fGpz, fTlx, fTly, fTlz,
    fTexFlags,              // r | g << 8 | b << 16 | have_hit << 24 | gi0_hit << 25 | gi1_hit << 26
    fRndZ, fRndW,           // Philox words of dimensions 2,3 (GI bounce 1)
    fSumR, fSumG, fSumB, fLensX, fLensY,
    kChainFields
};

struct Chain {
    volatile uint32_t* w;   // already offset by the thread index
    __device__ __forceinline__ float f(int i) const { return __uint_as_float(w[i * 128]); }
    __device__ __forceinline__ uint32_t u(int i) const { return w[i * 128]; }
    __device__ __forceinline__ void set(int i, float v) { w[i * 128] = __float_as_uint(v); }
    __device__ __forceinline__ void setu(int i, uint32_t v) { w[i * 128] = v; }
};

__device__ __forceinline__ uint32_t mul_u8s(uint32_t c, float f) { return uint32_t(uint8_t(fminf(255.0f, float(c) * f))); }   // utils.cpp:43-48

__device__ __forceinline__ void view_to_world_s(const float* m, float vx, float vy, float vz, float& x, float& y, float& z) {
    x = (m[0] * vx + m[1] * vy) + m[2] * vz;                               // v * rot_mat, camera_controller.hpp:51-54
    y = (m[3] * vx + m[4] * vy) + m[5] * vz;
    z = (m[6] * vx + m[7] * vy) + m[8] * vz;
}

}  // namespace

template <typename Nodes, int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) render_smem_kernel(Nodes nodes, RenderLaunch L, uint32_t* __restrict__ accum,
                                                             unsigned long long* __restrict__ counters) {
    extern __shared__ uint2 smem[];
    __shared__ uint32_t s_stats[12];                                        // rays / Σ complexity per class
    Stack64<128> stack{smem + threadIdx.x};
    Chain cs{reinterpret_cast<volatile uint32_t*>(smem + (L.depth + 1) * 128) + threadIdx.x};
    nodes.slots = pin(nodes.slots);
    const int guard = pin(L.guard);
    const int depth_offset = pin(kSvoMaxDepth - L.depth);
    if (threadIdx.x < 12) s_stats[threadIdx.x] = 0u;
    __syncthreads();

    // 8x4 pixel tile per warp, 4 tiles side by side per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_x = (L.width + 31) / 32;
    const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
    const int x = bx * 32 + warp * 8 + (lane & 7);
    const int y = L.row_begin + (by * L.tile_step + L.tile_index) * 4 + (lane >> 3);

    if (x < L.width && y < L.row_end) {
        const float SCALE = 1.0f / float(1 << L.depth);                       // raycaster.hpp:123-124 / main.cpp:82
        const float n_norm = SCALE * 0.0078125f * 2.0f;                       // raycaster.hpp:171-172
        const uint32_t pixel = uint32_t(y) * uint32_t(L.width) + uint32_t(x);
        {
            const float aspect = float(L.width) / float(L.height);            // main.cpp:133
            cs.set(fLensX, float(x) / float(L.height) - aspect * 0.5f);       // main.cpp:145
            cs.set(fLensY, float(y) / float(L.height) - 0.5f);                // main.cpp:146
            cs.setu(fSumR, 0u); cs.setu(fSumG, 0u); cs.setu(fSumB, 0u);
        }

        for (int s = 0; s < L.spp; ++s) {
            float ox, oy, oz, dx, dy, dz, coef;
            {   // Camera::getRay, camera_controller.hpp:34-49
                const uint4 rnd0 = philox4x32_10(pixel, uint32_t(L.sample_offset + s), 0u, 0u, L.seed_lo, L.seed_hi);
                cs.setu(fRndZ, rnd0.z); cs.setu(fRndW, rnd0.w);
                const float u0 = lattice(rnd0.x, -0.5f, 0.5f), u1 = lattice(rnd0.y, -0.5f, 0.5f);
                float fx = cs.f(fLensX), fy = cs.f(fLensY), fz = L.cam.fov;
                normalize3(fx, fy, fz);
                fx *= L.cam.focal_length; fy *= L.cam.focal_length; fz *= L.cam.focal_length;
                const float rx = L.cam.aperture * u0, ry = L.cam.aperture * u1, rz = L.cam.aperture * 0.0f;
                float qx = fx - rx, qy = fy - ry, qz = fz - rz;
                normalize3(qx, qy, qz);
                float wx, wy, wz;
                view_to_world_s(L.cam.rot_mat, qx, qy, qz, dx, dy, dz);
                view_to_world_s(L.cam.rot_mat, rx, ry, rz, wx, wy, wz);
                ox = (L.cam.position[0] + wx) * SCALE + 1.0f;                  // main.cpp:149
                oy = (L.cam.position[1] + wy) * SCALE + 1.0f;
                oz = (L.cam.position[2] + wz) * SCALE + 1.0f;
                coef = 0.0f;
                cs.set(fLight, 0.0f); cs.set(fIrr0, 0.0f); cs.set(fIrr1, 0.0f);
                cs.setu(fTexFlags, 0u);
            }

            int stage = sPrimary;
            while (stage != sDone) {
                LsvoResult r;
                lsvo_cast_ray(nodes, stack, depth_offset, guard, ox, oy, oz, dx, dy, dz, coef, 0.0f, r);
                atomicAdd(&s_stats[stage], 1u);
                if (atomicAdd(&s_stats[6 + stage], r.complexity) >= 0x80000000u) {   // keep the 32-bit counter from wrapping
                    atomicSub(&s_stats[6 + stage], 0x80000000u);
                    atomicAdd(counters + 6 + stage, 0x80000000ull);
                }
                LsvoHit h;
                if (r.hit) lsvo_finish(r, ox, oy, oz, L.depth, h);

                int next = sDone;
                switch (stage) {
                    case sPrimary: {                                           // raycaster.hpp:131-145
                        if (!r.hit) break;
                        const float nx = h.normal[0], ny = h.normal[1], nz = h.normal[2];
                        cs.set(fNx, nx); cs.set(fNy, ny); cs.set(fNz, nz);
                        const uint8_t* tex = (ny != 0.0f) ? L.tex_top : L.tex_side;    // :211-215
                        const float u = fminf(fmaxf(h.uv[0], 0.0f), 1.0f), v = fminf(fmaxf(h.uv[1], 0.0f), 1.0f);   // :237-238
                        const uint32_t tx = uint32_t(16.0f * u), ty = uint32_t(16.0f * v);                          // :239
                        const uint8_t* texel = tex + 3u * (ty * 16u + tx);
                        cs.setu(fTexFlags, uint32_t(__ldg(texel)) | (uint32_t(__ldg(texel + 1)) << 8) | (uint32_t(__ldg(texel + 2)) << 16) | (1u << 24));
                        cs.set(fGpx, h.pos[0]); cs.set(fGpy, h.pos[1]); cs.set(fGpz, h.pos[2]);
                        // sun shadow ray, :139,:152-153
                        ox = h.pos[0] + nx * SCALE * 0.001f; oy = h.pos[1] + ny * SCALE * 0.001f; oz = h.pos[2] + nz * SCALE * 0.001f;
                        float tlx = L.light[0] - ox, tly = L.light[1] - oy, tlz = L.light[2] - oz;
                        normalize3(tlx, tly, tlz);
                        cs.set(fTlx, tlx); cs.set(fTly, tly); cs.set(fTlz, tlz);
                        dx = tlx; dy = tly; dz = tlz; coef = 0.0f;
                        next = sShadow;
                        break;
                    }
                    case sShadow: {                                            // :155-157
                        const float nx = cs.f(fNx), ny = cs.f(fNy), nz = cs.f(fNz);
                        if (!r.hit) cs.set(fLight, fmaxf(0.0f, dot3(cs.f(fTlx), cs.f(fTly), cs.f(fTlz), nx, ny, nz)));
                        if (!L.use_gi) break;
                        // getGlobalIllumination level 0, :169-194 — from the primary hit
                        const float c1 = lattice(cs.u(fRndZ), -1000.0f, 1000.0f), c2 = lattice(cs.u(fRndW), -1000.0f, 1000.0f);
                        float ax, ay, az;
                        if (nx != 0.0f) { ax = 0.0f; ay = c1; az = c2; }
                        else if (ny != 0.0f) { ax = c1; ay = 0.0f; az = c2; }
                        else if (nz != 0.0f) { ax = c1; ay = c2; az = 0.0f; }
                        else break;                                            // start inside a solid: no estimate
                        ox = cs.f(fGpx) + nx * n_norm; oy = cs.f(fGpy) + ny * n_norm; oz = cs.f(fGpz) + nz * n_norm;   // :174
                        dx = (nx + ax) * n_norm; dy = (ny + ay) * n_norm; dz = (nz + az) * n_norm;                      // :192
                        normalize3(dx, dy, dz);
                        cs.set(fDotGi0, dot3(dx, dy, dz, nx, ny, nz));        // :193
                        coef = 0.5f;
                        next = sGi0;
                        break;
                    }
                    case sGi0:
                    case sGi1: {                                               // :194-198
                        if (!r.hit) break;
                        cs.setu(fTexFlags, cs.u(fTexFlags) | (stage == sGi0 ? (1u << 25) : (1u << 26)));
                        const float gnx = h.normal[0], gny = h.normal[1], gnz = h.normal[2];
                        cs.set(fGnx, gnx); cs.set(fGny, gny); cs.set(fGnz, gnz);
                        cs.set(fGpx, h.pos[0]); cs.set(fGpy, h.pos[1]); cs.set(fGpz, h.pos[2]);
                        ox = h.pos[0] + gnx * n_norm; oy = h.pos[1] + gny * n_norm; oz = h.pos[2] + gnz * n_norm;   // :196
                        float tlx = L.light[0] - ox, tly = L.light[1] - oy, tlz = L.light[2] - oz;                  // :197
                        normalize3(tlx, tly, tlz);
                        cs.set(fTlx, tlx); cs.set(fTly, tly); cs.set(fTlz, tlz);
                        dx = tlx; dy = tly; dz = tlz; coef = 0.5f;
                        next = stage + 1;
                        break;
                    }
                    case sGi0Shadow: {                                         // :199-200
                        const float gnx = cs.f(fGnx), gny = cs.f(fGny), gnz = cs.f(fGnz);
                        if (!r.hit) cs.set(fIrr0, fmaxf(0.0f, dot3(gnx, gny, gnz, cs.f(fTlx), cs.f(fTly), cs.f(fTlz))));
                        if (L.gi_bounces < 2) break;
                        // second bounce (extension): the same estimator from the GI hit, dimensions 4,5
                        const uint4 rnd1 = philox4x32_10(pixel, uint32_t(L.sample_offset + s), 1u, 0u, L.seed_lo, L.seed_hi);
                        const float c1 = lattice(rnd1.x, -1000.0f, 1000.0f), c2 = lattice(rnd1.y, -1000.0f, 1000.0f);
                        float ax, ay, az;
                        if (gnx != 0.0f) { ax = 0.0f; ay = c1; az = c2; }
                        else if (gny != 0.0f) { ax = c1; ay = 0.0f; az = c2; }
                        else if (gnz != 0.0f) { ax = c1; ay = c2; az = 0.0f; }
                        else break;
                        ox = cs.f(fGpx) + gnx * n_norm; oy = cs.f(fGpy) + gny * n_norm; oz = cs.f(fGpz) + gnz * n_norm;
                        dx = (gnx + ax) * n_norm; dy = (gny + ay) * n_norm; dz = (gnz + az) * n_norm;
                        normalize3(dx, dy, dz);
                        cs.set(fDotGi1, dot3(dx, dy, dz, gnx, gny, gnz));
                        coef = 0.5f;
                        next = sGi1;
                        break;
                    }
                    case sGi1Shadow: {
                        if (!r.hit) cs.set(fIrr1, fmaxf(0.0f, dot3(cs.f(fGnx), cs.f(fGny), cs.f(fGnz), cs.f(fTlx), cs.f(fTly), cs.f(fTlz))));
                        break;
                    }
                    default: break;
                }
                stage = next;
            }

            const uint32_t tf = cs.u(fTexFlags);
            if (tf & (1u << 24)) {                                            // the primary ray hit
                float gi = 0.0f;
                if (L.use_gi && (tf & (1u << 25))) {
                    float irr = cs.f(fIrr0);
                    if (L.gi_bounces >= 2) irr = irr + ((tf & (1u << 26)) ? fminf(0.5f, cs.f(fIrr1) * cs.f(fDotGi1)) : 0.0f);
                    gi = fmaxf(0.0f, 1000000.0f * fminf(0.5f, irr * cs.f(fDotGi0)) / 1.0f);   // :201,:206
                }
                const float f = fminf(1.0f, fmaxf(0.0f, cs.f(fLight) + gi));   // :163
                cs.setu(fSumR, cs.u(fSumR) + mul_u8s(tf & 0xffu, f));
                cs.setu(fSumG, cs.u(fSumG) + mul_u8s((tf >> 8) & 0xffu, f));
                cs.setu(fSumB, cs.u(fSumB) + mul_u8s((tf >> 16) & 0xffu, f));
            }
        }
        uint4* a = reinterpret_cast<uint4*>(accum) + pixel;                   // Sample, raycaster.hpp:18-24,87-90
        uint4 v = *a;
        v.x += cs.u(fSumR); v.y += cs.u(fSumG); v.z += cs.u(fSumB); v.w += uint32_t(L.spp);
        *a = v;
    }
    __syncthreads();
    if (threadIdx.x < 12 && s_stats[threadIdx.x]) atomicAdd(counters + threadIdx.x, (unsigned long long)s_stats[threadIdx.x]);
}

template <int kMinBlocks>
static cudaError_t launch_smem_variant(const RefNodes& nv, const RenderLaunch& L, uint32_t* d_accum, unsigned long long* d_counters,
                                       unsigned grid, size_t smem, cudaStream_t stream) {
    auto kernel = render_smem_kernel<RefNodes, kMinBlocks>;
    // kMinBlocks CTAs of (stack + chain) shared memory must fit next to L1: ask for the carve-out that allows it
    int carve = int((smem * kMinBlocks * 100 + 228 * 1024 - 1) / (228 * 1024)) + 8;
    if (carve > 100) carve = 100;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    kernel<<<grid, 128, smem, stream>>>(nv, L, d_accum, d_counters);
    return cudaGetLastError();
}

cudaError_t launch_render_smem(const uint2* nodes, const RenderLaunch& L, uint32_t* d_accum, unsigned long long* d_counters,
                               int min_blocks, cudaStream_t stream) {
    const int rows = L.row_end - L.row_begin;
    if (rows <= 0 || L.width <= 0 || L.spp <= 0) return cudaSuccess;
    const int block = 128;
    const int tiles_x = (L.width + 31) / 32, tiles_y = ((rows + 3) / 4 + L.tile_step - 1 - L.tile_index) / L.tile_step;
    if (tiles_y <= 0) return cudaSuccess;
    const size_t smem = size_t(L.depth + 1) * block * 8 + size_t(kChainFields) * block * 4;
    RefNodes nv{nodes};
    const unsigned grid = unsigned(tiles_x) * unsigned(tiles_y);
    switch (min_blocks) {
        case 8: return launch_smem_variant<8>(nv, L, d_accum, d_counters, grid, smem, stream);
        case 7: return launch_smem_variant<7>(nv, L, d_accum, d_counters, grid, smem, stream);
        case 5: return launch_smem_variant<5>(nv, L, d_accum, d_counters, grid, smem, stream);
        default: return launch_smem_variant<6>(nv, L, d_accum, d_counters, grid, smem, stream);
    }
}

}  // namespace vrt

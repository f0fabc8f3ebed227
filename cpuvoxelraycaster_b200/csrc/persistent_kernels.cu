// Persistent-thread kernels with per-lane ray regeneration (kernels K1p and K4p).
//
// One CTA per SM slot, resident for the whole launch.  A warp fetches work in chunks with one atomic
// (warp work-fetching), hands individual rays / pixels to lanes by ballot rank, and runs the LSVO
// traversal loop one trip at a time (lsvo_step.cuh).  A lane whose ray has terminated parks; as soon as
// `refill` lanes are parked the warp leaves the traversal loop, retires their results, regenerates
// their next rays (shadow / GI / next sample / next pixel, or the next ray of the buffer) and resumes.
// Every lane's arithmetic is exactly that of LSVO<D>::castRay / RayCaster::castRay in the reference
// (see lsvo_step.cuh, render_kernels.cu), so results do not depend on the scheduling.
#include "lsvo_step.cuh"
#include "kernels.h"

namespace vrt {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ void store_hit_record(vrt_hit* out, const LsvoResult& r, const LsvoHit& h, int depth) {
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

// ---- K1p: LSVO<D>::castRay over a ray buffer ---------------------------------------------------------------
// Work is fetched 64 rays at a time (one atomic per fetch) so that the tail stays balanced.
// refill > 0: fixed regeneration threshold.  refill == 0 (default): warp-adaptive — a warp starts in
// "coherent" mode (threshold 32: 32 rays walk in lock-step like the reference's pixel neighbours, which keeps
// the push/advance/pop branches uniform); after every synchronous batch it measures the batch's SIMT
// efficiency sum(iters) / (32 * max(iters)) and drops to threshold 8 when rays turn out incoherent
// (ncu: 4.7 of 32 lanes active on random rays without regeneration), probing coherent mode again now and then.
constexpr int kRayChunk = 64;

template <typename Nodes>
__global__ void __launch_bounds__(128, 4) lsvo_cast_persistent_kernel(Nodes nodes, int depth, int guard,
                                                                      const float* __restrict__ origin,
                                                                      const float* __restrict__ dir, float coef, float bias,
                                                                      uint64_t n, vrt_hit* __restrict__ out,
                                                                      unsigned long long* __restrict__ counters, int refill) {
    extern __shared__ uint2 smem[];
    Stack64<128> stack{smem + threadIdx.x};
    nodes.slots = pin(nodes.slots);
    guard = pin(guard);
    const int depth_offset = pin(kSvoMaxDepth - depth);
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const bool adaptive = refill <= 0;
    int threshold = adaptive ? 32 : refill;                // warp-uniform
    int refills_since_probe = 0;
    bool sync_batch = false;                               // all 32 lanes were started in the same refill phase

    Trav t;
    bool alive = false, has_result = false, exhausted = false;
    uint64_t ray = 0, chunk_next = 0, chunk_end = 0;       // chunk_* are warp-uniform
    unsigned long long iter_sum = 0;

    for (;;) {
        // ---- refill phase: retire parked lanes, hand out new rays ----
        const bool retiring = !alive && has_result;
        if (adaptive) {
            const unsigned rmask = __ballot_sync(kFull, retiring);
            if (threshold == 32 && sync_batch && rmask == kFull) {
                uint32_t sum = t.iters, mx = t.iters;
                for (int o = 16; o > 0; o >>= 1) {
                    sum += __shfl_xor_sync(kFull, sum, o);
                    mx = max(mx, __shfl_xor_sync(kFull, mx, o));
                }
                if (sum * 100u < 55u * 32u * mx) threshold = 8;      // < 55 % of the lanes did useful work
            } else if (threshold != 32 && ++refills_since_probe >= 256) {
                threshold = 32;                                      // drain, then measure one synchronous batch
                refills_since_probe = 0;
            }
        }
        if (retiring) {
            LsvoResult r;
            t.result(r);
            LsvoHit h;
            if (r.hit) lsvo_finish(r, t.ox, t.oy, t.oz, depth, h);
            store_hit_record(out + ray, r, h, depth);
            iter_sum += t.iters;
            has_result = false;
        }
        unsigned want = __ballot_sync(kFull, !alive);
        sync_batch = want == kFull;
        while (want && !exhausted) {
            if (chunk_next == chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counters + 1, (unsigned long long)kRayChunk);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = (base + kRayChunk < n) ? base + kRayChunk : n;
            }
            const unsigned rank = __popc(want & lt);
            const bool take = ((want >> lane) & 1u) && (uint64_t(rank) < chunk_end - chunk_next);
            if (take) {
                ray = chunk_next + rank;
                t.init(origin[3 * ray], origin[3 * ray + 1], origin[3 * ray + 2], dir[3 * ray], dir[3 * ray + 1], dir[3 * ray + 2],
                       coef, bias);
                alive = true;
                has_result = true;
            }
            const unsigned took = __ballot_sync(kFull, take);
            chunk_next += __popc(took);
            want &= ~took;
        }
        if (want) sync_batch = false;                      // the buffer ran dry: not a full batch
        if (!__ballot_sync(kFull, alive)) break;
        // ---- traversal phase: step until `threshold` lanes are parked (or, at the tail, until all are) ----
        if (exhausted || threshold == 32) {
            // lock-step batch (or tail): every lane runs its ray to the end, no per-trip vote needed
            while (alive) alive = t.step(nodes, stack, depth_offset, guard);
            __syncwarp();
        } else {
            const int min_alive = 32 - threshold + 1;
            do {
                if (alive) alive = t.step(nodes, stack, depth_offset, guard);
            } while (__popc(__ballot_sync(kFull, alive)) >= min_alive);
        }
    }
    for (int o = 16; o > 0; o >>= 1) iter_sum += __shfl_xor_sync(kFull, iter_sum, o);
    if (lane == 0 && iter_sum) atomicAdd(counters, iter_sum);
}

// ---- K4p: frame rendering ------------------------------------------------------------------------------------
enum Stage : int { kPrimary = 0, kShadow = 1, kGi0 = 2, kGi0Shadow = 3, kGi1 = 4, kGi1Shadow = 5, kDone = 6 };

__device__ __forceinline__ uint8_t mul_u8p(uint8_t c, float f) { return uint8_t(fminf(255.0f, float(c) * f)); }   // utils.cpp:43-48

__device__ __forceinline__ void view_to_world_p(const float* m, float vx, float vy, float vz, float& x, float& y, float& z) {
    x = (m[0] * vx + m[1] * vy) + m[2] * vz;                               // v * rot_mat, camera_controller.hpp:51-54
    y = (m[3] * vx + m[4] * vy) + m[5] * vz;
    z = (m[6] * vx + m[7] * vy) + m[8] * vz;
}

template <typename Nodes>
__global__ void __launch_bounds__(128, 4) render_persistent_kernel(Nodes nodes, RenderLaunch L, uint32_t* __restrict__ accum,
                                                                   unsigned long long* __restrict__ counters, int refill) {
    extern __shared__ uint2 smem[];
    __shared__ uint32_t s_stats[12];                                        // rays / complexity per class, spilled to global at 2^31
    Stack64<128> stack{smem + threadIdx.x};
    nodes.slots = pin(nodes.slots);
    const int guard = pin(L.guard);
    if (threadIdx.x < 12) s_stats[threadIdx.x] = 0u;
    __syncthreads();
    const int depth_offset = pin(kSvoMaxDepth - L.depth);
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();

    // work space: owned 4-row tiles, cut into 8x4 pixel tiles of 32 ordinals each (coherent at warp start)
    const int rows = L.row_end - L.row_begin;
    const int tile_rows = ((rows + 3) / 4 + L.tile_step - 1 - L.tile_index) / L.tile_step;
    const int tiles_x = (L.width + 7) / 8;
    const unsigned long long n_ord = (unsigned long long)(tile_rows > 0 ? tile_rows : 0) * tiles_x * 32ull;

    const float SCALE = 1.0f / float(1 << L.depth);                         // raycaster.hpp:123-124 / main.cpp:82
    const float n_norm = SCALE * 0.0078125f * 2.0f;                         // raycaster.hpp:171-172
    const float aspect = float(L.width) / float(L.height);                  // main.cpp:133

    Trav t;
    bool alive = false, has_ray = false, have_pixel = false, exhausted = false;
    unsigned long long chunk_next = 0, chunk_end = 0;                       // warp-uniform
    // pixel state
    uint32_t pixel = 0, sum_r = 0, sum_g = 0, sum_b = 0;
    float lens_x = 0.f, lens_y = 0.f;
    int s = 0;
    // chain state (one sample)
    int stage = kPrimary;
    uint4 rnd0 = make_uint4(0, 0, 0, 0);
    float nx = 0.f, ny = 0.f, nz = 0.f, light = 0.f;
    float dot_gi0 = 0.f, dot_gi1 = 0.f, irr0 = 0.f, irr1 = 0.f;
    float gnx = 0.f, gny = 0.f, gnz = 0.f, gpx = 0.f, gpy = 0.f, gpz = 0.f;
    float tlx = 0.f, tly = 0.f, tlz = 0.f;
    uint8_t tex_r = 0, tex_g = 0, tex_b = 0;
    bool have_hit = false, gi0_hit = false, gi1_hit = false;

    for (;;) {
        // ================= refill phase =================
        float rox = 0.f, roy = 0.f, roz = 0.f, rdx = 0.f, rdy = 0.f, rdz = 0.f, rcoef = 0.f;
        bool new_ray = false, want_primary = false;
        if (!alive && has_ray) {
            // ---- A: the lane's ray has terminated: advance the sample's chain (raycaster.hpp:118-207) ----
            has_ray = false;
            LsvoResult r;
            t.result(r);
            LsvoHit h;
            if (r.hit) lsvo_finish(r, t.ox, t.oy, t.oz, L.depth, h);
            atomicAdd(&s_stats[stage], 1u);
            if (atomicAdd(&s_stats[6 + stage], r.complexity) >= 0x80000000u) {
                atomicSub(&s_stats[6 + stage], 0x80000000u);
                atomicAdd(counters + 6 + stage, 0x80000000ull);
            }
            int next = kDone;
            switch (stage) {
                case kPrimary: {                                             // :131-145
                    if (!r.hit) break;
                    have_hit = true;
                    nx = h.normal[0]; ny = h.normal[1]; nz = h.normal[2];
                    const uint8_t* tex = (ny != 0.0f) ? L.tex_top : L.tex_side;            // :211-215
                    const float u = fminf(fmaxf(h.uv[0], 0.0f), 1.0f), v = fminf(fmaxf(h.uv[1], 0.0f), 1.0f);   // :237-238
                    const uint32_t tx = uint32_t(16.0f * u), ty = uint32_t(16.0f * v);      // :239
                    const uint8_t* texel = tex + 3u * (ty * 16u + tx);
                    tex_r = __ldg(texel); tex_g = __ldg(texel + 1); tex_b = __ldg(texel + 2);
                    gpx = h.pos[0]; gpy = h.pos[1]; gpz = h.pos[2];
                    rox = h.pos[0] + nx * SCALE * 0.001f; roy = h.pos[1] + ny * SCALE * 0.001f; roz = h.pos[2] + nz * SCALE * 0.001f;   // :139
                    tlx = L.light[0] - rox; tly = L.light[1] - roy; tlz = L.light[2] - roz;   // :152
                    normalize3(tlx, tly, tlz);
                    rdx = tlx; rdy = tly; rdz = tlz; rcoef = 0.0f;
                    next = kShadow;
                    break;
                }
                case kShadow: {                                              // :155-157
                    if (!r.hit) light = fmaxf(0.0f, dot3(tlx, tly, tlz, nx, ny, nz));
                    if (!L.use_gi) break;
                    const float c1 = lattice(rnd0.z, -1000.0f, 1000.0f), c2 = lattice(rnd0.w, -1000.0f, 1000.0f);   // :180-181
                    float ax, ay, az;
                    if (nx != 0.0f) { ax = 0.0f; ay = c1; az = c2; }         // :182-190
                    else if (ny != 0.0f) { ax = c1; ay = 0.0f; az = c2; }
                    else if (nz != 0.0f) { ax = c1; ay = c2; az = 0.0f; }
                    else break;
                    rox = gpx + nx * n_norm; roy = gpy + ny * n_norm; roz = gpz + nz * n_norm;      // :174
                    rdx = (nx + ax) * n_norm; rdy = (ny + ay) * n_norm; rdz = (nz + az) * n_norm;   // :192
                    normalize3(rdx, rdy, rdz);
                    dot_gi0 = dot3(rdx, rdy, rdz, nx, ny, nz);               // :193
                    rcoef = 0.5f;
                    next = kGi0;
                    break;
                }
                case kGi0:
                case kGi1: {                                                 // :194-198
                    if (!r.hit) break;
                    if (stage == kGi0) gi0_hit = true; else gi1_hit = true;
                    gnx = h.normal[0]; gny = h.normal[1]; gnz = h.normal[2];
                    gpx = h.pos[0]; gpy = h.pos[1]; gpz = h.pos[2];
                    rox = gpx + gnx * n_norm; roy = gpy + gny * n_norm; roz = gpz + gnz * n_norm;   // :196
                    tlx = L.light[0] - rox; tly = L.light[1] - roy; tlz = L.light[2] - roz;         // :197
                    normalize3(tlx, tly, tlz);
                    rdx = tlx; rdy = tly; rdz = tlz; rcoef = 0.5f;
                    next = stage + 1;
                    break;
                }
                case kGi0Shadow: {                                           // :199-200
                    if (!r.hit) irr0 = fmaxf(0.0f, dot3(gnx, gny, gnz, tlx, tly, tlz));
                    if (L.gi_bounces < 2) break;
                    const uint4 rnd1 = philox4x32_10(pixel, uint32_t(L.sample_offset + s), 1u, 0u, L.seed_lo, L.seed_hi);
                    const float c1 = lattice(rnd1.x, -1000.0f, 1000.0f), c2 = lattice(rnd1.y, -1000.0f, 1000.0f);
                    float ax, ay, az;
                    if (gnx != 0.0f) { ax = 0.0f; ay = c1; az = c2; }
                    else if (gny != 0.0f) { ax = c1; ay = 0.0f; az = c2; }
                    else if (gnz != 0.0f) { ax = c1; ay = c2; az = 0.0f; }
                    else break;
                    rox = gpx + gnx * n_norm; roy = gpy + gny * n_norm; roz = gpz + gnz * n_norm;
                    rdx = (gnx + ax) * n_norm; rdy = (gny + ay) * n_norm; rdz = (gnz + az) * n_norm;
                    normalize3(rdx, rdy, rdz);
                    dot_gi1 = dot3(rdx, rdy, rdz, gnx, gny, gnz);
                    rcoef = 0.5f;
                    next = kGi1;
                    break;
                }
                case kGi1Shadow: {
                    if (!r.hit) irr1 = fmaxf(0.0f, dot3(gnx, gny, gnz, tlx, tly, tlz));
                    break;
                }
                default: break;
            }
            if (next != kDone) {
                stage = next;
                new_ray = true;
            } else {
                // the sample is complete: colour (raycaster.hpp:161-163) and accumulation (:87-90)
                if (have_hit) {
                    float gi = 0.0f;
                    if (L.use_gi && gi0_hit) {
                        float irr = irr0;
                        if (L.gi_bounces >= 2) irr = irr + (gi1_hit ? fminf(0.5f, irr1 * dot_gi1) : 0.0f);
                        gi = fmaxf(0.0f, 1000000.0f * fminf(0.5f, irr * dot_gi0) / 1.0f);           // :201,:206
                    }
                    const float f = fminf(1.0f, fmaxf(0.0f, light + gi));                           // :163
                    sum_r += mul_u8p(tex_r, f); sum_g += mul_u8p(tex_g, f); sum_b += mul_u8p(tex_b, f);
                }
                ++s;
                if (s < L.spp) {
                    want_primary = true;
                } else {
                    uint4* a = reinterpret_cast<uint4*>(accum) + pixel;       // Sample, raycaster.hpp:18-24
                    uint4 v = *a;
                    v.x += sum_r; v.y += sum_g; v.z += sum_b; v.w += uint32_t(L.spp);
                    *a = v;
                    have_pixel = false;
                }
            }
        }
        // ---- B: lanes without a pixel take the next ordinals of the warp's chunk ----
        unsigned want = __ballot_sync(kFull, !alive && !new_ray && !want_primary && !have_pixel);
        while (want && !exhausted) {
            if (chunk_next == chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counters + 12, 32ull);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n_ord) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = base + 32ull;
            }
            const unsigned rank = __popc(want & lt);
            const bool take = ((want >> lane) & 1u) && ((unsigned long long)rank < chunk_end - chunk_next);
            bool valid = false;
            if (take) {
                const unsigned long long q = chunk_next + rank;
                const int tile = int(q >> 5), k = int(q & 31ull);
                const int ty = tile / tiles_x, txi = tile - ty * tiles_x;
                const int x = txi * 8 + (k & 7);
                const int y = L.row_begin + (ty * L.tile_step + L.tile_index) * 4 + (k >> 3);
                if (x < L.width && y < L.row_end) {
                    valid = true;
                    pixel = uint32_t(y) * uint32_t(L.width) + uint32_t(x);
                    lens_x = float(x) / float(L.height) - aspect * 0.5f;     // main.cpp:145
                    lens_y = float(y) / float(L.height) - 0.5f;              // main.cpp:146
                    sum_r = sum_g = sum_b = 0u;
                    s = 0;
                    have_pixel = true;
                    want_primary = true;
                }
            }
            const unsigned took = __ballot_sync(kFull, take);
            chunk_next += __popc(took);
            want &= ~__ballot_sync(kFull, take && valid);                    // off-frame ordinals ask again
        }
        // ---- C: Camera::getRay (camera_controller.hpp:34-49) for lanes starting a sample ----
        if (want_primary) {
            rnd0 = philox4x32_10(pixel, uint32_t(L.sample_offset + s), 0u, 0u, L.seed_lo, L.seed_hi);
            const float u0 = lattice(rnd0.x, -0.5f, 0.5f), u1 = lattice(rnd0.y, -0.5f, 0.5f);
            float fx = lens_x, fy = lens_y, fz = L.cam.fov;
            normalize3(fx, fy, fz);
            fx *= L.cam.focal_length; fy *= L.cam.focal_length; fz *= L.cam.focal_length;
            const float rx = L.cam.aperture * u0, ry = L.cam.aperture * u1, rz = L.cam.aperture * 0.0f;
            float qx = fx - rx, qy = fy - ry, qz = fz - rz;
            normalize3(qx, qy, qz);
            float wx, wy, wz;
            view_to_world_p(L.cam.rot_mat, qx, qy, qz, rdx, rdy, rdz);
            view_to_world_p(L.cam.rot_mat, rx, ry, rz, wx, wy, wz);
            rox = (L.cam.position[0] + wx) * SCALE + 1.0f;                   // main.cpp:149
            roy = (L.cam.position[1] + wy) * SCALE + 1.0f;
            roz = (L.cam.position[2] + wz) * SCALE + 1.0f;
            rcoef = 0.0f;
            stage = kPrimary;
            light = 0.f; irr0 = 0.f; irr1 = 0.f;
            have_hit = false; gi0_hit = false; gi1_hit = false;
            new_ray = true;
        }
        // ---- D: prologue of castRay for every regenerated ray ----
        if (new_ray) {
            t.init(rox, roy, roz, rdx, rdy, rdz, rcoef, 0.0f);
            alive = true;
            has_ray = true;
        }
        if (!__ballot_sync(kFull, alive)) break;
        // ================= traversal phase =================
        const int min_alive = exhausted ? 1 : 32 - refill + 1;
        do {
            if (alive) alive = t.step(nodes, stack, depth_offset, guard);
        } while (__popc(__ballot_sync(kFull, alive)) >= min_alive);
    }
    __syncthreads();
    if (threadIdx.x < 12 && s_stats[threadIdx.x]) atomicAdd(counters + threadIdx.x, (unsigned long long)s_stats[threadIdx.x]);
}

// ---- launchers -----------------------------------------------------------------------------------------------
template <typename K>
static int resident_blocks(K kernel, int block, size_t smem) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * sms;            // a whole number of CTAs per SM: 148 x occupancy on B200
}

cudaError_t launch_lsvo_cast_persistent(const uint2* nodes, int depth, int guard, const float* d_origin, const float* d_dir,
                                        float coef, float bias, uint64_t n, vrt_hit* d_out, unsigned long long* d_counters,
                                        int refill, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int block = 128;
    const size_t smem = size_t(depth + 1) * block * 8;
    auto kernel = lsvo_cast_persistent_kernel<RefNodes>;
    uint64_t grid = uint64_t(resident_blocks(kernel, block, smem));
    const uint64_t need = (n + block - 1) / block;
    if (need < grid) grid = need;
    RefNodes nv{nodes};
    kernel<<<unsigned(grid), block, smem, stream>>>(nv, depth, guard, d_origin, d_dir, coef, bias, n, d_out, d_counters, refill);
    return cudaGetLastError();
}

cudaError_t launch_render_persistent(const uint2* nodes, const RenderLaunch& L, uint32_t* d_accum, unsigned long long* d_counters,
                                     int refill, cudaStream_t stream) {
    const int rows = L.row_end - L.row_begin;
    if (rows <= 0 || L.width <= 0 || L.spp <= 0) return cudaSuccess;
    const int tile_rows = ((rows + 3) / 4 + L.tile_step - 1 - L.tile_index) / L.tile_step;
    if (tile_rows <= 0) return cudaSuccess;
    const int block = 128;
    const size_t smem = size_t(L.depth + 1) * block * 8;
    auto kernel = render_persistent_kernel<RefNodes>;
    uint64_t grid = uint64_t(resident_blocks(kernel, block, smem));
    const uint64_t need = (uint64_t(tile_rows) * ((L.width + 7) / 8) * 32 + block - 1) / block;
    if (need < grid) grid = need;
    RefNodes nv{nodes};
    kernel<<<unsigned(grid), block, smem, stream>>>(nv, L, d_accum, d_counters, refill);
    return cudaGetLastError();
}

}  // namespace vrt

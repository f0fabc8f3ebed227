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
#include "render_chain.cuh"

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

#ifndef VRT_K1P_MIN_CTAS
#define VRT_K1P_MIN_CTAS 8
#endif
template <typename Nodes, bool kCone, bool kGuard>
__global__ void __launch_bounds__(128, VRT_K1P_MIN_CTAS) lsvo_cast_persistent_kernel(Nodes nodes, int depth, int guard,
                                                                      const float* __restrict__ origin,
                                                                      const float* __restrict__ dir, float coef, float bias,
                                                                      uint64_t n, vrt_hit* __restrict__ out,
                                                                      unsigned long long* __restrict__ counters, int refill,
                                                                      const unsigned long long* __restrict__ gate, unsigned long long want) {
    extern __shared__ uint2 smem[];
    if (gate && *gate != want) return;                     // the automatic K1b / K1p choice (lsvo_kernels.cu, classify_rays_kernel)
    nodes.slots = pin(nodes.slots);
    const float guard_sf = keep_in_register(guard_scale_f(guard), smem + threadIdx.x);
    guard = keep_in_register(guard, smem + threadIdx.x);
    Stack64s<128> stack = Stack64s<128>::make(smem + threadIdx.x, kSvoMaxDepth - depth);
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const bool adaptive = refill <= 0;
    int threshold = adaptive ? 32 : refill;                // warp-uniform
    int refills_since_probe = 0;
    bool sync_batch = false;                               // all 32 lanes were started in the same refill phase

    Trav2<kCone, false, kGuard> t;
    bool alive = false, has_result = false, exhausted = false;
    uint64_t ray = 0, chunk_next = 0, chunk_end = 0;       // chunk_* are warp-uniform
    unsigned long long iter_sum = 0;

    for (;;) {
        // ---- refill phase: retire parked lanes, hand out new rays ----
        const bool retiring = !alive && has_result;
        if (adaptive) {
            const unsigned rmask = __ballot_sync(kFull, retiring);
            if (threshold == 32 && sync_batch && rmask == kFull) {
                uint32_t sum = uint32_t(t.iters_f), mx = sum;
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
            t.result(r, guard_sf);
            LsvoHit h;
            if (r.hit) lsvo_finish(r, origin[3 * ray], origin[3 * ray + 1], origin[3 * ray + 2], depth, h);   // re-read, not carried in registers
            store_hit_record(out + ray, r, h, depth);
            iter_sum += r.complexity;
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
                t.prime(nodes);
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
            while (alive) alive = t.step(nodes, stack, guard, guard_sf);
            __syncwarp();
        } else {
            const int min_alive = 32 - threshold + 1;
            do {
                if (alive) alive = t.step(nodes, stack, guard, guard_sf);
            } while (__popc(__ballot_sync(kFull, alive)) >= min_alive);
        }
    }
    for (int o = 16; o > 0; o >>= 1) iter_sum += __shfl_xor_sync(kFull, iter_sum, o);
    if (lane == 0 && iter_sum) atomicAdd(counters, iter_sum);
}

// ---- K4p: frame rendering ------------------------------------------------------------------------------------
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
    uint32_t pixel = 0, sum_r = 0, sum_g = 0, sum_b = 0;                    // pixel state
    float lens_x = 0.f, lens_y = 0.f;
    int s = 0;
    int stage = kPrimary;                                                   // chain state (one sample)
    ChainState c;

    for (;;) {
        // ================= refill phase =================
        NextRay nr;
        bool new_ray = false, want_primary = false;
        if (!alive && has_ray) {
            // ---- A: the lane's ray has terminated: advance the sample's chain (render_chain.cuh) ----
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
            const int next = chain_advance(L, c, stage, r, h, pixel, uint32_t(L.sample_offset + s), SCALE, n_norm, nr);
            if (next != kDone) {
                stage = next;
                new_ray = true;
            } else {
                chain_colour(L, c, sum_r, sum_g, sum_b);                     // the sample is complete
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
        // ---- C: Camera::getRay for lanes starting a sample ----
        if (want_primary) {
            chain_begin(L, c, pixel, uint32_t(L.sample_offset + s), lens_x, lens_y, SCALE, L.focal ? __ldg(L.focal) : L.cam.focal_length, nr);
            stage = kPrimary;
            new_ray = true;
        }
        // ---- D: prologue of castRay for every regenerated ray ----
        if (new_ray) {
            t.init(nr.ox, nr.oy, nr.oz, nr.dx, nr.dy, nr.dz, nr.coef, 0.0f);
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

template <typename Nodes>
static cudaError_t cast_persistent(Nodes nv, int depth, int guard, const float* d_origin, const float* d_dir, float coef, float bias,
                                   uint64_t n, vrt_hit* d_out, unsigned long long* d_counters, int refill, cudaStream_t stream,
                                   const unsigned long long* d_gate, unsigned long long want) {
    const int block = 128;
    const size_t smem = size_t(depth + 1) * block * 8;
    auto launch = [&](auto kernel) {
        uint64_t grid = uint64_t(resident_blocks(kernel, block, smem));
        const uint64_t need = (n + block - 1) / block;
        if (need < grid) grid = need;
        kernel<<<unsigned(grid), block, smem, stream>>>(nv, depth, guard, d_origin, d_dir, coef, bias, n, d_out, d_counters, refill, d_gate, want);
    };
    // the cone test is compiled out when it cannot fire (coef = bias = 0: lsvo_step.cuh, Trav2)
    // ... and the loop guard when it cannot bind (guard below the voxels' own scale)
    const bool cone = !(coef == 0.0f && bias == 0.0f), guarded = guard >= kSvoMaxDepth - depth;
    if (cone) { if (guarded) launch(lsvo_cast_persistent_kernel<Nodes, true, true>); else launch(lsvo_cast_persistent_kernel<Nodes, true, false>); }
    else { if (guarded) launch(lsvo_cast_persistent_kernel<Nodes, false, true>); else launch(lsvo_cast_persistent_kernel<Nodes, false, false>); }
    return cudaGetLastError();
}

cudaError_t launch_lsvo_cast_persistent(const uint2* nodes, bool compact, int depth, int guard, const float* d_origin, const float* d_dir,
                                        float coef, float bias, uint64_t n, vrt_hit* d_out, unsigned long long* d_counters,
                                        int refill, cudaStream_t stream, const unsigned long long* d_gate, unsigned long long want) {
    if (n == 0) return cudaSuccess;
    return compact ? cast_persistent(CompactNodes{nodes}, depth, guard, d_origin, d_dir, coef, bias, n, d_out, d_counters, refill, stream, d_gate, want)
                   : cast_persistent(RefNodes{nodes}, depth, guard, d_origin, d_dir, coef, bias, n, d_out, d_counters, refill, stream, d_gate, want);
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

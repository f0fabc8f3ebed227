// Scene construction on the GPU: the demo terrain T(D) straight into the reference's LNode layout, in HBM.
//
// Replaces, with byte-identical output (tests compare against the host builder and the reference's compileSVO):
//   src/main.cpp:61-76                   FastNoise SimplexFractal heights + the SVO::setCell fill
//   include/svo.hpp:72-114               SVO::setCell / rec_setCell (80-byte pointer nodes — never built)
//   include/lsvo_utils.hpp:45-55,
//   src/lsvo_utils.cpp:4-49              compileSVO / compileSVO_rec: DFS pre-order, 8 slots per non-empty node,
//                                        children visited x-outer / y / z-inner, slot = z*4 + y*2 + x
// The reference needs 10 s at 1024^3 and cannot build 4096^3 at all; the host builder (scene_host.cpp) takes
// 0.7 s / 2.9 s at 2048^3 / 4096^3 plus the PCIe upload of 1.35 / 5.4 GB.  Here nothing crosses PCIe.
//
// DFS numbering without a DFS.  Interior nodes of the terrain at level l (cubes of edge 2^l) are, per (x,z)
// column of that level, the contiguous run yi = bottom>>l .. top_l(x,z)>>l.  With
//     size(N)      = 8 + sum of size(C) over N's interior children          (slots emitted in N's subtree)
//     child_pos(N) = index of N's 8-slot child block
// the recursive pre-order of compileSVO_rec is:  child_pos(root) = 1;  for N's children in visit order,
//     self(C) = child_pos(N) + slot(C);  child_pos(C) = child_pos(N) + 8 + sum of size() of earlier interior siblings.
// One bottom-up sweep gives size(), one top-down sweep gives self()/child_pos(), one sweep writes the slots.
#include <vector>

#include "host_util.h"
#include "kernels.h"
#include "vrt_device.cuh"

namespace vrt {

namespace {

struct Level {
    int n;                 // columns per side at this level: S >> l
    int ybase;             // bottom >> l
    const int32_t* top;    // [n*n] highest solid y of the column square
    const uint32_t* off;   // [n*n] exclusive prefix sum of node counts
    uint32_t* size;        // per node
    uint32_t* cpos;
    uint32_t* self;
};

__constant__ uint8_t c_perm[512];
__constant__ uint8_t c_perm12[512];

__device__ __forceinline__ int ffloor_dev(float f) { return f >= 0 ? int(f) : int(f) - 1; }

__device__ __forceinline__ float simplex_corner(uint8_t offset, int ix, int iy, float fx, float fy) {
    float t = 0.5f - fx * fx - fy * fy;
    if (t < 0) return 0.0f;
    const float gx[12] = {1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
    const float gy[12] = {1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
    const uint8_t g = c_perm12[(ix & 0xff) + c_perm[(iy & 0xff) + offset]];
    t *= t;
    return t * t * (fx * gx[g] + fy * gy[g]);
}

// FastNoise 0.4.1 SingleSimplex 2-D, lib/fastnoise/FastNoise.cpp:1275-1333
__device__ __forceinline__ float simplex_octave(uint8_t offset, float x, float y) {
    const float sqrt3 = 1.7320508075688772935274463415059f;
    const float skew = 0.5f * (sqrt3 - 1.0f), unskew = (3.0f - sqrt3) / 6.0f;
    float t = (x + y) * skew;
    const int i = ffloor_dev(x + t), j = ffloor_dev(y + t);
    t = float(i + j) * unskew;
    const float x0 = x - (float(i) - t), y0 = y - (float(j) - t);
    const int i1 = x0 > y0 ? 1 : 0, j1 = 1 - i1;
    const float x1 = x0 - float(i1) + unskew, y1 = y0 - float(j1) + unskew;
    const float x2 = x0 - 1 + 2 * unskew, y2 = y0 - 1 + 2 * unskew;
    const float n0 = simplex_corner(offset, i, j, x0, y0);
    const float n1 = simplex_corner(offset, i + i1, j + j1, x1, y1);
    const float n2 = simplex_corner(offset, i + 1, j + 1, x2, y2);
    return 70 * (n0 + n1 + n2);
}

// main.cpp:68 height = int32(64 * GetNoise(.75x, .75z) + 32) and main.cpp:71-76: y in [1, max(16, min(S, height)))
// stored at y + S/2  →  top = S/2 + hmax - 1
__global__ void heights_kernel(int S, float bounding, int32_t* __restrict__ heights, int32_t* __restrict__ top0) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
    if (z >= S) return;
    float fx = (0.75f * float(uint32_t(x))) * 0.01f, fz = (0.75f * float(uint32_t(z))) * 0.01f;   // GetNoise: x *= m_frequency
    float sum = simplex_octave(c_perm[0], fx, fz), amp = 1.0f;                                   // FBM, :1191-1207
    for (int o = 1; o < 3; ++o) {
        fx *= 2.0f; fz *= 2.0f;
        amp *= 0.5f;
        sum += simplex_octave(c_perm[o], fx, fz) * amp;
    }
    const int32_t h = int32_t(64.0f * (sum * bounding) + 32);
    heights[size_t(x) * S + z] = h;
    const int32_t hmax = max(16, min(S, h));
    top0[size_t(x) * S + z] = S / 2 + hmax - 1;
}

// the same fill rule for caller-supplied column heights (dynamic scenes)
__global__ void tops_kernel(int S, const int32_t* __restrict__ heights, int32_t* __restrict__ top0) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
    if (z >= S) return;
    const int32_t hmax = max(16, min(S / 2, heights[size_t(x) * S + z]));   // min(S, h) wherever y + S/2 stays inside the world
    top0[size_t(x) * S + z] = S / 2 + hmax - 1;
}

__global__ void pyramid_kernel(int n, const int32_t* __restrict__ lo, int32_t* __restrict__ hi) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
    if (z >= n) return;
    const size_t m = size_t(n) * 2;
    const int32_t a = lo[(2 * size_t(x)) * m + 2 * z], b = lo[(2 * size_t(x)) * m + 2 * z + 1];
    const int32_t c = lo[(2 * size_t(x) + 1) * m + 2 * z], d = lo[(2 * size_t(x) + 1) * m + 2 * z + 1];
    hi[size_t(x) * n + z] = max(max(a, b), max(c, d));
}

__global__ void count_kernel(int n2, int l, int ybase, const int32_t* __restrict__ top, uint32_t* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    cnt[i] = uint32_t((top[i] >> l) - ybase + 1);        // top >= bottom always (hmax >= 16)
}

// ---- exclusive scan (three small kernels; the arrays are at most 16 M entries) ----
constexpr int kScanBlock = 1024;
__global__ void scan_block_sums(const uint32_t* __restrict__ in, int n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t s[32];
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = s[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = w;
    }
}
__global__ void scan_sums_serial(uint32_t* sums, int nb, uint32_t* total) {      // nb <= 16384: one thread is enough
    uint32_t run = 0;
    for (int i = 0; i < nb; ++i) { const uint32_t v = sums[i]; sums[i] = run; run += v; }
    *total = run;
}
__global__ void scan_apply(const uint32_t* __restrict__ in, int n, const uint32_t* __restrict__ sums, uint32_t* __restrict__ out) {
    __shared__ uint32_t s[kScanBlock];
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    s[threadIdx.x] = i < n ? in[i] : 0u;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {           // Hillis-Steele inclusive scan
        const uint32_t v = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += v;
        __syncthreads();
    }
    if (i < n) out[i] = sums[blockIdx.x] + s[threadIdx.x] - in[i];
}

// node of level `lv` at column (x,z), run index yi: storage index, or 0xffffffff when the cube is empty
__device__ __forceinline__ uint32_t node_index(const Level& lv, int l, int x, int z, int yi) {
    const size_t c = size_t(x) * lv.n + z;
    if (yi < lv.ybase || yi > (lv.top[c] >> l)) return 0xffffffffu;
    return lv.off[c] + uint32_t(yi - lv.ybase);
}

// The three sweeps below give one WARP to a column of level l (blockDim = 32 x 4: lane = node of the column's run, threadIdx.y =
// column): the nodes of a column are stored consecutively, so a warp's loads and stores of size / self / cpos are coalesced.
// (One thread per column, walking its run alone, wrote 32 different sectors per warp store: 0.95 ms for the placement sweep of a
// 2048^3 world against 0.2-0.3 ms now; it matters for edits, which re-flatten the world.)
constexpr int kColumnsPerBlock = 4;
#define VRT_COLUMN_OF_WARP()                                                     \
    const int z = blockIdx.x * kColumnsPerBlock + threadIdx.y, x = blockIdx.y;   \
    if (z >= cur.n) return;                                                      \
    const size_t c = size_t(x) * cur.n + z;                                      \
    const int y_hi = cur.top[c] >> l

// bottom-up: size(N) = 8 + sum of interior children sizes
__global__ void size_kernel(Level cur, Level below, int l) {
    VRT_COLUMN_OF_WARP();
    for (int yi = cur.ybase + int(threadIdx.x); yi <= y_hi; yi += 32) {
        uint32_t sz = 8u;
        if (l > 1)
            for (int k = 0; k < 8; ++k) {
                const uint32_t ci = node_index(below, l - 1, 2 * x + (k & 1), 2 * z + ((k >> 2) & 1), 2 * yi + ((k >> 1) & 1));
                if (ci != 0xffffffffu) sz += below.size[ci];
            }
        cur.size[cur.off[c] + uint32_t(yi - cur.ybase)] = sz;
    }
}

// top-down: hand self()/child_pos() to the interior children, in compileSVO_rec's visit order (lsvo_utils.cpp:29-31)
__global__ void place_kernel(Level cur, Level below, int l) {
    VRT_COLUMN_OF_WARP();
    for (int yi = cur.ybase + int(threadIdx.x); yi <= y_hi; yi += 32) {
        const uint32_t me = cur.off[c] + uint32_t(yi - cur.ybase);
        const uint32_t P = cur.cpos[me];
        uint32_t running = P + 8u;
        for (int cx = 0; cx < 2; ++cx)
            for (int cy = 0; cy < 2; ++cy)
                for (int cz = 0; cz < 2; ++cz) {
                    const uint32_t ci = node_index(below, l - 1, 2 * x + cx, 2 * z + cz, 2 * yi + cy);
                    if (ci == 0xffffffffu) continue;
                    below.self[ci] = P + uint32_t(cz * 4 + cy * 2 + cx);          // slot = z*4 + y*2 + x, :34
                    below.cpos[ci] = running;
                    running += below.size[ci];
                }
    }
}

// writes the slot of every interior node of level l (all other slots keep the LNode() default written beforehand)
__global__ void emit_kernel(Level cur, Level below, int l, int bottom, const int32_t* __restrict__ top0, int S,
                            uint2* __restrict__ slots) {
    VRT_COLUMN_OF_WARP();
    for (int yi = cur.ybase + int(threadIdx.x); yi <= y_hi; yi += 32) {
        const uint32_t me = cur.off[c] + uint32_t(yi - cur.ybase);
        uint32_t mask = 0u;
        for (int k = 0; k < 8; ++k) {                      // k = slot = z*4 + y*2 + x
            const int cxx = 2 * x + (k & 1), czz = 2 * z + ((k >> 2) & 1), cyy = 2 * yi + ((k >> 1) & 1);
            bool present;
            if (l > 1) present = node_index(below, l - 1, cxx, czz, cyy) != 0xffffffffu;
            else present = cyy >= bottom && cyy <= top0[size_t(cxx) * S + czz];    // voxel solid
            if (present) mask |= 1u << k;
        }
        const uint32_t self = cur.self[me];
        slots[self] = make_uint2(1u | (mask << 8) | ((l == 1 ? mask : 0u) << 16), cur.cpos[me] - self);   // lsvo_utils.hpp:14-17
    }
}

__global__ void fill_default_kernel(uint2* __restrict__ slots, uint64_t n) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
        slots[i] = make_uint2(1u, 0u);                     // LNode(): color 1, masks 0, offset 0 (lsvo_utils.hpp:7-12)
}

__global__ void root_kernel(Level root) {                  // the root: slot 0, child block at 1 (lsvo_utils.hpp:49-52)
    root.self[0] = 0u;
    root.cpos[0] = 1u;
}

using Scratch = PoolScratch;

#define VRT_TRY(call)                          \
    do {                                       \
        cudaError_t e_ = (call);               \
        if (e_ != cudaSuccess) return e_;      \
    } while (0)

}  // namespace

// Builds T(depth) on the device — or, with d_heights_in, the same fill rule (main.cpp:70-76) over caller-supplied
// column heights.  *d_slots is cudaMalloc'ed (caller frees); optional d_heights_out [S*S] int32 receives the heights.
cudaError_t device_build_terrain_lsvo(int depth, uint2** d_slots, uint64_t* n_slots, int32_t* d_heights_out, cudaStream_t stream,
                                      const int32_t* d_heights_in, BuildPool* pool, uint64_t* capacity_slots) {
    const int S = 1 << depth, bottom = S / 2 + 1;
    Scratch sc(pool);
    int32_t* d_heights = nullptr;
    std::vector<int32_t*> top(depth + 1, nullptr);
    for (int l = 0; l <= depth; ++l) VRT_TRY(sc.alloc(&top[l], size_t(S >> l) * (S >> l)));
    const dim3 blk(128);
    if (d_heights_in) {
        d_heights = const_cast<int32_t*>(d_heights_in);
        tops_kernel<<<dim3((S + 127) / 128, S), blk, 0, stream>>>(S, d_heights_in, top[0]);
    } else {
        uint8_t perm[512], perm12[512];
        float bounding;
        host_simplex_tables(perm, perm12, &bounding);
        VRT_TRY(cudaMemcpyToSymbolAsync(c_perm, perm, 512, 0, cudaMemcpyHostToDevice, stream));
        VRT_TRY(cudaMemcpyToSymbolAsync(c_perm12, perm12, 512, 0, cudaMemcpyHostToDevice, stream));
        VRT_TRY(sc.alloc(&d_heights, size_t(S) * S));
        heights_kernel<<<dim3((S + 127) / 128, S), blk, 0, stream>>>(S, bounding, d_heights, top[0]);
    }
    for (int l = 1; l <= depth; ++l) {
        const int n = S >> l;
        pyramid_kernel<<<dim3((n + 127) / 128, n), blk, 0, stream>>>(n, top[l - 1], top[l]);
    }
    if (d_heights_out && d_heights_out != d_heights)
        VRT_TRY(cudaMemcpyAsync(d_heights_out, d_heights, size_t(S) * S * 4, cudaMemcpyDeviceToDevice, stream));

    // node counts and storage offsets per level
    std::vector<Level> lv(depth + 1);
    std::vector<uint32_t> totals(depth + 1, 0);
    uint32_t* d_total = nullptr;
    VRT_TRY(sc.alloc(&d_total, depth + 1));
    for (int l = 1; l <= depth; ++l) {
        const int n = S >> l, n2 = n * n, nb = (n2 + kScanBlock - 1) / kScanBlock;
        uint32_t *cnt = nullptr, *off = nullptr, *sums = nullptr;
        VRT_TRY(sc.alloc(&cnt, n2));
        VRT_TRY(sc.alloc(&off, n2));
        VRT_TRY(sc.alloc(&sums, nb));
        count_kernel<<<(n2 + 255) / 256, 256, 0, stream>>>(n2, l, bottom >> l, top[l], cnt);
        scan_block_sums<<<nb, kScanBlock, 0, stream>>>(cnt, n2, sums);
        scan_sums_serial<<<1, 1, 0, stream>>>(sums, nb, d_total + l);
        scan_apply<<<nb, kScanBlock, 0, stream>>>(cnt, n2, sums, off);
        lv[l].n = n; lv[l].ybase = bottom >> l; lv[l].top = top[l]; lv[l].off = off;
    }
    VRT_TRY(cudaMemcpyAsync(totals.data(), d_total, (depth + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VRT_TRY(cudaStreamSynchronize(stream));
    uint64_t interior = 0;
    for (int l = 1; l <= depth; ++l) {
        interior += totals[l];
        VRT_TRY(sc.alloc(&lv[l].size, totals[l]));
        VRT_TRY(sc.alloc(&lv[l].cpos, totals[l]));
        VRT_TRY(sc.alloc(&lv[l].self, totals[l]));
    }
    const uint64_t n = 1 + 8 * interior;                   // root slot + one 8-slot block per non-empty node
    if (n > 0xffffffffull) return cudaErrorInvalidValue;
    uint2* slots = nullptr;
    uint64_t capacity = n;
    if (pool && pool->spare && pool->spare_slots >= n) {   // build into the array the previous edit replaced
        slots = pool->spare;
        capacity = pool->spare_slots;
        pool->spare = nullptr;
        pool->spare_slots = 0;
    } else {
        if (pool) capacity = n + n / 16 + 4096;            // room for the next edits to grow into
        VRT_TRY(cudaMalloc(&slots, capacity * sizeof(uint2)));
    }
    fill_default_kernel<<<148 * 8, 256, 0, stream>>>(slots, n);

    lv[0] = Level{S, bottom, top[0], nullptr, nullptr, nullptr, nullptr};
    const dim3 wblk(32, kColumnsPerBlock);
    for (int l = 1; l <= depth; ++l)
        size_kernel<<<dim3((lv[l].n + kColumnsPerBlock - 1) / kColumnsPerBlock, lv[l].n), wblk, 0, stream>>>(lv[l], lv[l - 1], l);
    root_kernel<<<1, 1, 0, stream>>>(lv[depth]);
    for (int l = depth; l >= 2; --l)
        place_kernel<<<dim3((lv[l].n + kColumnsPerBlock - 1) / kColumnsPerBlock, lv[l].n), wblk, 0, stream>>>(lv[l], lv[l - 1], l);
    for (int l = 1; l <= depth; ++l)
        emit_kernel<<<dim3((lv[l].n + kColumnsPerBlock - 1) / kColumnsPerBlock, lv[l].n), wblk, 0, stream>>>(lv[l], lv[l - 1], l, bottom, top[0], S, slots);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !pool) e = cudaStreamSynchronize(stream);   // the work arrays are freed on return
    if (e != cudaSuccess) {
        if (pool) { if (pool->spare) cudaFree(pool->spare); pool->spare = slots; pool->spare_slots = capacity; }
        else cudaFree(slots);
        return e;
    }
    *d_slots = slots;
    *n_slots = n;
    if (capacity_slots) *capacity_slots = capacity;
    return cudaSuccess;
}

// ---- compaction: reference layout → breadth-first array of live nodes -------------------------------------------
namespace {

__global__ void compact_count_kernel(const uint2* __restrict__ ref, const uint32_t* __restrict__ list, uint32_t m,
                                     uint32_t* __restrict__ cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t raw = ref[list[i]].x;
    cnt[i] = __popc((raw >> 8) & ~(raw >> 16) & 0xffu);
}

__global__ void compact_emit_kernel(const uint2* __restrict__ ref, const uint32_t* __restrict__ list, uint32_t m,
                                    const uint32_t* __restrict__ off, uint32_t cur_base, uint32_t next_base,
                                    uint2* __restrict__ out, uint32_t* __restrict__ next_list) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t id = list[i];
    const uint2 w = ref[id];
    out[cur_base + i] = make_uint2(w.x, next_base + off[i]);
    uint32_t interior = (w.x >> 8) & ~(w.x >> 16) & 0xffu, k = off[i];
    while (interior) {
        const uint32_t s = __ffs(interior) - 1;
        interior &= interior - 1;
        next_list[k++] = id + w.y + s;                     // child slot in the reference layout: parent + child_offset + s
    }
}

}  // namespace

// Builds the compact array from a reference-layout array on the device.  *d_out is cudaMalloc'ed (caller frees).
cudaError_t device_compact_lsvo(const uint2* d_ref, uint64_t n_ref, int depth, uint2** d_out, uint64_t* n_out, cudaStream_t stream) {
    (void)depth;
    const uint64_t cap = (n_ref - 1) / 8 + 1;              // live nodes = root + one per 8-slot block at most
    Scratch sc(nullptr);
    uint2* out = nullptr;
    VRT_TRY(cudaMalloc(&out, cap * sizeof(uint2)));
    uint32_t *list_a = nullptr, *list_b = nullptr, *cnt = nullptr, *off = nullptr, *sums = nullptr, *d_total = nullptr;
    cudaError_t e = sc.alloc(&list_a, cap);
    if (e == cudaSuccess) e = sc.alloc(&list_b, cap);
    if (e == cudaSuccess) e = sc.alloc(&cnt, cap);
    if (e == cudaSuccess) e = sc.alloc(&off, cap);
    if (e == cudaSuccess) e = sc.alloc(&sums, cap / kScanBlock + 2);
    if (e == cudaSuccess) e = sc.alloc(&d_total, 1);
    if (e == cudaSuccess) e = cudaMemsetAsync(list_a, 0, sizeof(uint32_t), stream);     // level 0: the root, slot 0
    if (e != cudaSuccess) { cudaFree(out); return e; }
    uint32_t m = 1, base = 0;
    uint64_t total_nodes = 0;
    while (m > 0) {
        const uint32_t nb = (m + kScanBlock - 1) / kScanBlock;
        compact_count_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_ref, list_a, m, cnt);
        scan_block_sums<<<nb, kScanBlock, 0, stream>>>(cnt, int(m), sums);
        scan_sums_serial<<<1, 1, 0, stream>>>(sums, int(nb), d_total);
        scan_apply<<<nb, kScanBlock, 0, stream>>>(cnt, int(m), sums, off);
        uint32_t next_m = 0;
        e = cudaMemcpyAsync(&next_m, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess && uint64_t(base) + m + next_m > cap) e = cudaErrorInvalidValue;   // malformed input
        if (e != cudaSuccess) { cudaFree(out); return e; }
        compact_emit_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_ref, list_a, m, off, base, base + m, out, list_b);
        total_nodes += m;
        base += m;
        m = next_m;
        uint32_t* t = list_a; list_a = list_b; list_b = t;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { cudaFree(out); return e; }
    *d_out = out;
    *n_out = total_nodes;
    return cudaSuccess;
}


// ---- bounds of the solid voxels (for Trav2's kBounds walks) ------------------------------------------------------------
// A breadth-first sweep of the reference-layout array from the root: one queue entry per non-empty node (slot index + low
// corner in voxel units), expanded level by level on the device until the cubes are 8 voxels wide (or a level would not fit
// the queue): the min / max corners of those cubes are the bounds, at most 8 voxels loose — plenty for "the ray has left
// everything solid".  Scene construction, not the hot path: a handful of small launches.
namespace {
constexpr uint32_t kBoundsQueue = 1u << 19;                 // entries per queue (8 MB each)
struct BoundsEntry { uint32_t node, x, y, z; };

__global__ void bounds_expand_kernel(const uint2* __restrict__ slots, const BoundsEntry* __restrict__ in, const uint32_t* __restrict__ n_in_ptr,
                                     int level, int stop_level, BoundsEntry* __restrict__ out, uint32_t* __restrict__ n_out,
                                     int* __restrict__ box /* lo xyz (min), hi xyz (max) in voxels */) {
    const uint32_t n_in = *n_in_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
        const BoundsEntry e = in[i];
        const uint2 w = __ldg(slots + e.node);
        const uint32_t child_mask = (w.x >> 8) & 0xffu, leaf_mask = (w.x >> 16) & 0xffu;
        const uint32_t half = 1u << (level - 1);
        for (uint32_t s = 0; s < 8; ++s) {
            if (!((child_mask >> s) & 1u)) continue;
            // slot bit clear = upper half (the tree stores the mirrored octant, lsvo.hpp:79)
            const uint32_t x = e.x + ((s & 1u) ? 0u : half), y = e.y + ((s & 2u) ? 0u : half), z = e.z + ((s & 4u) ? 0u : half);
            bool terminal = ((leaf_mask >> s) & 1u) || level - 1 <= stop_level;
            if (!terminal) {
                const uint32_t at = atomicAdd(n_out, 1u);
                if (at < kBoundsQueue) out[at] = BoundsEntry{e.node + w.y + s, x, y, z};
                else terminal = true;                        // queue full: take the whole cube (still a bound)
            }
            if (terminal) {
                atomicMin(box + 0, int(x)); atomicMin(box + 1, int(y)); atomicMin(box + 2, int(z));
                atomicMax(box + 3, int(x + half)); atomicMax(box + 4, int(y + half)); atomicMax(box + 5, int(z + half));
            }
        }
    }
}
__global__ void bounds_clamp_count_kernel(uint32_t* n) { if (*n > kBoundsQueue) *n = kBoundsQueue; }
__global__ void bounds_finish_kernel(const int* __restrict__ box, int depth, float margin, float* __restrict__ out) {
    const float S = float(1 << depth);
    if (box[0] > box[3]) {                                  // nothing solid: an empty box in front of every ray
        for (int a = 0; a < 3; ++a) { out[a] = 1.5f; out[3 + a] = 1.5f; }
        return;
    }
    for (int a = 0; a < 3; ++a) {
        out[a] = fmaxf(1.0f, 1.0f + (float(box[a]) - margin) / S);
        out[3 + a] = fminf(2.0f, 1.0f + (float(box[3 + a]) + margin) / S);
    }
}
}  // namespace

size_t bounds_work_bytes() { return 2 * size_t(kBoundsQueue) * sizeof(BoundsEntry) + 64; }

cudaError_t device_scene_bounds(const uint2* d_nodes, int depth, float margin_voxels, float* d_bounds, void* d_work, cudaStream_t stream) {
    char* base = static_cast<char*>(d_work);
    BoundsEntry* q[2] = {reinterpret_cast<BoundsEntry*>(base), reinterpret_cast<BoundsEntry*>(base + size_t(kBoundsQueue) * sizeof(BoundsEntry))};
    uint32_t* counts = reinterpret_cast<uint32_t*>(base + 2 * size_t(kBoundsQueue) * sizeof(BoundsEntry));   // [0], [1]: queue sizes
    int* box = reinterpret_cast<int*>(counts + 2);
    const BoundsEntry root{0u, 0u, 0u, 0u};
    const uint32_t one = 1u;
    const int init_box[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};
    cudaError_t e = cudaMemcpyAsync(q[0], &root, sizeof(root), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, &one, sizeof(one), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(box, init_box, sizeof(init_box), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    const int stop_level = depth > 3 ? 3 : 0;               // cubes of 8 voxels
    int cur = 0;
    for (int level = depth; level >= 1; --level) {
        e = cudaMemsetAsync(counts + (cur ^ 1), 0, sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
        bounds_expand_kernel<<<296, 256, 0, stream>>>(d_nodes, q[cur], counts + cur, level, stop_level, q[cur ^ 1], counts + (cur ^ 1), box);
        bounds_clamp_count_kernel<<<1, 1, 0, stream>>>(counts + (cur ^ 1));
        cur ^= 1;
        if (level - 1 <= stop_level) break;
    }
    bounds_finish_kernel<<<1, 1, 0, stream>>>(box, depth, margin_voxels, d_bounds);
    return cudaGetLastError();
}

}  // namespace vrt

// Dense-grid traversal — kernels K2 (Grid3D), K2m (MipmapGrid3D) and K3 (pointer-SVO semantics).
//
//   K2   Grid3D<X,Y,Z>::castRay            reference include/grid_3d.hpp:35-132 (Amanatides-Woo DDA)
//   K2m  MipmapGrid3D                      reference include/mipmap_grid3D.hpp:14-17 is an empty stub; defined
//                                          here as "Grid3D::castRay results, bit-identical, with fewer fetches"
//   K3   SVO<N>::castRay + rec_castRay     reference include/svo.hpp:62-70,140-194 with fillHitResult's
//                                          commented body (svo.hpp:116-138) restored
//
// Device format: the reference stores an 8-byte Cell per voxel (1 GiB at 512^3); traversal only asks
// `type != Empty`, so the device keeps ONE BIT per cell, z-contiguous like m_cells[x][y][z] (16 MiB at
// 512^3 — L2 resident), plus an OR-pyramid of the same bit grids (level l = cubes of edge 2^l).
// The per-cell recurrence (t_max += t_d, one float add per step) is kept exactly as written, so every
// result is bit-identical; the pyramid is only used to SKIP FETCHES while the ray is inside a cube already
// known to be empty (a multi-cell jump would change the rounding of t_max, hence the hit on grazing rays).
#include "vrt_device.cuh"
#include "kernels.h"

namespace vrt {

__device__ __forceinline__ bool grid_bit(const GridLevels& g, int l, int x, int y, int z) {
    const GridLevel& L = g.level[l];
    const uint64_t i = (uint64_t(x >> l) * uint64_t(L.ny) + uint64_t(y >> l)) * uint64_t(L.nz) + uint64_t(z >> l);
    return (__ldg(L.bits + (i >> 5)) >> (i & 31u)) & 1u;
}

template <bool kMip>
__global__ void __launch_bounds__(256) grid_cast_kernel(GridLevels g, const float* __restrict__ origin, const float* __restrict__ dir,
                                                        uint64_t n, vrt_hit* __restrict__ out,
                                                        unsigned long long* __restrict__ counters) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t iter = 0u, fetches = 0u;
    if (i < n) {
        const float ox = origin[3 * i], oy = origin[3 * i + 1], oz = origin[3 * i + 2];
        const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        const int X = g.X, Y = g.Y, Z = g.Z;
        const float tdx = fabsf(1.0f / dx), tdy = fabsf(1.0f / dy), tdz = fabsf(1.0f / dz);       // grid_3d.hpp:42-44
        const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1, sz = dz < 0 ? -1 : 1;               // :48-50
        int cx = int(ox), cy = int(oy), cz = int(oz);                                              // :58-60
        float tmx = (float(cx + (sx > 0 ? 1 : 0)) - ox) / dx;                                      // :62-64
        float tmy = (float(cy + (sy > 0 ? 1 : 0)) - oy) / dy;
        float tmz = (float(cz + (sz > 0 ? 1 : 0)) - oz) / dz;
        // cube currently known to be empty (mip variant): level and coordinates at that level
        int e_level = -1, ex = 0, ey = 0, ez = 0;
        bool hit = false;
        int side = 0;
        float t = 0.0f;
        while (cx >= 0 && cy >= 0 && cz >= 0 && cx < X && cy < Y && cz < Z && iter < 2048u) {      // :68-70
            ++iter;
            side = (tmx < tmy) ? ((tmx < tmz) ? 0 : 2) : ((tmy < tmz) ? 1 : 2);                     // :73-99
            if (side == 0) { t = tmx; tmx += tdx; cx += sx; }
            else if (side == 1) { t = tmy; tmy += tdy; cy += sy; }
            else { t = tmz; tmz += tdz; cz += sz; }
            if (cx >= 0 && cy >= 0 && cz >= 0 && cx < X && cy < Y && cz < Z) {                     // :101
                bool solid;
                if (kMip) {
                    if (e_level >= 0 && (cx >> e_level) == ex && (cy >> e_level) == ey && (cz >> e_level) == ez) {
                        solid = false;                       // still inside the empty cube: no fetch
                    } else {
                        e_level = -1;
                        solid = true;
                        for (int l = g.n_levels - 1; l >= 0; --l) {   // coarse to fine, stop at the first empty cube
                            ++fetches;
                            if (!grid_bit(g, l, cx, cy, cz)) {
                                solid = false;
                                if (l > 0) { e_level = l; ex = cx >> l; ey = cy >> l; ez = cz >> l; }
                                break;
                            }
                        }
                    }
                } else {
                    ++fetches;
                    solid = grid_bit(g, 0, cx, cy, cz);                                            // :103-104
                }
                if (solid) { hit = true; break; }
            }
        }
        float4* q = reinterpret_cast<float4*>(out + i);
        if (hit) {
            const float hx = ox + t * dx, hy = oy + t * dy, hz = oz + t * dz;                      // :105-107
            float nx = 0.0f, ny = 0.0f, nz = 0.0f, u, v;
            if (side == 0) { nx = float(-sx); u = 1.0f - fracf(hz); v = fracf(hy); }               // :112-121
            else if (side == 1) { ny = float(-sy); u = fracf(hx); v = fracf(hz); }
            else { nz = float(-sz); u = fracf(hx); v = fracf(hy); }
            q[0] = make_float4(hx, hy, hz, t);
            q[1] = make_float4(nx, ny, nz, __uint_as_float(iter));                                 // complexity = iter, :124
            q[2] = make_float4(u, v, __uint_as_float(VRT_HIT_FLAG_HIT), 0.0f);
            q[3] = make_float4(__int_as_float(cx), __int_as_float(cy), __int_as_float(cz), __uint_as_float(1u << side));
        } else {
            // HitPoint::complexity stays 0 on a miss (only written on a hit, :124)
            q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // counters[0] += loop iterations (the unit of the DDA roofline), counters[1] += occupancy fetches
    for (int o = 16; o > 0; o >>= 1) {
        iter += __shfl_xor_sync(0xffffffffu, iter, o);
        fetches += __shfl_xor_sync(0xffffffffu, fetches, o);
    }
    if ((threadIdx.x & 31) == 0 && iter) {
        atomicAdd(counters, (unsigned long long)iter);
        atomicAdd(counters + 1, (unsigned long long)fetches);
    }
}

// ---- K3: SVO<N>::castRay (svo.hpp:62-70) ----------------------------------------------------------------
// The recursion of rec_castRay (svo.hpp:140-194) is unrolled into an explicit per-level frame; a node
// "exists" iff the occupancy pyramid bit of its cube is set, and it is a leaf iff its edge is one voxel.
struct SvoFrame {
    float px, py, pz;       // `position` argument of this level (re-based, :164)
    float cix, ciy, ciz;    // cell_pos_i
    float tmx, tmy, tmz;    // t_max
    float tmm;              // t_max_min
    float t_total;          // copy taken at entry (:151)
    int bx, by, bz;         // low corner of this node's cube in voxel units
};

__global__ void __launch_bounds__(128) svo_cast_kernel(GridLevels g, int depth, const float* __restrict__ origin,
                                                       const float* __restrict__ dir, uint32_t max_iter, uint64_t n,
                                                       vrt_hit* __restrict__ out, unsigned long long* __restrict__ counters) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t complexity = 0u;
    if (i < n) {
        const float sx0 = origin[3 * i], sy0 = origin[3 * i + 1], sz0 = origin[3 * i + 2];
        const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        // Ray, volumetric.hpp:28-40
        const float rtx = fabsf(1.0f / dx), rty = fabsf(1.0f / dy), rtz = fabsf(1.0f / dz);
        const float stx = dx >= 0.0f ? 1.0f : -1.0f, sty = copysignf(1.0f, dy), stz = copysignf(1.0f, dz);
        const float pdx = dx > 0.0f ? 1.0f : 0.0f, pdy = dy > 0.0f ? 1.0f : 0.0f, pdz = dz > 0.0f ? 1.0f : 0.0f;
        int side = 0;
        float ray_t_total = 0.0f;
        bool hit = false;
        float hit_t = 0.0f;
        int hvx = 0, hvy = 0, hvz = 0;

        SvoFrame st[13];
        int lvl = 0;                                   // frame index; children of frame `lvl` have edge 2^(depth-1-lvl)
        bool enter = true;                             // true: initialise frame `lvl` (function entry), false: resume after a return
        st[0].px = sx0; st[0].py = sy0; st[0].pz = sz0; st[0].bx = 0; st[0].by = 0; st[0].bz = 0;
        while (lvl >= 0) {
            SvoFrame& f = st[lvl];
            const int cl = depth - 1 - lvl;            // log2 of cell_size at this level
            const float cs = float(1u << cl);
            bool resume = !enter;                      // resuming after a child returned without a hit (:167-169)
            if (enter) {                                                                             // :142-151
                f.cix = float(int(f.px / cs)); f.ciy = float(int(f.py / cs)); f.ciz = float(int(f.pz / cs));
                f.cix = f.cix > 1.0f ? 1.0f : (f.cix < 0.0f ? 0.0f : f.cix);
                f.ciy = f.ciy > 1.0f ? 1.0f : (f.ciy < 0.0f ? 0.0f : f.ciy);
                f.ciz = f.ciz > 1.0f ? 1.0f : (f.ciz < 0.0f ? 0.0f : f.ciz);
                f.tmx = ((f.cix + pdx) * cs - f.px) / dx;
                f.tmy = ((f.ciy + pdy) * cs - f.py) / dy;
                f.tmz = ((f.ciz + pdz) * cs - f.pz) / dz;
                f.tmm = 0.0f;
                f.t_total = ray_t_total;
            }
            for (;;) {
                if (!resume) {
                    if (!(f.cix >= 0 && f.ciy >= 0 && f.ciz >= 0 && f.cix < 2 && f.ciy < 2 && f.ciz < 2 && complexity < max_iter)) {   // :152
                        --lvl; enter = false;          // return to the caller
                        break;
                    }
                    ++complexity;                                                                    // :154
                    const int cx = f.bx + int(uint32_t(f.cix)) * (1 << cl), cy = f.by + int(uint32_t(f.ciy)) * (1 << cl),
                              cz = f.bz + int(uint32_t(f.ciz)) * (1 << cl);
                    if (grid_bit(g, cl, cx, cy, cz)) {                                               // sub_node != nullptr, :156-157
                        if (cl == 0) {                                                               // leaf, :158-161
                            hit = true; hit_t = f.t_total + f.tmm; hvx = cx; hvy = cy; hvz = cz;
                            lvl = -1;
                            break;
                        }
                        SvoFrame& c = st[lvl + 1];                                                   // :163-166
                        c.px = (f.px + f.tmm * dx) - f.cix * cs;
                        c.py = (f.py + f.tmm * dy) - f.ciy * cs;
                        c.pz = (f.pz + f.tmm * dz) - f.ciz * cs;
                        c.bx = cx; c.by = cy; c.bz = cz;
                        ray_t_total = f.t_total + f.tmm;
                        ++lvl; enter = true;
                        break;
                    }
                }
                resume = false;
                const float tx = cs * rtx, ty = cs * rty, tz = cs * rtz;                             // :148
                if (f.tmx < f.tmy) {                                                                 // :173-192
                    if (f.tmx < f.tmz) { f.tmm = f.tmx; f.tmx += tx; f.cix += stx; side = 0; }
                    else { f.tmm = f.tmz; f.tmz += tz; f.ciz += stz; side = 2; }
                } else {
                    if (f.tmy < f.tmz) { f.tmm = f.tmy; f.tmy += ty; f.ciy += sty; side = 1; }
                    else { f.tmm = f.tmz; f.tmz += tz; f.ciz += stz; side = 2; }
                }
            }
        }
        float4* q = reinterpret_cast<float4*>(out + i);
        if (hit) {                                                                                   // fillHitResult, :116-138
            const float hx = sx0 + hit_t * dx, hy = sy0 + hit_t * dy, hz = sz0 + hit_t * dz;
            float nx = 0.0f, ny = 0.0f, nz = 0.0f, u, v;
            if (side == 0) { nx = -stx; u = 1.0f - fracf(hz); v = fracf(hy); }
            else if (side == 1) { ny = -sty; u = fracf(hx); v = fracf(hz); }
            else { nz = -stz; u = fracf(hx); v = fracf(hy); }
            q[0] = make_float4(hx, hy, hz, hit_t);
            q[1] = make_float4(nx, ny, nz, __uint_as_float(complexity));
            q[2] = make_float4(u, v, __uint_as_float(VRT_HIT_FLAG_HIT), 0.0f);
            q[3] = make_float4(__int_as_float(hvx), __int_as_float(hvy), __int_as_float(hvz), __uint_as_float(1u << side));
        } else {
            q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[1] = make_float4(0.f, 0.f, 0.f, __uint_as_float(complexity));   // complexity counts on a miss too (:154)
            q[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (int o = 16; o > 0; o >>= 1) complexity += __shfl_xor_sync(0xffffffffu, complexity, o);
    if ((threadIdx.x & 31) == 0 && complexity) atomicAdd(counters, (unsigned long long)complexity);
}

// ---- grid frames: RayCaster semantics over a dense grid, with mirror reflections (extension) ----
struct DdaHit {
    bool hit;
    int side, cx, cy, cz;
    float t;
    uint32_t steps;
};

// Grid3D::castRay (grid_3d.hpp:35-132) as a device function; same recurrence as grid_cast_kernel.
template <bool kMip>
__device__ __forceinline__ void grid_dda(const GridLevels& g, float ox, float oy, float oz, float dx, float dy, float dz, DdaHit& r) {
    const int X = g.X, Y = g.Y, Z = g.Z;
    const float tdx = fabsf(1.0f / dx), tdy = fabsf(1.0f / dy), tdz = fabsf(1.0f / dz);
    const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1, sz = dz < 0 ? -1 : 1;
    int cx = int(ox), cy = int(oy), cz = int(oz);
    float tmx = (float(cx + (sx > 0 ? 1 : 0)) - ox) / dx;
    float tmy = (float(cy + (sy > 0 ? 1 : 0)) - oy) / dy;
    float tmz = (float(cz + (sz > 0 ? 1 : 0)) - oz) / dz;
    int e_level = -1, ex = 0, ey = 0, ez = 0;
    uint32_t iter = 0u;
    int side = 0;
    float t = 0.0f;
    bool hit = false;
    while (cx >= 0 && cy >= 0 && cz >= 0 && cx < X && cy < Y && cz < Z && iter < 2048u) {
        ++iter;
        side = (tmx < tmy) ? ((tmx < tmz) ? 0 : 2) : ((tmy < tmz) ? 1 : 2);
        if (side == 0) { t = tmx; tmx += tdx; cx += sx; }
        else if (side == 1) { t = tmy; tmy += tdy; cy += sy; }
        else { t = tmz; tmz += tdz; cz += sz; }
        if (cx >= 0 && cy >= 0 && cz >= 0 && cx < X && cy < Y && cz < Z) {
            bool solid;
            if (kMip) {
                if (e_level >= 0 && (cx >> e_level) == ex && (cy >> e_level) == ey && (cz >> e_level) == ez) {
                    solid = false;
                } else {
                    e_level = -1;
                    solid = true;
                    for (int l = g.n_levels - 1; l >= 0; --l)
                        if (!grid_bit(g, l, cx, cy, cz)) {
                            solid = false;
                            if (l > 0) { e_level = l; ex = cx >> l; ey = cy >> l; ez = cz >> l; }
                            break;
                        }
                }
            } else {
                solid = grid_bit(g, 0, cx, cy, cz);
            }
            if (solid) { hit = true; break; }
        }
    }
    r.hit = hit; r.side = side; r.cx = cx; r.cy = cy; r.cz = cz; r.t = t; r.steps = iter;
}

// Grid3D::castRay (grid_3d.hpp:35-132) on the bordered bit grid (GridLevels::pad_bits) — the default DDA.
// ncu on the first kernels (profiles/r02_ncu_summaries.txt, grid_cast_kernel): issue slots 86 % busy, ALU pipe 74-83 %,
// math-pipe-throttle the top stall — the loop is bound by integer / compare instructions, ~75 of them per cell step:
// twelve bounds compares, 64-bit index multiplies, a copy of t per step.  Here a step is ~19 instructions:
//   * the cell is ONE 32-bit linear bit index, stepped by a per-axis stride (two's complement for negative steps);
//   * no bounds test at all: a ray that leaves the grid lands on a border cell, which is set — the loop ends on it like
//     on a solid cell, and the decoded coordinate tells the two apart afterwards (once per ray);
//   * `t_max += t_d` of the stepped axis (:78,86,94) is applied AFTER the occupancy test of the trip, so when the loop stops
//     the un-incremented t_max of the stepped axis IS the hit distance (:77,85,93) — nothing is copied per step;
//   * trips are counted with an FADD (FMA pipe; exact below 2^24);
//   * the 2048-iteration cap (:70) is compiled in only for grids a ray can cross in 2048 steps or more (kCap).
// The float recurrence and the comparisons are the reference's, so records are byte-identical to the generic loop above.
template <bool kCap>
__device__ __forceinline__ void grid_dda_fast(const GridLevels& g, float ox, float oy, float oz, float dx, float dy, float dz, DdaHit& r) {
    const int X = g.X, Y = g.Y, Z = g.Z;
    const float tdx = fabsf(1.0f / dx), tdy = fabsf(1.0f / dy), tdz = fabsf(1.0f / dz);       // grid_3d.hpp:42-44
    const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1, sz = dz < 0 ? -1 : 1;               // :48-50
    int cx = int(ox), cy = int(oy), cz = int(oz);                                              // :58-60
    float tmx = (float(cx + (sx > 0 ? 1 : 0)) - ox) / dx;                                      // :62-64
    float tmy = (float(cy + (sy > 0 ? 1 : 0)) - oy) / dy;
    float tmz = (float(cz + (sz > 0 ? 1 : 0)) - oz) / dz;
    r.hit = false; r.side = 0; r.cx = cx; r.cy = cy; r.cz = cz; r.t = 0.0f; r.steps = 0u;
    if (!(cx >= 0 && cy >= 0 && cz >= 0 && cx < X && cy < Y && cz < Z)) return;               // :68-69: the loop is not entered
    const uint32_t PY = g.pad_y, PZ = g.pad_z;
    uint32_t idx = (uint32_t(cx + 1) * PY + uint32_t(cy + 1)) * PZ + uint32_t(cz + 1);
    // strides and the grid pointer pinned in registers (ptxas otherwise re-derives them from the constant bank every trip)
    const uint32_t dix = uint32_t(sx) * PY * PZ + blockIdx.y, diy = uint32_t(sy) * PZ + blockIdx.y, diz = uint32_t(sz) + blockIdx.y;
    const uint32_t* __restrict__ bits = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(g.pad_bits) + blockIdx.y + threadIdx.z);
    // The loop is bound by the half-rate ALU pipe (ncu: ALU 82 %, FMA 14 %): compares, selects, integer adds, shifts and logic all
    // issue there.  The step is therefore written with explicit predication: the chosen axis' t_max grows by a predicated FADD and the
    // cell index moves by a predicated IMAD (`stride * one + idx`, `one` an opaque 1: both FMA pipe) instead of selects and an add;
    // the hit distance — the stepped axis' t_max before its increment (:77,85,93), which is the smallest of the three by the very
    // comparisons that chose the axis (ties hold equal values) — is a min3; the axis of the last step is recovered after the loop
    // from the index difference.  The fp32 operations are those of grid_3d.hpp:73-99, in order.
    const uint32_t one = 1u + blockIdx.y;
    float it = 0.0f, t_old = 0.0f;
    uint32_t idx_prev = idx;
    bool stopped = false;
    for (;;) {
        if (kCap) { if (!(it < 2048.0f)) break; }                                              // :70
        it += 1.0f;
        asm volatile(
            "{\n\t"
            ".reg .pred q, px, py, pn;\n\t"
            ".reg .f32 m;\n\t"
            "setp.lt.f32 q, %0, %1;\n\t"                  // t_max_x < t_max_y                               :73
            "setp.lt.and.f32 px, %0, %2, q;\n\t"          //   ... and t_max_x < t_max_z: step x             :74
            "setp.lt.and.f32 py, %1, %2, !q;\n\t"         // else t_max_y < t_max_z: step y                  :84
            "or.pred pn, px, py;\n\t"                     // neither: step z
            "min.f32 m, %1, %2;\n\t"
            "min.f32 %4, %0, m;\n\t"                      // the stepped axis' t_max before its increment
            "mov.u32 %5, %3;\n\t"
            "@px add.f32 %0, %0, %6;\n\t"                 // :78
            "@px mad.lo.u32 %3, %9, %12, %3;\n\t"
            "@py add.f32 %1, %1, %7;\n\t"                 // :86
            "@py mad.lo.u32 %3, %10, %12, %3;\n\t"
            "@!pn add.f32 %2, %2, %8;\n\t"                // :94
            "@!pn mad.lo.u32 %3, %11, %12, %3;\n\t"
            "}"
            : "+f"(tmx), "+f"(tmy), "+f"(tmz), "+r"(idx), "=f"(t_old), "=r"(idx_prev)
            : "f"(tdx), "f"(tdy), "f"(tdz), "r"(dix), "r"(diy), "r"(diz), "r"(one));
        if (__ldg(bits + (idx >> 5)) & (1u << (idx & 31u))) { stopped = true; break; }         // :103-104, or the border
    }
    const uint32_t moved = idx - idx_prev;                                                     // strides differ: PY * PZ > PZ > 1
    const bool p0 = moved == dix * one, p1 = moved == diy * one;
    r.steps = uint32_t(it);
    if (!stopped) return;
    const uint32_t pz = idx % PZ, q = idx / PZ, py = q % PY, px = q / PY;
    cx = int(px) - 1; cy = int(py) - 1; cz = int(pz) - 1;
    if (cx < 0 || cy < 0 || cz < 0 || cx >= X || cy >= Y || cz >= Z) return;                   // left the grid: miss
    r.hit = true;
    r.side = p0 ? 0 : (p1 ? 1 : 2);
    r.t = t_old;                                                                               // :77,85,93
    r.cx = cx; r.cy = cy; r.cz = cz;
}

// kMode: 0 = generic loop on the flat grid, 1 = generic loop with the fetch-skipping pyramid, 2 = bordered grid, 3 = bordered + cap
template <int kMode>
__device__ __forceinline__ void grid_dda_any(const GridLevels& g, float ox, float oy, float oz, float dx, float dy, float dz, DdaHit& r) {
    if (kMode == 0) grid_dda<false>(g, ox, oy, oz, dx, dy, dz, r);
    else if (kMode == 1) grid_dda<true>(g, ox, oy, oz, dx, dy, dz, r);
    else if (kMode == 2) grid_dda_fast<false>(g, ox, oy, oz, dx, dy, dz, r);
    else grid_dda_fast<true>(g, ox, oy, oz, dx, dy, dz, r);
}

// K2f: Grid3D::castRay over a ray buffer on the bordered grid; same 64-byte records as grid_cast_kernel
template <bool kCap>
__global__ void __launch_bounds__(256) grid_cast_fast_kernel(GridLevels g, const float* __restrict__ origin, const float* __restrict__ dir,
                                                             uint64_t n, vrt_hit* __restrict__ out,
                                                             unsigned long long* __restrict__ counters) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t iter = 0u;
    if (i < n) {
        const float ox = origin[3 * i], oy = origin[3 * i + 1], oz = origin[3 * i + 2];
        const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        DdaHit h;
        grid_dda_fast<kCap>(g, ox, oy, oz, dx, dy, dz, h);
        iter = h.steps;
        float4* q = reinterpret_cast<float4*>(out + i);
        if (h.hit) {
            const float t = h.t;
            const float hx = ox + t * dx, hy = oy + t * dy, hz = oz + t * dz;                  // grid_3d.hpp:105-107
            float nx = 0.0f, ny = 0.0f, nz = 0.0f, u, v;
            if (h.side == 0) { nx = float(dx < 0 ? 1 : -1); u = 1.0f - fracf(hz); v = fracf(hy); }   // :112-121
            else if (h.side == 1) { ny = float(dy < 0 ? 1 : -1); u = fracf(hx); v = fracf(hz); }
            else { nz = float(dz < 0 ? 1 : -1); u = fracf(hx); v = fracf(hy); }
            q[0] = make_float4(hx, hy, hz, t);
            q[1] = make_float4(nx, ny, nz, __uint_as_float(iter));                             // complexity = iter, :124
            q[2] = make_float4(u, v, __uint_as_float(VRT_HIT_FLAG_HIT), 0.0f);
            q[3] = make_float4(__int_as_float(h.cx), __int_as_float(h.cy), __int_as_float(h.cz), __uint_as_float(1u << h.side));
        } else {
            q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (int o = 16; o > 0; o >>= 1) iter += __shfl_xor_sync(0xffffffffu, iter, o);
    if ((threadIdx.x & 31) == 0 && iter) {
        atomicAdd(counters, (unsigned long long)iter);
        atomicAdd(counters + 1, (unsigned long long)iter);       // one occupancy fetch per step
    }
}

__device__ __forceinline__ uint8_t mul_u8g(uint8_t c, float f) { return uint8_t(fminf(255.0f, float(c) * f)); }   // utils.cpp:43-48

// 12 CTAs per SM (40 registers): the DDA loops need few registers and the kernel waits on instruction latency — left to itself
// ptxas takes 72 registers (7 CTAs per SM) and still spills 24 bytes.  cfg-3 frame (4K, mirror lake), same box: unbounded / 8 / 10 / 12
// CTAs per SM = 13.4 / 9.46 / 8.61 / 8.48 ms (profiles/r02_ab_grid_k4.txt).
#ifndef VRT_GRID_RENDER_MIN_CTAS
#define VRT_GRID_RENDER_MIN_CTAS 12
#endif
template <int kMode>
__global__ void __launch_bounds__(128, VRT_GRID_RENDER_MIN_CTAS) grid_render_kernel(GridLevels g, RenderLaunch L, uint32_t* __restrict__ accum,
                                                          unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_x = (L.width + 31) / 32;
    const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
    const int x = bx * 32 + warp * 8 + (lane & 7);
    const int y = L.row_begin + (by * L.tile_step + L.tile_index) * 4 + (lane >> 3);
    uint32_t n_rays[3] = {0, 0, 0}, n_steps[3] = {0, 0, 0};      // primary, shadow, reflection
    if (x < L.width && y < L.row_end && (!L.checker || (x & 1) == checker_x_parity(L.checker, L.checker_area_height, y))) {
        const float aspect = float(L.width) / float(L.height);
        const float lens_x = float(x) / float(L.height) - aspect * 0.5f, lens_y = float(y) / float(L.height) - 0.5f;
        const uint32_t pixel = uint32_t(y) * uint32_t(L.width) + uint32_t(x);
        uint32_t sum_r = 0, sum_g = 0, sum_b = 0;
        for (int s = 0; s < L.spp; ++s) {
            const uint32_t sample = uint32_t(L.sample_offset + s);
            float ox, oy, oz, dx, dy, dz;
            {   // Camera::getRay (camera_controller.hpp:34-49) in voxel units
                const uint4 rnd0 = philox4x32_10(pixel, sample, 0u, 0u, L.seed_lo, L.seed_hi);
                const float u0 = lattice(rnd0.x, -0.5f, 0.5f), u1 = lattice(rnd0.y, -0.5f, 0.5f);
                float fx = lens_x, fy = lens_y, fz = L.cam.fov;
                normalize3(fx, fy, fz);
                fx *= L.cam.focal_length; fy *= L.cam.focal_length; fz *= L.cam.focal_length;
                const float rx = L.cam.aperture * u0, ry = L.cam.aperture * u1, rz = L.cam.aperture * 0.0f;
                float qx = fx - rx, qy = fy - ry, qz = fz - rz;
                normalize3(qx, qy, qz);
                const float* m = L.cam.rot_mat;
                dx = (m[0] * qx + m[1] * qy) + m[2] * qz; dy = (m[3] * qx + m[4] * qy) + m[5] * qz; dz = (m[6] * qx + m[7] * qy) + m[8] * qz;
                ox = L.cam.position[0] + ((m[0] * rx + m[1] * ry) + m[2] * rz);
                oy = L.cam.position[1] + ((m[3] * rx + m[4] * ry) + m[5] * rz);
                oz = L.cam.position[2] + ((m[6] * rx + m[7] * ry) + m[8] * rz);
            }
            float tint = 1.0f;
            int bounds = 0;
            for (;;) {
                DdaHit h;
                grid_dda_any<kMode>(g, ox, oy, oz, dx, dy, dz, h);
                const int cls = bounds == 0 ? 0 : 2;
                n_rays[0] += cls == 0; n_rays[2] += cls == 2;
                n_steps[0] += cls == 0 ? h.steps : 0u; n_steps[2] += cls == 2 ? h.steps : 0u;
                if (!h.hit) break;
                const float hx = ox + h.t * dx, hy = oy + h.t * dy, hz = oz + h.t * dz;          // grid_3d.hpp:105-107
                float nx = 0.0f, ny = 0.0f, nz = 0.0f, u, v;
                if (h.side == 0) { nx = float(dx < 0 ? 1 : -1); u = 1.0f - fracf(hz); v = fracf(hy); }   // :112-121
                else if (h.side == 1) { ny = float(dy < 0 ? 1 : -1); u = fracf(hx); v = fracf(hz); }
                else { nz = float(dz < 0 ? 1 : -1); u = fracf(hx); v = fracf(hy); }
                const uint64_t ci = (uint64_t(h.cx) * uint64_t(g.Y) + uint64_t(h.cy)) * uint64_t(g.Z) + uint64_t(h.cz);
                const bool mirror = g.mirror && ((__ldg(g.mirror + (ci >> 5)) >> (ci & 31u)) & 1u);
                if (mirror && bounds < L.max_bounds) {
                    ox = hx + nx * 0.001f; oy = hy + ny * 0.001f; oz = hz + nz * 0.001f;
                    if (h.side == 0) dx = -dx; else if (h.side == 1) dy = -dy; else dz = -dz;
                    // dimensions 8+3b, 9+3b, 10+3b of the sample's Philox stream
                    float r[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t dim = 8u + 3u * uint32_t(bounds) + uint32_t(k);
                        const uint4 w = philox4x32_10(pixel, sample, dim >> 2, 0u, L.seed_lo, L.seed_hi);
                        const uint32_t word = (dim & 3u) == 0u ? w.x : ((dim & 3u) == 1u ? w.y : ((dim & 3u) == 2u ? w.z : w.w));
                        r[k] = lattice(word, -0.5f, 0.5f);
                    }
                    dx = dx + L.roughness * r[0]; dy = dy + L.roughness * r[1]; dz = dz + L.roughness * r[2];
                    normalize3(dx, dy, dz);
                    tint = tint * 0.8f;
                    ++bounds;
                    continue;
                }
                const uint8_t* tex = (ny != 0.0f) ? L.tex_top : L.tex_side;
                u = fminf(fmaxf(u, 0.0f), 1.0f); v = fminf(fmaxf(v, 0.0f), 1.0f);
                const uint32_t tx = min(15u, uint32_t(16.0f * u)), ty = min(15u, uint32_t(16.0f * v));
                const uint8_t* texel = tex + 3u * (ty * 16u + tx);
                const float sox = hx + nx * 0.001f, soy = hy + ny * 0.001f, soz = hz + nz * 0.001f;
                float tlx = L.light[0] - sox, tly = L.light[1] - soy, tlz = L.light[2] - soz;
                normalize3(tlx, tly, tlz);
                DdaHit sh;
                grid_dda_any<kMode>(g, sox, soy, soz, tlx, tly, tlz, sh);
                n_rays[1] += 1u; n_steps[1] += sh.steps;
                float light = 0.0f;
                if (!sh.hit) light = fmaxf(0.0f, dot3(tlx, tly, tlz, nx, ny, nz));
                const float f = fminf(1.0f, fmaxf(0.0f, light));
                sum_r += mul_u8g(mul_u8g(__ldg(texel), f), tint);
                sum_g += mul_u8g(mul_u8g(__ldg(texel + 1), f), tint);
                sum_b += mul_u8g(mul_u8g(__ldg(texel + 2), f), tint);
                break;
            }
        }
        uint4* a = reinterpret_cast<uint4*>(accum) + pixel;
        uint4 v4 = *a;
        v4.x += sum_r; v4.y += sum_g; v4.z += sum_b; v4.w += uint32_t(L.spp);
        *a = v4;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t a = n_rays[k], b = n_steps[k];
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0 && a) { atomicAdd(counters + k, (unsigned long long)a); atomicAdd(counters + 6 + k, (unsigned long long)b); }
    }
}

cudaError_t launch_grid_render(const GridLevels& g, bool use_mip, const RenderLaunch& L, uint32_t* d_accum,
                               unsigned long long* d_counters, cudaStream_t stream) {
    const int rows = L.row_end - L.row_begin;
    if (rows <= 0 || L.width <= 0 || L.spp <= 0) return cudaSuccess;
    const int tiles_x = (L.width + 31) / 32, tiles_y = ((rows + 3) / 4 + L.tile_step - 1 - L.tile_index) / L.tile_step;
    if (tiles_y <= 0) return cudaSuccess;
    const unsigned grid = unsigned(tiles_x) * unsigned(tiles_y);
    const bool cap = g.X + g.Y + g.Z >= 2048;
    if (g.pad_bits && L.grid_variant == 0) {
        if (cap) grid_render_kernel<3><<<grid, 128, 0, stream>>>(g, L, d_accum, d_counters);
        else grid_render_kernel<2><<<grid, 128, 0, stream>>>(g, L, d_accum, d_counters);
    } else if (use_mip && g.n_levels > 1) grid_render_kernel<1><<<grid, 128, 0, stream>>>(g, L, d_accum, d_counters);
    else grid_render_kernel<0><<<grid, 128, 0, stream>>>(g, L, d_accum, d_counters);
    return cudaGetLastError();
}

cudaError_t launch_grid_cast(const GridLevels& g, bool use_mip, int variant, const float* d_origin, const float* d_dir, uint64_t n,
                             vrt_hit* d_out, unsigned long long* d_counters, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    const unsigned grid = unsigned((n + 255) / 256);
    if (g.pad_bits && variant == 0) {                     // the bordered grid: Grid3D and MipmapGrid3D alike (identical records)
        if (g.X + g.Y + g.Z >= 2048) grid_cast_fast_kernel<true><<<grid, 256, 0, stream>>>(g, d_origin, d_dir, n, d_out, d_counters);
        else grid_cast_fast_kernel<false><<<grid, 256, 0, stream>>>(g, d_origin, d_dir, n, d_out, d_counters);
        return cudaGetLastError();
    }
    if (use_mip && g.n_levels > 1) grid_cast_kernel<true><<<grid, 256, 0, stream>>>(g, d_origin, d_dir, n, d_out, d_counters);
    else grid_cast_kernel<false><<<grid, 256, 0, stream>>>(g, d_origin, d_dir, n, d_out, d_counters);
    return cudaGetLastError();
}

cudaError_t launch_svo_cast(const GridLevels& g, int depth, const float* d_origin, const float* d_dir, uint32_t max_iter,
                            uint64_t n, vrt_hit* d_out, unsigned long long* d_counters, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    svo_cast_kernel<<<unsigned((n + 127) / 128), 128, 0, stream>>>(g, depth, d_origin, d_dir, max_iter, n, d_out, d_counters);
    return cudaGetLastError();
}

}  // namespace vrt

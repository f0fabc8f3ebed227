// Beam floors: a conservative start distance for the primary rays of every screen tile.
//
// Primary rays are 41 % of all loop trips of the headline frame (42 trips per ray): most of them walk through the empty air
// between the camera and the terrain.  LSVO<D>::castRay (lsvo.hpp:33-172) starts at t_min = max(0, entry into the root cube)
// (:54-57); starting it at any t_floor <= (the ray's hit distance) gives the SAME HitPoint — position, distance, normal,
// voxel_coord are functions of the hit cell and of the plane through which the ray enters it, not of the path that led
// there — as long as nothing solid lies before t_floor.  Only HitPoint::complexity (the trip count) gets smaller.
//
// So for every 8x8-pixel tile this kernel computes a RIGOROUS lower bound of the hit distance of every primary ray of the
// tile (every pixel of the tile, every lens sample): the distance from the camera to the nearest non-empty octree node that
// intersects the tile's frustum — a front-to-back search of the octree against the four side planes of the frustum (and a
// fifth plane in front of the camera).
//   * a ray can only hit voxels, every voxel lies in the non-empty nodes of every level, and the ray stays inside the
//     frustum of its tile, so the node it hits is among those the search sees (box-vs-plane tests only ever err towards
//     "intersects");
//   * depth of field: a lens sample starts at camera + r, |r| <= aperture / sqrt(2), and aims at the focal point of its
//     pixel (camera_controller.hpp:37-42); at distance t it is within |r| * (1 + t / focal_length) of the pixel's centre
//     ray, so the planes are moved outwards by that amount (evaluated at the node's far corner);
//   * the search stops descending at nodes smaller than half the tile's footprint at their distance (a finer bound buys
//     little) and the floor is the nearest such node's box distance minus the lens radius and two voxels of slack; every ray
//     lowers it further by the stretch over which its own plane crossings are numerically uncertain (render_chain.cuh,
//     beam_floor_of).
// Unlike the cone-marching beam optimisation of Laine & Karras (which follows only the tile's corner rays and can step over
// geometry that pokes into the beam between them), this bound holds for arbitrary voxel sets: frames are byte-identical with
// and without it (tests/test_gpu_render.py::test_beam_floors_do_not_change_frames).
#include "lsvo_traverse.cuh"
#include "kernels.h"

namespace vrt {

namespace {

constexpr int kBeamThreads = 128;                 // 16 tiles per block: 8 lanes per tile
constexpr int kBeamStack = 88;                    // <= 7 pushes per level, depth <= 12

}  // namespace

// Eight lanes per tile: lane `sub` of a group tests child `sub` of the node the group has popped, so the eight box-vs-frustum
// tests of a node run side by side instead of one after the other (one thread per tile took 0.30 ms on the headline frame whatever
// the number of tiles: the time of the longest search, one dependent test after the other).  The group's stack lives in shared
// memory; the children that have to be searched are pushed farthest first, so the nearest is popped next (front to back).
// floor[ty * tiles_x + tx] = t (normalised units, as castRay measures it) below which no primary ray of the tile can hit anything.
__global__ void __launch_bounds__(kBeamThreads) beam_floor_kernel(const uint2* __restrict__ slots, RenderLaunch L, int tile, int tiles_x,
                                                                  int tiles_y, int tile_y0, float* __restrict__ floor) {
    __shared__ uint4 stacks[kBeamThreads / 8][kBeamStack];      // x | y << 16, z | level << 16, node, bits of d_min
    const int group = threadIdx.x >> 3, sub = threadIdx.x & 7, lane = threadIdx.x & 31;
    const unsigned gmask = 0xffu << (lane & 24);
    const int id = blockIdx.x * (kBeamThreads / 8) + group;
    if (id >= tiles_x * tiles_y) return;
    const int tx = id % tiles_x, ty = tile_y0 + id / tiles_x;
    if (L.tile_step > 1) {                                 // multi-GPU tile split: only tiles that contain rows this rank renders
        bool mine = false;
        for (int row = ty * tile; row < ty * tile + tile; row += 4)
            mine = mine || (((row - L.row_begin) >> 2) % L.tile_step == L.tile_index);
        if (tile < 4) mine = ((ty * tile - L.row_begin) >> 2) % L.tile_step == L.tile_index;
        if (!mine) return;
    }
    const float S = float(1 << L.depth);
    // camera centre in voxel units of the castRay cube: (position * SCALE + 1 - 1) * S = position
    const float cx = L.cam.position[0], cy = L.cam.position[1], cz = L.cam.position[2];
    const float aspect = float(L.width) / float(L.height);
    // the four corner directions of the tile in world space (not normalised): pixel (x, y) looks along (x/H - aspect/2, y/H - 1/2, fov)
    float dir[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float px = float(tx * tile + ((k & 1) ? tile : 0)), py = float(ty * tile + ((k & 2) ? tile : 0));
        const float lx = px / float(L.height) - aspect * 0.5f, ly = py / float(L.height) - 0.5f, lz = L.cam.fov;
        const float* m = L.cam.rot_mat;
        dir[k][0] = (m[0] * lx + m[1] * ly) + m[2] * lz;
        dir[k][1] = (m[3] * lx + m[4] * ly) + m[5] * lz;
        dir[k][2] = (m[6] * lx + m[7] * ly) + m[8] * lz;
    }
    // inward unit normals of the side planes: left (00,01), right (11,10), top (10,00), bottom (01,11), oriented by the opposite corner
    float pn[5][3];
    const int pa[4] = {0, 3, 1, 2}, pb[4] = {2, 1, 0, 3}, inside[4] = {1, 0, 2, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* a = dir[pa[k]];
        const float* b = dir[pb[k]];
        float nx = a[1] * b[2] - a[2] * b[1], ny = a[2] * b[0] - a[0] * b[2], nz = a[0] * b[1] - a[1] * b[0];
        const float* q = dir[inside[k]];
        if (nx * q[0] + ny * q[1] + nz * q[2] < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
        const float inv = rsqrtf(nx * nx + ny * ny + nz * nz);
        pn[k][0] = nx * inv; pn[k][1] = ny * inv; pn[k][2] = nz * inv;
    }
    // fifth plane: in front of the camera.  Without it a box BEHIND the camera passes the side-plane tests whenever the frustum
    // there is narrower than the lens slack (small tiles): sky tiles would "see" the ground behind the camera and lose their skip.
    // Every ray point is r + t * u with t >= 0 and u within the tile: dot(fwd, p) >= -|r| for all of them.
    {
        const float fx = dir[0][0] + dir[1][0] + dir[2][0] + dir[3][0], fy = dir[0][1] + dir[1][1] + dir[2][1] + dir[3][1],
                    fz = dir[0][2] + dir[1][2] + dir[2][2] + dir[3][2];
        const float inv = rsqrtf(fx * fx + fy * fy + fz * fz);
        pn[4][0] = fx * inv; pn[4][1] = fy * inv; pn[4][2] = fz * inv;
    }
    // lens radius (voxels) and its growth with distance, with a little slack for the rounding of everything above
    const float lens = fabsf(L.cam.aperture) * 0.70710678f * 1.01f + 0.01f;
    const float focal = fmaxf(L.focal ? __ldg(L.focal) : L.cam.focal_length, 1e-3f);
    const float lens_growth = lens / focal;
    // footprint of the tile per unit distance (voxels per voxel): its diagonal
    const float spread = float(tile) * 1.4142136f / float(L.height);

    uint4* stack = stacks[group];
    int sp = 1;
    if (sub == 0) stack[0] = make_uint4(0u, uint32_t(L.depth) << 16, 0u, 0u);
    __syncwarp(gmask);
    float best = 3.0e9f;                                   // nearest qualifying node so far (voxels); the same in all lanes of a group
    while (sp > 0) {
        const uint4 f = stack[--sp];
        __syncwarp(gmask);                                 // everyone has read the frame before anyone overwrites its slot
        if (__uint_as_float(f.w) >= best) continue;        // something nearer was found since this node was pushed
        const int fx = int(f.x & 0xffffu), fy = int(f.x >> 16), fz = int(f.y & 0xffffu), level = int(f.y >> 16);
        const uint2 w = __ldg(slots + f.z);
        const uint32_t child_mask = (w.x >> 8) & 0xffu, leaf_mask = (w.x >> 16) & 0xffu;
        const int half = 1 << (level - 1);
        // front to back: lane 0 takes the octant on the camera's side, lane 7 the opposite one
        const uint32_t near_upper = (cx >= float(fx + half) ? 1u : 0u) | (cy >= float(fy + half) ? 2u : 0u) | (cz >= float(fz + half) ? 4u : 0u);
        const uint32_t upper = uint32_t(sub) ^ near_upper;
        const uint32_t s = ~upper & 7u;                    // slot bit clear = upper half (the tree stores the mirrored octant, lsvo.hpp:79)
        const int bx = fx + ((upper & 1u) ? half : 0), by = fy + ((upper & 2u) ? half : 0), bz = fz + ((upper & 4u) ? half : 0);
        float d_min = 3.0e9f;
        bool search = false, terminal = false;
        if ((child_mask >> s) & 1u) {
            const float x0 = float(bx) - cx, y0 = float(by) - cy, z0 = float(bz) - cz, e = float(half);
            // distance from the camera to the box (0 inside) and to its far corner
            const float nxd = fmaxf(fmaxf(x0, -(x0 + e)), 0.0f), nyd = fmaxf(fmaxf(y0, -(y0 + e)), 0.0f), nzd = fmaxf(fmaxf(z0, -(z0 + e)), 0.0f);
            d_min = sqrtf(nxd * nxd + nyd * nyd + nzd * nzd);
            if (d_min < best) {
                const float ax = fmaxf(fabsf(x0), fabsf(x0 + e)), ay = fmaxf(fabsf(y0), fabsf(y0 + e)), az = fmaxf(fabsf(z0), fabsf(z0 + e));
                const float d_far = sqrtf(ax * ax + ay * ay + az * az);
                const float slack = lens + lens_growth * d_far + 1e-3f * e + 1e-4f * d_far;
                bool outside = false;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    // the box corner farthest along the inward normal
                    const float qx = pn[k][0] >= 0.0f ? x0 + e : x0, qy = pn[k][1] >= 0.0f ? y0 + e : y0, qz = pn[k][2] >= 0.0f ? z0 + e : z0;
                    if (pn[k][0] * qx + pn[k][1] * qy + pn[k][2] * qz < -slack) outside = true;
                }
                if (!outside) {
                    const bool leaf = (leaf_mask >> s) & 1u;
                    terminal = leaf || level - 1 == 0 || e <= 0.5f * spread * d_min;
                    search = !terminal;
                }
            }
        }
        // the nearest terminal child bounds everything else
        float cand = terminal ? d_min : 3.0e9f;
        cand = fminf(cand, __shfl_xor_sync(gmask, cand, 1));
        cand = fminf(cand, __shfl_xor_sync(gmask, cand, 2));
        cand = fminf(cand, __shfl_xor_sync(gmask, cand, 4));
        best = fminf(best, cand);
        search = search && d_min < best;
        const unsigned pushing = (__ballot_sync(gmask, search) >> (lane & 24)) & 0xffu;
        const int n_push = __popc(pushing);
        if (sp + n_push > kBeamStack) {                    // cannot happen for depth <= 12: stay conservative
            float m = search ? d_min : 3.0e9f;
            m = fminf(m, __shfl_xor_sync(gmask, m, 1));
            m = fminf(m, __shfl_xor_sync(gmask, m, 2));
            m = fminf(m, __shfl_xor_sync(gmask, m, 4));
            best = fminf(best, m);
        } else {
            // farthest first: lane `sub` goes below every pushing lane nearer than itself
            if (search) stack[sp + __popc(pushing >> (sub + 1))] = make_uint4(uint32_t(bx) | uint32_t(by) << 16, uint32_t(bz) | uint32_t(level - 1) << 16,
                                                                             f.z + w.y + s, __float_as_uint(d_min));
            sp += n_push;
        }
        __syncwarp(gmask);
    }
    if (sub == 0) {
        const float t = (best - lens - 2.0f) / S;          // castRay's t is distance in the unit cube's units; two voxels of slack
        floor[id] = best > 2.9e9f ? 3.0f : fmaxf(0.0f, t - 1e-5f * fabsf(t));   // empty frustum: beyond the cube's diagonal
    }
}

cudaError_t launch_beam_floor(const uint2* nodes, const RenderLaunch& L, int tile, float* d_floor, cudaStream_t stream) {
    const int tiles_x = (L.width + tile - 1) / tile;
    const int ty0 = L.row_begin / tile, ty1 = (L.row_end + tile - 1) / tile;
    const int n = tiles_x * (ty1 - ty0);
    if (n <= 0) return cudaSuccess;
    const int per_block = kBeamThreads / 8;
    beam_floor_kernel<<<(n + per_block - 1) / per_block, kBeamThreads, 0, stream>>>(nodes, L, tile, tiles_x, ty1 - ty0, ty0, d_floor + size_t(ty0) * tiles_x);
    return cudaGetLastError();
}

}  // namespace vrt

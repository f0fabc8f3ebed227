// LSVO stack traversal on the device — the engine's restatement of LSVO<D>::castRay
// (reference include/lsvo.hpp:33-172), op for op in fp32 without contraction so that the hit record
// and the iteration count (HitPoint::complexity) are bit-identical to the CPU reference.
//
// The traversal state lives in registers; the per-ray stack (lsvo.hpp:42, OctreeStack
// lsvo_utils.hpp:35-39) is addressed through a policy so that kernels can keep it in shared memory.
#pragma once
#include "vrt_device.cuh"

namespace vrt {

// One octree node as the traversal sees it, whatever the memory format.
struct NodeView {
    uint32_t raw;          // first word of the slot: color | child_mask << 8 | leaf_mask << 16
    uint32_t child_mask;   // LNode::child_mask
    uint32_t leaf_mask;    // LNode::leaf_mask
    uint32_t child_base;   // index such that child slot s lives at child_base + s
};

// Reference layout (lsvo_utils.hpp:5-18): 8-byte slots, children at parent + child_offset + slot.
struct RefNodes {
    const uint2* __restrict__ slots;
    __device__ __forceinline__ NodeView fetch(uint32_t id) const {
        const uint2 w = __ldg(slots + id);                 // one coalescable 8-byte load (lsvo.hpp:74)
        NodeView v;
        v.raw = w.x;
        v.child_mask = (w.x >> 8) & 0xffu;
        v.leaf_mask = (w.x >> 16) & 0xffu;
        v.child_base = id + w.y;
        return v;
    }
    __device__ __forceinline__ uint32_t child(const NodeView& v, uint32_t slot) const { return v.child_base + slot; }
};

// Per-thread stack in local memory (v0 policy).
struct LocalStack {
    uint32_t parent[kSvoMaxDepth + 1];
    float t_max[kSvoMaxDepth + 1];
    __device__ __forceinline__ void push(int i, uint32_t p, float t) { parent[i] = p; t_max[i] = t; }
    __device__ __forceinline__ void pop(int i, uint32_t& p, float& t) const { p = parent[i]; t = t_max[i]; }
};

// Stack in shared memory, one column per thread: entry i of thread t at [i * blockDim + t] (conflict-free).
struct SharedStack {
    uint32_t* parent;   // already offset by threadIdx.x
    float* t_max;
    int stride;
    __device__ __forceinline__ void push(int i, uint32_t p, float t) { parent[i * stride] = p; t_max[i * stride] = t; }
    __device__ __forceinline__ void pop(int i, uint32_t& p, float& t) const { p = parent[i * stride]; t = t_max[i * stride]; }
};

struct LsvoResult {
    float px, py, pz;      // cell low corner in the mirrored frame (un-mirrored by finish())
    float t_min;
    float scale_f;
    int scale;
    uint32_t face;         // step mask that entered the cell (`normal` in lsvo.hpp:69,122)
    uint32_t mirror;
    uint32_t complexity;
    bool hit;
    float dx, dy, dz;      // direction after the |d| >= 2^-23 clamp (lsvo.hpp:44-46)
};

template <typename Nodes, typename Stack>
__device__ __forceinline__ void lsvo_cast(const Nodes& nodes, Stack& stack, int depth, int guard, float ox, float oy,
                                          float oz, float dx, float dy, float dz, float coef, float bias, LsvoResult& r) {
    const int depth_offset = kSvoMaxDepth - depth;                       // lsvo.hpp:38
    if (fabsf(dx) < kEps) dx = copysignf(kEps, dx);                      // lsvo.hpp:44-46
    if (fabsf(dy) < kEps) dy = copysignf(kEps, dy);
    if (fabsf(dz) < kEps) dz = copysignf(kEps, dz);
    const float tcx = -1.0f / fabsf(dx), tcy = -1.0f / fabsf(dy), tcz = -1.0f / fabsf(dz);   // :47
    float tox = ox * tcx, toy = oy * tcy, toz = oz * tcz;                // :48
    uint32_t mirror = 7u;
    if (dx > 0.0f) { mirror ^= 1u; tox = 3.0f * tcx - tox; }             // :50-52
    if (dy > 0.0f) { mirror ^= 2u; toy = 3.0f * tcy - toy; }
    if (dz > 0.0f) { mirror ^= 4u; toz = 3.0f * tcz - toz; }
    float t_min = fmaxf(2.0f * tcx - tox, fmaxf(2.0f * tcy - toy, 2.0f * tcz - toz));   // :54
    float t_max = fminf(tcx - tox, fminf(tcy - toy, tcz - toz));                        // :55
    float h = t_max;
    t_min = fmaxf(0.0f, t_min);
    t_max = fminf(1.0f, t_max);
    uint32_t parent = 0u, child = 0u, face = 0u;
    int scale = kSvoMaxDepth - 1;
    float px = 1.0f, py = 1.0f, pz = 1.0f, scale_f = 0.5f;
    if (1.5f * tcx - tox > t_min) { child ^= 1u; px = 1.5f; }           // :66-68
    if (1.5f * tcy - toy > t_min) { child ^= 2u; py = 1.5f; }
    if (1.5f * tcz - toz > t_min) { child ^= 4u; pz = 1.5f; }
    bool hit = false;
    uint32_t iters = 0u;

    while (scale < kSvoMaxDepth && scale > guard) {                      // :72
        ++iters;
        const NodeView nd = nodes.fetch(parent);                         // :74
        const float cx = px * tcx - tox, cy = py * tcy - toy, cz = pz * tcz - toz;   // :76
        const float tc_max = fminf(cx, fminf(cy, cz));
        const uint32_t shift = child ^ mirror;                           // :79
        if (((nd.child_mask >> shift) & 1u) && t_min <= t_max) {         // :80-81
            if (tc_max * coef + bias >= scale_f) { hit = true; break; }  // :82-85
            const float tv_max = fminf(t_max, tc_max);
            const float half = scale_f * 0.5f;
            if (t_min <= tv_max) {                                       // :89
                if ((nd.leaf_mask >> shift) & 1u) { hit = true; break; } // :90-95
                if (tc_max < h) stack.push(scale - depth_offset, parent, t_max);   // :97-100
                h = tc_max;
                parent = nodes.child(nd, shift);                         // :103
                child = 0u;
                --scale;
                scale_f = half;
                if (half * tcx + cx > t_min) { child ^= 1u; px += scale_f; }   // :88,107-109
                if (half * tcy + cy > t_min) { child ^= 2u; py += scale_f; }
                if (half * tcz + cz > t_min) { child ^= 4u; pz += scale_f; }
                t_max = tv_max;
                continue;
            }
        }
        uint32_t step = 0u;                                              // :115-118
        if (cx <= tc_max) { step ^= 1u; px -= scale_f; }
        if (cy <= tc_max) { step ^= 2u; py -= scale_f; }
        if (cz <= tc_max) { step ^= 4u; pz -= scale_f; }
        t_min = tc_max;
        child ^= step;
        face = step;
        if (child & step) {                                              // :124-145
            const uint32_t ix = __float_as_uint(px), iy = __float_as_uint(py), iz = __float_as_uint(pz);
            uint32_t diff = 0u;
            if (step & 1u) diff |= ix ^ __float_as_uint(px + scale_f);
            if (step & 2u) diff |= iy ^ __float_as_uint(py + scale_f);
            if (step & 4u) diff |= iz ^ __float_as_uint(pz + scale_f);
            scale = int((__float_as_uint(__uint2float_rn(diff)) >> 23) - 127u);   // :132
            scale_f = __uint_as_float(uint32_t(scale - kSvoMaxDepth + 127) << 23);  // :133
            if (scale >= kSvoMaxDepth) break;   // left the root cube; the reference's stack read here is dead
            stack.pop(scale - depth_offset, parent, t_max);              // :134-136
            const uint32_t sx = ix >> scale, sy = iy >> scale, sz = iz >> scale;
            px = __uint_as_float(sx << scale);
            py = __uint_as_float(sy << scale);
            pz = __uint_as_float(sz << scale);
            child = (sx & 1u) | ((sy & 1u) << 1) | ((sz & 1u) << 2);
            h = 0.0f;
        }
    }
    r.px = px; r.py = py; r.pz = pz;
    r.t_min = t_min; r.scale_f = scale_f; r.scale = scale; r.face = face; r.mirror = mirror;
    r.complexity = iters; r.hit = hit;
    r.dx = dx; r.dy = dy; r.dz = dz;
}

// Hit epilogue, lsvo.hpp:148-169.  Only valid when r.hit.
struct LsvoHit {
    float pos[3];
    float normal[3];
    float uv[2];
    float distance;
    float corner[3];   // un-mirrored low corner of the hit cell
};

__device__ __forceinline__ float glm_sign(float x) { return float(0.0f < x) - float(x < 0.0f); }

__device__ __forceinline__ void lsvo_finish(const LsvoResult& r, float ox, float oy, float oz, int depth, LsvoHit& h) {
    h.normal[0] = (-glm_sign(r.dx)) * float(r.face & 1u);                // :149
    h.normal[1] = (-glm_sign(r.dy)) * float(r.face & 2u);
    h.normal[2] = (-glm_sign(r.dz)) * float(r.face & 4u);
    float px = r.px, py = r.py, pz = r.pz;
    if ((r.mirror & 1u) == 0u) px = 3.0f - r.scale_f - px;               // :151-153
    if ((r.mirror & 2u) == 0u) py = 3.0f - r.scale_f - py;
    if ((r.mirror & 4u) == 0u) pz = 3.0f - r.scale_f - pz;
    h.corner[0] = px; h.corner[1] = py; h.corner[2] = pz;
    h.distance = r.t_min;                                                // :155
    h.pos[0] = fminf(fmaxf(ox + r.t_min * r.dx, px + kEps), px + r.scale_f - kEps);   // :156-158
    h.pos[1] = fminf(fmaxf(oy + r.t_min * r.dy, py + kEps), py + r.scale_f - kEps);
    h.pos[2] = fminf(fmaxf(oz + r.t_min * r.dz, pz + kEps), pz + r.scale_f - kEps);
    const float S = float(1 << depth);                                   // :39
    if (h.normal[0] != 0.0f) { h.uv[0] = fracf(h.pos[2] * S); h.uv[1] = fracf(h.pos[1] * S); }        // :160-168
    else if (h.normal[1] != 0.0f) { h.uv[0] = fracf(h.pos[0] * S); h.uv[1] = fracf(h.pos[2] * S); }
    else if (h.normal[2] != 0.0f) { h.uv[0] = fracf(h.pos[0] * S); h.uv[1] = fracf(h.pos[1] * S); }
    else { h.uv[0] = 0.0f; h.uv[1] = 0.0f; }   // uninitialised in the reference (start inside a solid cell)
}

}  // namespace vrt

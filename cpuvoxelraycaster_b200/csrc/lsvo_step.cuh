// LSVO traversal as a resumable per-lane state machine: init() = prologue of LSVO<D>::castRay
// (reference include/lsvo.hpp:44-70), step() = exactly one trip of its while loop (:72-146),
// lsvo_finish() = the hit epilogue (:148-169, in lsvo_traverse.cuh).
//
// Same fp32 operations in the same order as the reference, so results are bit-identical.  What is NOT
// copied is the integer bookkeeping around them, which the first ncu source-level capture showed to be
// a quarter of all issued instructions (profiles/r01_summary.md):
//   * POP (lsvo.hpp:124-145): the reference finds the highest differing bit by converting the XOR of two
//     float bit patterns to float and reading its exponent (:132).  The operand always has < 24
//     significant bits (positions are multiples of 2^-12 in [0.5, 2], so its low 10 bits are zero), the
//     conversion is exact and the exponent is the index of the highest set bit: one FLO (31 - clz).
//     Truncating the position to the new scale ((x >> s) << s, :137-142) is one AND with a mask.
//   * PUSH/POP use ONE 8-byte shared-memory access per stack entry ({parent, t_max} interleaved, entry-major,
//     compile-time stride), instead of two 4-byte accesses with run-time address arithmetic.
//   * the node base pointer and the loop guard are pinned in registers (they were re-read from the constant
//     bank on every trip, in the dependent chain in front of the node fetch).
#pragma once
#include "kernels.h"
#include "lsvo_traverse.cuh"

namespace vrt {

// Stack in shared memory: entry i of thread t at base[i * kThreads + t] (uint2 = {parent index, t_max bits}).
template <int kThreads>
struct Stack64 {
    uint2* base;   // already offset by the thread index
    __device__ __forceinline__ void push(int i, uint32_t p, float t) { base[i * kThreads] = make_uint2(p, __float_as_uint(t)); }
    __device__ __forceinline__ void pop(int i, uint32_t& p, float& t) const {
        const uint2 e = base[i * kThreads];
        p = e.x;
        t = __uint_as_float(e.y);
    }
};

// Keeps a kernel parameter in a register across the traversal loop.  ptxas re-materialises parameters from the
// constant bank (an LDC in front of every node fetch); adding blockIdx.y — always 0, every launch is 1-D, but
// not provably so — makes the value a computed one that has to stay in a register.
__device__ __forceinline__ const uint2* pin(const uint2* p) {
    // threadIdx.z (always 0: every launch is 1-D) rather than a block index: the sum is then a per-thread value in a vector
    // register pair and the node address is ONE IMAD.WIDE with the immediate 8 (with the pointer in a uniform register ptxas
    // has to materialise the 8 in a register first, every trip)
    return reinterpret_cast<const uint2*>(reinterpret_cast<uintptr_t>(p) + blockIdx.y + threadIdx.z);
}
__device__ __forceinline__ int pin(int v) { return v + int(blockIdx.y); }
// A value ptxas cannot re-derive: written to and read back from the thread's own shared-memory slot (volatile), once per
// kernel.  For loop invariants that must live in a register across the traversal loop — ptxas re-materialises anything it can
// trace to kernel parameters or special registers, every trip, when registers are scarce.
__device__ __forceinline__ uint32_t keep_in_register(uint32_t v, const void* own_shared_slot) {
    asm volatile("st.volatile.shared.b32 [%1], %0;\n\tld.volatile.shared.b32 %0, [%1];" : "+r"(v) : "r"(uint32_t(__cvta_generic_to_shared(own_shared_slot))) : "memory");
    return v;
}
__device__ __forceinline__ int keep_in_register(int v, const void* slot) { return int(keep_in_register(uint32_t(v), slot)); }
__device__ __forceinline__ float keep_in_register(float v, const void* slot) { return __uint_as_float(keep_in_register(__float_as_uint(v), slot)); }

struct Trav {
    // ray (direction after the |d| >= 2^-23 clamp) and cone
    float ox, oy, oz, dx, dy, dz, coef, bias;
    // traversal state
    float tcx, tcy, tcz, tox, toy, toz;
    float px, py, pz;
    float t_min, t_max, h;
    uint32_t parent;
    uint32_t child, mirror, face;
    int scale;
    uint32_t iters;
    bool hit;

    __device__ __forceinline__ float scale_f() const { return __uint_as_float(uint32_t(scale - kSvoMaxDepth + 127) << 23); }

    // A ray with a non-finite origin or direction never leaves the reference's loop (every comparison with NaN fails: no
    // descent, no step, no pop); here it is a miss of complexity 0.  To keep the loop itself untouched such a ray is
    // replaced by one that starts outside the cube pointing away (one trip, no descent) and is marked in bit 3 of
    // `mirror`, which result() reads.
    __device__ __forceinline__ void init(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_, float coef_, float bias_) {
        // inf * 0 and NaN * 0 are NaN; per component, so that large finite components cannot overflow a sum
        const bool finite = (((ox_ * 0.0f + oy_ * 0.0f) + oz_ * 0.0f) + ((dx_ * 0.0f + dy_ * 0.0f) + dz_ * 0.0f)) == 0.0f;
        if (!finite) { ox_ = 3.0f; oy_ = 3.0f; oz_ = 3.0f; dx_ = 1.0f; dy_ = 1.0f; dz_ = 1.0f; }
        ox = ox_; oy = oy_; oz = oz_; coef = coef_; bias = bias_;
        if (fabsf(dx_) < kEps) dx_ = copysignf(kEps, dx_);                 // lsvo.hpp:44-46
        if (fabsf(dy_) < kEps) dy_ = copysignf(kEps, dy_);
        if (fabsf(dz_) < kEps) dz_ = copysignf(kEps, dz_);
        dx = dx_; dy = dy_; dz = dz_;
        tcx = -1.0f / fabsf(dx); tcy = -1.0f / fabsf(dy); tcz = -1.0f / fabsf(dz);   // :47
        tox = ox * tcx; toy = oy * tcy; toz = oz * tcz;                     // :48
        mirror = finite ? 7u : 15u;
        if (dx > 0.0f) { mirror ^= 1u; tox = 3.0f * tcx - tox; }           // :50-52
        if (dy > 0.0f) { mirror ^= 2u; toy = 3.0f * tcy - toy; }
        if (dz > 0.0f) { mirror ^= 4u; toz = 3.0f * tcz - toz; }
        t_min = fmaxf(2.0f * tcx - tox, fmaxf(2.0f * tcy - toy, 2.0f * tcz - toz));   // :54
        t_max = fminf(tcx - tox, fminf(tcy - toy, tcz - toz));                         // :55
        h = t_max;
        t_min = fmaxf(0.0f, t_min);
        t_max = fminf(1.0f, t_max);
        parent = 0u; child = 0u; face = 0u;
        scale = kSvoMaxDepth - 1;
        px = 1.0f; py = 1.0f; pz = 1.0f;
        if (1.5f * tcx - tox > t_min) { child ^= 1u; px = 1.5f; }          // :66-68
        if (1.5f * tcy - toy > t_min) { child ^= 2u; py = 1.5f; }
        if (1.5f * tcz - toz > t_min) { child ^= 4u; pz = 1.5f; }
        iters = 0u;
        hit = false;
    }

    // One trip of the loop.  Returns true while the ray is alive (the loop condition :72 still holds and no hit).
    template <typename Nodes, typename Stack>
    __device__ __forceinline__ bool step(const Nodes& nodes, Stack& stack, int depth_offset, int guard) {
        ++iters;
        const float sf = scale_f();
        const NodeView nd = nodes.fetch(parent);                             // :74
        const float cx = px * tcx - tox, cy = py * tcy - toy, cz = pz * tcz - toz;   // :76
        const float tc_max = fminf(cx, fminf(cy, cz));
        const uint32_t shift = child ^ mirror;                               // :79
        const uint32_t child_bit = 0x100u << shift;                          // child_mask lives in bits 8..15 of the word
        if ((nd.raw & child_bit) && t_min <= t_max) {                        // :80-81
            if (tc_max * coef + bias >= sf) { hit = true; return false; }    // :82-85
            const float tv_max = fminf(t_max, tc_max);
            const float half = sf * 0.5f;
            if (t_min <= tv_max) {                                           // :89
                if (nd.raw & (child_bit << 8)) { hit = true; return false; }  // leaf_mask: bits 16..23, :90-95
                if (tc_max < h) stack.push(scale - depth_offset, parent, t_max);  // :97-100
                h = tc_max;
                parent = nodes.child(nd, shift);                             // :103
                child = 0u;
                --scale;
                if (half * tcx + cx > t_min) { child ^= 1u; px += half; }    // :88,107-109
                if (half * tcy + cy > t_min) { child ^= 2u; py += half; }
                if (half * tcz + cz > t_min) { child ^= 4u; pz += half; }
                t_max = tv_max;
                return scale > guard;                                        // :72 (scale < 23 holds after a descent)
            }
        }
        const uint32_t ox_bits = __float_as_uint(px), oy_bits = __float_as_uint(py), oz_bits = __float_as_uint(pz);
        uint32_t step_mask = 0u;                                             // :115-118
        if (cx <= tc_max) { step_mask ^= 1u; px -= sf; }
        if (cy <= tc_max) { step_mask ^= 2u; py -= sf; }
        if (cz <= tc_max) { step_mask ^= 4u; pz -= sf; }
        t_min = tc_max;
        child ^= step_mask;
        face = step_mask;
        if (child & step_mask) {                                             // :124-145
            const uint32_t ix = __float_as_uint(px), iy = __float_as_uint(py), iz = __float_as_uint(pz);
            // :126-131: on a stepped axis p + scale_f is exactly the position before the step (grid-aligned
            // values, exact subtraction and addition); on the other axes the position did not change, XOR = 0
            const uint32_t diff = (ix ^ ox_bits) | (iy ^ oy_bits) | (iz ^ oz_bits);
            scale = 31 - __clz(int(diff));        // == (floatAsInt((float)diff) >> 23) - 127 (:132), see the header
            if (scale >= kSvoMaxDepth) return false;                         // left the root cube: miss
            stack.pop(scale - depth_offset, parent, t_max);                  // :134-136
            const uint32_t keep = 0xffffffffu << scale;                      // (x >> s) << s == x & ~((1 << s) - 1)
            px = __uint_as_float(ix & keep);                                 // :137-142
            py = __uint_as_float(iy & keep);
            pz = __uint_as_float(iz & keep);
            child = ((ix >> scale) & 1u) | (((iy >> scale) & 1u) << 1) | (((iz >> scale) & 1u) << 2);   // :143
            h = 0.0f;
            return scale > guard;
        }
        return true;
    }

    __device__ __forceinline__ void result(LsvoResult& r) const {
        r.px = px; r.py = py; r.pz = pz;
        r.t_min = t_min; r.scale_f = scale_f(); r.scale = scale; r.face = face; r.mirror = mirror & 7u;
        r.complexity = (mirror & 8u) ? 0u : iters; r.hit = hit;
        r.dx = dx; r.dy = dy; r.dz = dz;
    }
};

// Whole ray: LSVO<D>::castRay (lsvo.hpp:33-147).  The hit epilogue is lsvo_finish().
template <typename Nodes, typename Stack>
__device__ __forceinline__ void lsvo_cast_ray(const Nodes& nodes, Stack& stack, int depth_offset, int guard, float ox, float oy,
                                              float oz, float dx, float dy, float dz, float coef, float bias, LsvoResult& r) {
    Trav t;
    t.init(ox, oy, oz, dx, dy, dz, coef, bias);
    while (t.step(nodes, stack, depth_offset, guard)) {}
    t.result(r);
}

// ---- Trav2: the same loop with its bookkeeping moved off the ALU pipe --------------------------------------------------
// ncu (profiles/r01, r02): the frame kernels issue at 75 % with the half-rate ALU pipe (integer, compare, min/max, select,
// logic) 75 % busy and the full-rate FMA pipe 22 % busy — half of all instructions of the loop are ALU-pipe instructions.
// Trav2 computes exactly the reference's fp32 operations (results bit-identical) but
//   * keeps the cell size sf = 2^(scale-23) as a float that is HALVED on a descent (an FMUL that the loop needs anyway for
//     `half`) instead of rebuilding it from an integer scale every trip (LEA) and decrementing that integer (IADD3); the
//     integer scale only exists inside POP, where FLO produces it;
//   * counts loop trips with an FADD (exact below 2^24 trips) instead of an IADD3;
//   * drops the cone test `tc_max * coef + bias >= sf` (lsvo.hpp:82-85) at compile time for rays cast with coef = bias = 0
//     (primary and sun-shadow rays: the test can never be true there — 0 >= sf is false for every sf > 0, and a NaN product,
//     inf * 0, compares false as well);
//   * addresses the stack from the bits of sf (push) / with one IMAD from the scale (pop);
// and, after the per-instruction view of the 8-CTA frame kernel (profiles/r02_summary.md: most warp-trips execute all three
// paths of the loop, so an instruction removed from any path is removed from most trips),
//   * applies the ADVANCE step once for both continuations (POP rebuilds the old position as p + sf, exactly);
//   * has one exit for both kinds of hit and no hit flag (how the walk ended is read off its final state, Trav2::hit);
//   * compiles the loop guard out where it cannot bind (kGuard) and uses one FFMA per axis for the child selection where the
//     product is exact (kUnit);
//   * keeps the current node's two words in registers across trips (loaded when `parent` changes only);
//   * pushes unconditionally (no `h`: the same pairs are read back, see step()).
// 107 -> 86 instructions per loop in the frame kernels; every step with byte-identical frames and trip counts.
template <int kThreads>
struct Stack64s {
    uint32_t addr;       // shared-memory byte address of the entry of scale 0 (entry of scale s at addr + s * kThreads * 8)
    // `thread_base` is the thread's first stack slot.  The address is passed through a volatile shared-memory round trip: ptxas
    // otherwise re-derives it from %tid / %ctaid and the kernel parameters on every push and pop (11 of the 14 instructions of a
    // push in the round-2 capture) instead of keeping one register alive across the loop.
    __device__ __forceinline__ static Stack64s make(uint2* thread_base, int depth_offset) {
        Stack64s st;
        uint32_t a = uint32_t(__cvta_generic_to_shared(thread_base)) - uint32_t(depth_offset * kThreads * 8);
        asm volatile("st.volatile.shared.b32 [%1], %0;\n\tld.volatile.shared.b32 %0, [%1];" : "+r"(a) : "r"(uint32_t(__cvta_generic_to_shared(thread_base))) : "memory");
        st.addr = a;
        return st;
    }
    // sf = 2^(scale - 23): bits = (scale + 104) << 23, so byte offset (scale + 104) * 128 * 8 = bits >> 13 (the low 23 bits of a
    // power of two are zero: no masking needed)
    __device__ __forceinline__ void push_sf(float sf, uint32_t p, float t) {
        static_assert(kThreads == 128, "the shift below assumes 128 threads per block");
        const uint32_t a = addr + (__float_as_uint(sf) >> 13);
        asm volatile("st.shared.v2.b32 [%0 + -106496], {%1, %2};" :: "r"(a), "r"(p), "r"(__float_as_uint(t)) : "memory");
    }
    __device__ __forceinline__ void pop(int scale, uint32_t& p, float& t) const {
        uint32_t ty;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(p), "=r"(ty) : "r"(addr + uint32_t(scale) * uint32_t(kThreads * 8)) : "memory");
        t = __uint_as_float(ty);
    }
    // Per-thread statistics pairs behind the stack, addressed from the same register: the stack's slots are the scales
    // 23 - depth .. 23, so whatever the depth the 8-byte slot of "scale" 24 + k is the k-th pair behind the thread's stack
    // column (the frame kernels reserve six: rays and loop trips per ray class).  One LDS.64 / STS.64 and an IMAD per ray; the
    // plain array form re-derived the address from %tid and the kernel parameters every time (20 instructions per ray).
    __device__ __forceinline__ void stat_zero(int k) const {
        asm volatile("st.shared.v2.b32 [%0], {%1, %1};" :: "r"(addr + uint32_t(kSvoMaxDepth + 1 + k) * uint32_t(kThreads * 8)), "r"(0u) : "memory");
    }
    __device__ __forceinline__ void stat_add(int k, uint32_t a, uint32_t b) const {
        const uint32_t at = addr + uint32_t(kSvoMaxDepth + 1 + k) * uint32_t(kThreads * 8);
        uint32_t x, y;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(at) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(at), "r"(x + a), "r"(y + b) : "memory");
    }
    __device__ __forceinline__ void stat_get(int k, uint32_t& a, uint32_t& b) const {
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr + uint32_t(kSvoMaxDepth + 1 + k) * uint32_t(kThreads * 8)) : "memory");
    }
};

// Bounds of everything solid in the scene (castRay's [1,2]^3 coordinates, a little enlarged): a ray that has left this box
// cannot hit anything any more.  kBounds walks stop there instead of at the root cube's far side (lsvo.hpp:72) — a shadow ray
// that has climbed out of the terrain no longer pays for the empty half of the world.  Misses stay misses; only their trip
// count shrinks, so this is for frames (like the beam floors), not for the batched casts that report HitPoint::complexity.
// Only for rays without a cone (coef = bias = 0: camera and sun-shadow rays): a cone ray "hits" the first NON-EMPTY NODE it
// finds smaller than its cone (lsvo.hpp:82-85), and such a node reaches far beyond the voxels it contains — measured: with the
// bound applied to the GI rays 0.4 % of them lose their hit.
// kGuard = false compiles the loop guard `scale > MAX_DEPTH` (lsvo.hpp:72) out: with guard < 23 - depth it can never bind — the
// deepest cells a walk can stand in are the voxels (scale 23 - depth: every child of a node of the level above is a leaf or
// empty, which vrt_lsvo_create validates and the device builders guarantee), the test after a descent compares a scale
// >= 23 - depth, and the one after a POP a scale larger than one that has passed already.  The launchers pick the variant.
// kUnit = true promises |d| <= 2^100 or so (the frame kernels: their directions are normalised, or a normalised vector times the
// camera's rotation matrix, which vrt_render* checks): the child selection of a descent may then use one FFMA per axis for
// half * tc + c.  That is the reference's FMUL + FADD bit for bit — half is a power of two and |tc| = 1 / |d| is far from the
// denormals, so the product is exact and the only rounding is the sum's, as in the two-instruction form.
template <bool kCone, bool kBounds = false, bool kGuard = true, bool kUnit = false>
struct Trav2 {
    float dx, dy, dz, coef, bias;
    float tcx, tcy, tcz, tox, toy, toz;
    float px, py, pz;
    float t_min, t_max;
    float sf, iters_f;
    float t_limit;       // kBounds: the time at which the ray leaves the scene bounds
    uint32_t parent, child, mirror, face;
    uint2 nd_;           // the two words of node `parent`, loaded when parent changes (a descent, a POP) instead of on every trip
    // There is no `hit` flag: a flag written inside the loop costs an instruction per trip on two of its three paths (ptxas sets
    // it in front of the exits).  How the walk ended is read off the state it ended in, see hit().

    // t_floor: the walk starts at max(t_floor, entry into the root cube) instead of max(0, ...) (:57).  Any t_floor below the
    // ray's hit distance leaves the HitPoint unchanged except for its complexity (beam_kernels.cu); 0 = the reference.
    __device__ __forceinline__ void init(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_, float coef_, float bias_,
                                         float t_floor = 0.0f, const SceneBounds* bounds = nullptr) {
        const bool finite = (((ox_ * 0.0f + oy_ * 0.0f) + oz_ * 0.0f) + ((dx_ * 0.0f + dy_ * 0.0f) + dz_ * 0.0f)) == 0.0f;
        if (!finite) { ox_ = 3.0f; oy_ = 3.0f; oz_ = 3.0f; dx_ = 1.0f; dy_ = 1.0f; dz_ = 1.0f; }
        coef = coef_; bias = bias_;
        if (fabsf(dx_) < kEps) dx_ = copysignf(kEps, dx_);                 // lsvo.hpp:44-46
        if (fabsf(dy_) < kEps) dy_ = copysignf(kEps, dy_);
        if (fabsf(dz_) < kEps) dz_ = copysignf(kEps, dz_);
        dx = dx_; dy = dy_; dz = dz_;
        // :47 — -1/|d| correctly rounded == -(correctly rounded 1/|d|): the reciprocal intrinsic is the shorter sequence
        tcx = -__frcp_rn(fabsf(dx)); tcy = -__frcp_rn(fabsf(dy)); tcz = -__frcp_rn(fabsf(dz));
        tox = ox_ * tcx; toy = oy_ * tcy; toz = oz_ * tcz;                  // :48
        mirror = finite ? 7u : 15u;
        if (dx > 0.0f) { mirror ^= 1u; tox = 3.0f * tcx - tox; }           // :50-52
        if (dy > 0.0f) { mirror ^= 2u; toy = 3.0f * tcy - toy; }
        if (dz > 0.0f) { mirror ^= 4u; toz = 3.0f * tcz - toz; }
        t_min = fmaxf(2.0f * tcx - tox, fmaxf(2.0f * tcy - toy, 2.0f * tcz - toz));   // :54
        t_max = fminf(tcx - tox, fminf(tcy - toy, tcz - toz));                         // :55 (h = t_max, :56: see the push in step())
        t_min = fmaxf(t_floor, t_min);                                     // :57 with t_floor = 0
        t_max = fminf(1.0f, t_max);
        if (kBounds) {
            // exit from the bounds box, with the expressions of :55 (in the mirrored frame every direction is negative, so the
            // exit planes are the low ones: lo on an unmirrored axis, 3 - hi on a mirrored one)
            const float bx = (mirror & 1u) ? bounds->lo[0] : 3.0f - bounds->hi[0];
            const float by = (mirror & 2u) ? bounds->lo[1] : 3.0f - bounds->hi[1];
            const float bz = (mirror & 4u) ? bounds->lo[2] : 3.0f - bounds->hi[2];
            t_limit = fminf(bx * tcx - tox, fminf(by * tcy - toy, bz * tcz - toz));
        }
        parent = 0u; child = 0u; face = 0u;
        sf = 0.5f;                                                         // scale = 22
        px = 1.0f; py = 1.0f; pz = 1.0f;
        if (1.5f * tcx - tox > t_min) { child ^= 1u; px = 1.5f; }          // :66-68
        if (1.5f * tcy - toy > t_min) { child ^= 2u; py = 1.5f; }
        if (1.5f * tcz - toz > t_min) { child ^= 4u; pz = 1.5f; }
        iters_f = 0.0f;
    }
    // the root node, before the first trip
    template <typename Nodes> __device__ __forceinline__ void prime(const Nodes& nodes) { nd_ = nodes.load(0u); }

    // one trip of the loop (:72-146); guard_sf = 2^(guard - 23)
    template <typename Nodes, typename Stack>
    __device__ __forceinline__ bool step(const Nodes& nodes, Stack& stack, int guard, float guard_sf) {
        iters_f += 1.0f;
        const uint32_t nd_raw = nd_.x;                                       // :74
        const float cx = px * tcx - tox, cy = py * tcy - toy, cz = pz * tcz - toz;   // :76
        const float tc_max = fminf(cx, fminf(cy, cz));
        const uint32_t shift = child ^ mirror;                               // :79
        const uint32_t masks = nd_raw >> shift;                              // bit 8: the child exists, bit 16: it is a leaf
        if ((masks & 0x100u) && t_min <= t_max) {                            // :80-81
            const float tv_max = fminf(t_max, tc_max);
            const bool inside = t_min <= tv_max;                             // :89
            // the two ways of ending on a hit — the cone is wider than the cell (:82-85, tested first in the reference), the child
            // is a leaf (:90-95) — leave the same state behind: one exit for both
            // (as a mask, not as a bool: ptxas turns the boolean form into a shift, an AND and an integer compare)
            uint32_t ends = masks & (inside ? 0x10000u : 0u);
            if (kCone) { if (tc_max * coef + bias >= sf) ends = 1u; }
            if (ends) return false;
            if (inside) {
                const float half = sf * 0.5f;
                // :97-100 push unconditionally.  The reference skips the write when the child's exit time equals `h` (the exit time
                // of the cell it is in, or 0 right after a POP).  Skipped or not, the slot of a level only ever holds (P, t_max(P))
                // for the node P the walk is inside at that level — t_max changes on descents and POPs only, so every write made
                // while inside P writes the same pair — and the reference skips exactly the writes that are redundant (after a
                // POP: the slot was just read) or never read (equal exit times: the POP that leaves the child leaves P too).  The
                // walk therefore reads the same pairs with or without `h`; without it the loop has three instructions and one
                // live register less.
                stack.push_sf(sf, parent, t_max);
                parent = nodes.child_of(parent, nd_, shift);                 // :103
                nd_ = nodes.load(parent);
                child = 0u;
                sf = half;                                                   // --scale
                const float mx = kUnit ? fmaf(half, tcx, cx) : half * tcx + cx;   // the centre planes' crossing times, :88
                const float my = kUnit ? fmaf(half, tcy, cy) : half * tcy + cy;
                const float mz = kUnit ? fmaf(half, tcz, cz) : half * tcz + cz;
                if (mx > t_min) { child ^= 1u; px += half; }                 // :107-109
                if (my > t_min) { child ^= 2u; py += half; }
                if (mz > t_min) { child ^= 4u; pz += half; }
                t_max = tv_max;
                return kGuard ? sf > guard_sf : true;                        // :72
            }
        }
        // ADVANCE (:113-122).  The step is applied once, for both continuations.  POP needs the position before the step as well:
        // on a stepped axis that is exactly p + sf (grid-aligned values, exact subtraction and addition) — three FADDs on the FMA
        // pipe inside the pop path instead of copies or selects on the ALU pipe, and popping lanes share the step with the others.
        const bool sx = cx <= tc_max, sy = cy <= tc_max, sz = cz <= tc_max;   // :115-118
        uint32_t step_mask = sx ? 1u : 0u;
        if (sy) step_mask ^= 2u;
        if (sz) step_mask ^= 4u;
        t_min = tc_max;
        // (opaque to the compiler, here and below: without it t_min = tc_max is re-materialised behind the bounds test and twice
        // inside the pop path, and the pop path reuses the differences p - sf: selects instead of predicated FADDs)
        asm volatile("" : "+f"(t_min));
        if (kBounds) { if (t_min > t_limit) return false; }                 // outside everything solid: a miss, whatever follows
        child ^= step_mask;
        face = step_mask;
        if (sx) px -= sf;
        if (sy) py -= sf;
        if (sz) pz -= sf;
        asm volatile("" : "+f"(px), "+f"(py), "+f"(pz));
        if (child & step_mask) {                                             // :124-145, see Trav::step
            uint32_t diff = 0u;                                              // the bits the step changed, stepped axes only
            if (sx) diff |= __float_as_uint(px) ^ __float_as_uint(px + sf);
            if (sy) diff |= __float_as_uint(py) ^ __float_as_uint(py + sf);
            if (sz) diff |= __float_as_uint(pz) ^ __float_as_uint(pz + sf);
            const uint32_t ix = __float_as_uint(px), iy = __float_as_uint(py), iz = __float_as_uint(pz);
            int scale;                                                       // index of the highest differing bit (:132): FLO
            asm("bfind.u32 %0, %1;" : "=r"(scale) : "r"(diff));
            if (scale >= kSvoMaxDepth) return false;                         // left the root cube: a miss (a stepped p is < 1 now)
            stack.pop(scale, parent, t_max);                                 // :134-136
            nd_ = nodes.load(parent);
            sf = __uint_as_float(uint32_t(scale + 104) << 23);               // :133
            const uint32_t bit = 1u << scale, keep = 0u - bit;               // (x >> s) << s == x & -(1 << s)
            px = __uint_as_float(ix & keep);                                 // :137-142
            py = __uint_as_float(iy & keep);
            pz = __uint_as_float(iz & keep);
            child = ((ix & bit) + 2u * (iy & bit) + 4u * (iz & bit)) >> scale;   // :143
            return kGuard ? scale > guard : true;
        }
        return true;
    }

    // How the walk ended, read off its final state.  step() returns false (a) on a hit, with the state of the trip's start; (b) when
    // the guard stops it: sf <= guard_sf, which no trip starts with; (c) on the bounds exit: t_min > t_limit — every ADVANCE
    // that continues has t_min <= t_limit, and a walk that STARTS beyond its limit cannot hit without an ADVANCE (it would have to
    // start inside a solid voxel, i.e. inside the bounds); (d) when a step leaves the root cube: only a step from p = 1 does
    // that, and it leaves p = 1 - sf < 1 behind — no other state has a coordinate below 1.
    __device__ __forceinline__ bool hit(float guard_sf) const {
        bool miss = fminf(px, fminf(py, pz)) < 1.0f;
        if (kGuard) miss = miss || !(sf > guard_sf);
        if (kBounds) miss = miss || t_min > t_limit;
        return !miss;
    }

    __device__ __forceinline__ void result(LsvoResult& r, float guard_sf) const {
        r.px = px; r.py = py; r.pz = pz;
        r.t_min = t_min; r.scale_f = sf; r.scale = int(__float_as_uint(sf) >> 23) - 104; r.face = face; r.mirror = mirror & 7u;
        r.complexity = (mirror & 8u) ? 0u : uint32_t(iters_f); r.hit = hit(guard_sf);
        r.dx = dx; r.dy = dy; r.dz = dz;
    }
};

template <bool kCone, bool kBounds = false, bool kGuard = true, bool kUnit = false, typename Nodes, typename Stack>
__device__ __forceinline__ void lsvo_cast_ray2(const Nodes& nodes, Stack& stack, int guard, float guard_sf, float ox, float oy, float oz,
                                               float dx, float dy, float dz, float coef, float bias, LsvoResult& r, float t_floor = 0.0f,
                                               const SceneBounds* bounds = nullptr) {
    Trav2<kCone, kBounds, kGuard, kUnit> t;
    t.init(ox, oy, oz, dx, dy, dz, coef, bias, t_floor, bounds);
    t.prime(nodes);
    while (t.step(nodes, stack, guard, guard_sf)) {}
    t.result(r, guard_sf);
}
__device__ __forceinline__ float guard_scale_f(int guard) { return __uint_as_float(uint32_t(guard + 104) << 23); }
__device__ __forceinline__ float pin(float v) { return v + float(blockIdx.y); }

}  // namespace vrt

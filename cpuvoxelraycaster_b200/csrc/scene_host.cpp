// Host-side scene construction for libvrt: the demo terrain's heightfield and the flattening of an
// octree into the reference's LNode layout, done directly from heights / voxel lists so that the
// 80-byte pointer nodes of SVO<N> (svo.hpp:7-25; ~13 GB at 2048^3) are never materialised.
//
// Replaces, with bit-identical output in every defined field:
//   src/main.cpp:61-76                     FastNoise SimplexFractal heights + SVO::setCell fill
//   include/svo.hpp:72-76,91-114            SVO::setCell / rec_setCell
//   include/lsvo_utils.hpp:45-55,
//   src/lsvo_utils.cpp:4-49                 compileSVO / compileSVO_rec (DFS pre-order, 8 slots per
//                                           non-empty node, children visited x-outer / z-inner)
// Compiled with -ffp-contract=off: the heights must match the reference's to the integer.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/vrt.h"
#include "host_util.h"

namespace vrt {

// ---- 2-D simplex FBM (FastNoise 0.4.1 defaults: seed 1337, frequency 0.01, 3 octaves, ------------
// ---- lacunarity 2, gain 0.5; lib/fastnoise/FastNoise.cpp:197-227,410-447,1191-1207,1275-1333) -----
class SimplexFbm2D {
public:
    explicit SimplexFbm2D(int seed = 1337) {
        std::mt19937_64 gen(seed);
        for (int i = 0; i < 256; ++i) perm_[i] = uint8_t(i);
        for (int j = 0; j < 256; ++j) {
            const int k = int(gen() % uint64_t(256 - j)) + j;
            std::swap(perm_[j], perm_[k]);
            perm_[j + 256] = perm_[j];
        }
        for (int j = 0; j < 512; ++j) perm12_[j] = uint8_t(perm_[j] % 12);
        float amp = gain_, total = 1.0f;
        for (int i = 1; i < octaves_; ++i) { total += amp; amp *= gain_; }
        bounding_ = 1.0f / total;
    }

    float fbm(float x, float y) const {
        x *= frequency_;
        y *= frequency_;
        float sum = octave(perm_[0], x, y), amp = 1.0f;
        for (int i = 1; i < octaves_; ++i) {
            x *= lacunarity_;
            y *= lacunarity_;
            amp *= gain_;
            sum += octave(perm_[i], x, y) * amp;
        }
        return sum * bounding_;
    }

    void tables(uint8_t perm[512], uint8_t perm12[512], float* bounding) const {
        std::memcpy(perm, perm_, 512);
        std::memcpy(perm12, perm12_, 512);
        *bounding = bounding_;
    }

private:
    static int floor_to_int(float f) { return f >= 0 ? int(f) : int(f) - 1; }

    float corner(uint8_t offset, int ix, int iy, float fx, float fy) const {
        float t = 0.5f - fx * fx - fy * fy;
        if (t < 0) return 0.0f;
        static const float gx[12] = {1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
        static const float gy[12] = {1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
        const uint8_t g = perm12_[(ix & 0xff) + perm_[(iy & 0xff) + offset]];
        t *= t;
        return t * t * (fx * gx[g] + fy * gy[g]);
    }

    float octave(uint8_t offset, float x, float y) const {
        const float sqrt3 = 1.7320508075688772935274463415059f;
        const float skew = 0.5f * (sqrt3 - 1.0f), unskew = (3.0f - sqrt3) / 6.0f;
        float t = (x + y) * skew;
        const int i = floor_to_int(x + t), j = floor_to_int(y + t);
        t = float(i + j) * unskew;
        const float x0 = x - (float(i) - t), y0 = y - (float(j) - t);
        const int i1 = x0 > y0 ? 1 : 0, j1 = 1 - i1;
        const float x1 = x0 - float(i1) + unskew, y1 = y0 - float(j1) + unskew;
        const float x2 = x0 - 1 + 2 * unskew, y2 = y0 - 1 + 2 * unskew;
        const float n0 = corner(offset, i, j, x0, y0);
        const float n1 = corner(offset, i + i1, j + j1, x1, y1);
        const float n2 = corner(offset, i + 1, j + 1, x2, y2);
        return 70 * (n0 + n1 + n2);
    }

    uint8_t perm_[512], perm12_[512];
    const float frequency_ = 0.01f, lacunarity_ = 2.0f, gain_ = 0.5f;
    const int octaves_ = 3;
    float bounding_;
};

// permutation tables + fractal bounding for the device-side noise (scene_device.cu)
void host_simplex_tables(uint8_t perm[512], uint8_t perm12[512], float* bounding) { SimplexFbm2D().tables(perm, perm12, bounding); }

float host_noise2d(float x, float y) {
    static const SimplexFbm2D noise;
    return noise.fbm(x, y);
}

void host_terrain_heights(int32_t size, int32_t* out) {
    const SimplexFbm2D noise;
    for (int32_t x = 0; x < size; ++x)
        for (int32_t z = 0; z < size; ++z)   // main.cpp:68
            out[size_t(x) * size + z] = int32_t(64.0f * noise.fbm(0.75f * float(uint32_t(x)), 0.75f * float(uint32_t(z))) + 32);
}

// ---- flattening ------------------------------------------------------------------------------------
namespace {

struct Emitter {
    vrt_lnode* out;
    uint64_t cap, count;
    void fresh_slot() {
        if (out && count < cap) out[count] = vrt_lnode{1, 0, 0, 0, 0};   // LNode() defaults, lsvo_utils.hpp:7-12
        ++count;
    }
    vrt_lnode* at(uint64_t i) { return (out && i < cap) ? out + i : nullptr; }
};

// Occupancy predicate for the terrain: column (x,z) is solid for y in [S/2+1, top(x,z)].
struct TerrainOcc {
    int S, bottom;
    std::vector<std::vector<int32_t>> top;   // top[l][(x>>l)*(S>>l)+(z>>l)] = max column top over the square
    bool nonempty(int l, int x0, int y0, int z0) const {
        const int32_t t = top[l][size_t(x0 >> l) * size_t(S >> l) + size_t(z0 >> l)];
        return t >= y0 && t >= bottom && y0 + (1 << l) - 1 >= bottom;
    }
};

template <typename Occ>
void flatten(Emitter& e, const Occ& occ, int depth) {
    struct Frame { uint64_t idx; int level, x, y, z; };
    // Explicit DFS; children are pushed in reverse visit order so that they pop in the reference's
    // order: x outer, y middle, z inner (lsvo_utils.cpp:29-31).
    std::vector<Frame> todo;
    e.fresh_slot();                                      // root, lsvo_utils.hpp:49
    todo.push_back({0, depth, 0, 0, 0});
    while (!todo.empty()) {
        const Frame f = todo.back();
        todo.pop_back();
        const uint64_t child_pos = e.count;              // lsvo_utils.cpp:7
        if (vrt_lnode* n = e.at(f.idx)) n->child_offset = uint32_t(child_pos - f.idx);
        const int half = 1 << (f.level - 1);
        uint32_t present = 0;
        for (int c = 0; c < 8; ++c)                      // c = z*4 + y*2 + x (lsvo_utils.cpp:34)
            if (occ.nonempty(f.level - 1, f.x + (c & 1) * half, f.y + ((c >> 1) & 1) * half, f.z + ((c >> 2) & 1) * half))
                present |= 1u << c;
        if (!present) continue;                          // :12-23 (only an empty root gets here)
        for (int i = 0; i < 8; ++i) e.fresh_slot();      // :25-27
        if (vrt_lnode* n = e.at(f.idx)) {
            n->child_mask = uint8_t(present);
            if (f.level == 1) n->leaf_mask = uint8_t(present);   // children are voxels: :40-42
        }
        if (f.level == 1) continue;
        // reference visit order of slots: (x,y,z) lexicographic → slot = z*4+y*2+x : 0,4,2,6,1,5,3,7
        static const int order[8] = {0, 4, 2, 6, 1, 5, 3, 7};
        for (int k = 7; k >= 0; --k) {
            const int c = order[k];
            if (present & (1u << c))
                todo.push_back({child_pos + uint64_t(c), f.level - 1, f.x + (c & 1) * half, f.y + ((c >> 1) & 1) * half,
                                f.z + ((c >> 2) & 1) * half});
        }
    }
}

}  // namespace

// NB: a DFS with an explicit LIFO only reproduces the recursive pre-order numbering if a node's whole
// subtree is emitted before its next sibling starts — which holds here because child blocks are
// allocated when a frame is popped, and frames pop in pre-order.
uint64_t host_build_terrain_lsvo(uint32_t depth, const int32_t* heights, vrt_lnode* out, uint64_t cap) {
    const int S = 1 << depth;
    TerrainOcc occ;
    occ.S = S;
    occ.bottom = S / 2 + 1;
    occ.top.resize(depth + 1);
    occ.top[0].resize(size_t(S) * S);
    for (size_t i = 0; i < size_t(S) * S; ++i) {
        // main.cpp:71-72; columns cannot leave the world: tops are clamped to S/2 like tops_kernel (scene_device.cu) does,
        // so the count-only call, the fill call and the device builder agree for any caller-supplied heights
        const int32_t hmax = std::max(16, std::min(S / 2, heights[i]));
        occ.top[0][i] = S / 2 + hmax - 1;                             // y in [1,hmax) stored at y + S/2
    }
    for (uint32_t l = 1; l <= depth; ++l) {
        const size_t n = size_t(S >> l), m = n * 2;
        occ.top[l].resize(n * n);
        const std::vector<int32_t>& lo = occ.top[l - 1];
        for (size_t x = 0; x < n; ++x)
            for (size_t z = 0; z < n; ++z)
                occ.top[l][x * n + z] = std::max(std::max(lo[(2 * x) * m + 2 * z], lo[(2 * x) * m + 2 * z + 1]),
                                                 std::max(lo[(2 * x + 1) * m + 2 * z], lo[(2 * x + 1) * m + 2 * z + 1]));
    }
    if (!out) {
        // slot count without the DFS: 1 + 8 * (number of non-empty cubes of edge >= 2)
        uint64_t interior = 0;
        for (uint32_t l = 1; l <= depth; ++l) {
            const size_t n = size_t(S >> l);
            const int edge = 1 << l;
            for (size_t i = 0; i < n * n; ++i) {
                const int32_t t = occ.top[l][i];
                if (t < occ.bottom) continue;
                interior += uint64_t(t / edge - occ.bottom / edge + 1);   // y-cubes intersecting [bottom, t]
            }
        }
        return 1 + 8 * interior;
    }
    Emitter e{out, cap, 0};
    flatten(e, occ, int(depth));
    return e.count;
}

// Voxel list → LNode array.  Voxels are keyed by their DFS visit order (per level the digit
// x*4+y*2+z, most significant level first), sorted, and the tree is emitted range by range.
uint64_t host_build_lsvo_from_voxels(uint32_t depth, const uint32_t* xyz, uint64_t n_voxels, vrt_lnode* out, uint64_t cap) {
    std::vector<uint64_t> keys(n_voxels);
    for (uint64_t i = 0; i < n_voxels; ++i) {
        const uint32_t x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        uint64_t k = 0;
        for (int b = int(depth) - 1; b >= 0; --b)
            k = (k << 3) | uint64_t((((x >> b) & 1u) << 2) | (((y >> b) & 1u) << 1) | ((z >> b) & 1u));
        keys[i] = k;
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());

    Emitter e{out, cap, 0};
    e.fresh_slot();
    struct Frame { uint64_t idx, lo, hi; int level; };   // keys[lo,hi) share the digits above `level`
    std::vector<Frame> todo;
    todo.push_back({0, 0, keys.size(), int(depth)});
    while (!todo.empty()) {
        const Frame f = todo.back();
        todo.pop_back();
        const uint64_t child_pos = e.count;
        if (vrt_lnode* n = e.at(f.idx)) n->child_offset = uint32_t(child_pos - f.idx);
        if (f.lo == f.hi) continue;                      // empty root
        for (int i = 0; i < 8; ++i) e.fresh_slot();
        const int shift = 3 * (f.level - 1);
        uint64_t bounds[9];
        bounds[0] = f.lo;
        for (int dgt = 0; dgt < 8; ++dgt) {              // digit = x*4 + y*2 + z, ascending = visit order
            uint64_t p = bounds[dgt];
            while (p < f.hi && ((keys[p] >> shift) & 7u) == uint64_t(dgt)) ++p;
            bounds[dgt + 1] = p;
        }
        uint32_t present = 0;
        for (int dgt = 0; dgt < 8; ++dgt)
            if (bounds[dgt + 1] > bounds[dgt]) {
                const int slot = ((dgt & 1) << 2) | (dgt & 2) | ((dgt >> 2) & 1);   // z*4 + y*2 + x
                present |= 1u << slot;
            }
        if (vrt_lnode* n = e.at(f.idx)) {
            n->child_mask = uint8_t(present);
            if (f.level == 1) n->leaf_mask = uint8_t(present);
        }
        if (f.level == 1) continue;
        for (int dgt = 7; dgt >= 0; --dgt)
            if (bounds[dgt + 1] > bounds[dgt]) {
                const int slot = ((dgt & 1) << 2) | (dgt & 2) | ((dgt >> 2) & 1);
                todo.push_back({child_pos + uint64_t(slot), bounds[dgt], bounds[dgt + 1], f.level - 1});
            }
    }
    return e.count;
}

// ---- camera basis: Camera::setViewAngle (camera_controller.hpp:27-32) --------------------------------
// generateRotationMatrix (utils.cpp:94-100) = rotate(I, -angle.y, x̂) * rotate(I, -angle.x, ŷ) with
// glm::rotate's Rodrigues form; written out for the two axis-aligned cases it is called with.
namespace {
struct Mat3 { float c[3][3]; };   // c[col][row]

// glm::rotate(mat4(1), a, axis) restricted to 3x3, axis = unit x (k=0) or unit y (k=1)
Mat3 axis_rotation(float a, int k) {
    const float c = std::cos(a), s = std::sin(a);
    float axis[3] = {0.0f, 0.0f, 0.0f};
    axis[k] = 1.0f;
    {   // glm normalises the axis: v * (1/sqrt(dot(v,v)))
        const float inv = 1.0f / std::sqrt((axis[0] * axis[0] + axis[1] * axis[1]) + axis[2] * axis[2]);
        for (float& v : axis) v *= inv;
    }
    float temp[3];
    for (int i = 0; i < 3; ++i) temp[i] = (1.0f - c) * axis[i];
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    // Result[j] = I[0]*R[j][0] + I[1]*R[j][1] + I[2]*R[j][2]  (identity input, summed left to right)
    Mat3 m;
    for (int j = 0; j < 3; ++j)
        for (int r = 0; r < 3; ++r) {
            const float i0 = r == 0 ? 1.0f : 0.0f, i1 = r == 1 ? 1.0f : 0.0f, i2 = r == 2 ? 1.0f : 0.0f;
            m.c[j][r] = (i0 * R[j][0] + i1 * R[j][1]) + i2 * R[j][2];
        }
    return m;
}
}  // namespace

void host_camera_rotation(const float view_angle[2], float rot_mat[9], float camera_vec[3]) {
    const Mat3 rx = axis_rotation(-view_angle[0], 1);   // about y
    const Mat3 ry = axis_rotation(-view_angle[1], 0);   // about x
    // (ry * rx)[j] = ry[0]*rx[j][0] + ry[1]*rx[j][1] + ry[2]*rx[j][2] (+ ry[3]*0 in the 4x4 original)
    for (int j = 0; j < 3; ++j)
        for (int r = 0; r < 3; ++r)
            rot_mat[3 * j + r] = ((ry.c[0][r] * rx.c[j][0] + ry.c[1][r] * rx.c[j][1]) + ry.c[2][r] * rx.c[j][2]) + 0.0f * 0.0f;
    // viewToWorld((0,0,1)) = v * rot_mat: component j = m[j][0]*0 + m[j][1]*0 + m[j][2]*1
    for (int j = 0; j < 3; ++j)
        camera_vec[j] = (rot_mat[3 * j + 0] * 0.0f + rot_mat[3 * j + 1] * 0.0f) + rot_mat[3 * j + 2] * 1.0f;
}

}  // namespace vrt

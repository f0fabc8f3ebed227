// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI harness around the two reference structures that need a
// build-time sed patch to be usable at all (the patch is applied to temporary copies by
// oracle/Makefile, the bodies of castRay are untouched):
//   Grid3D<X,Y,Z>  include/grid_3d.hpp   — `: public Volumetric` / ` override` removed (the 2-arg
//                                           castRay does not override the 4-arg pure virtual)
//   SVO<N>         include/svo.hpp       — the /* */ around fillHitResult's body (svo.hpp:118,137)
//                                           removed: the "intended SVO" that reports hits
// No arithmetic of its own; built into oracle/_ref/libvrt_ref_patched.so.
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <SFML/Graphics.hpp>
#include <glm/glm.hpp>

#include "svo.hpp"      // patched copy first on the include path
#include "grid_3d.hpp"  // patched copy

namespace sf {
const Color Color::Black(0, 0, 0), Color::White(255, 255, 255), Color::Red(255, 0, 0), Color::Green(0, 255, 0),
    Color::Blue(0, 0, 255), Color::Yellow(255, 255, 0), Color::Magenta(255, 0, 255), Color::Cyan(0, 255, 255),
    Color::Transparent(0, 0, 0, 0);
std::map<std::string, std::vector<Uint8>>& shim_texture_registry() {
    static std::map<std::string, std::vector<Uint8>> reg;
    return reg;
}
}  // namespace sf

extern "C" struct vrt_ref_hit {
    float position[3];
    float normal[3];
    float voxel_coord[2];
    float distance;
    uint32_t complexity;
    uint32_t hit;
    uint32_t pad;
};

namespace {

void store(const HitPoint& h, vrt_ref_hit* r) {
    std::memset(r, 0, sizeof(*r));
    r->complexity = h.complexity;
    if (h.cell) {
        r->hit = 1;
        r->position[0] = h.position.x; r->position[1] = h.position.y; r->position[2] = h.position.z;
        r->normal[0] = h.normal.x; r->normal[1] = h.normal.y; r->normal[2] = h.normal.z;
        r->voxel_coord[0] = h.voxel_coord.x; r->voxel_coord[1] = h.voxel_coord.y;
        r->distance = h.distance;
    }
}

struct GridBase {
    virtual ~GridBase() {}
    virtual void set(uint32_t x, uint32_t y, uint32_t z) = 0;
    virtual void cast(const float* o, const float* d, uint64_t n, vrt_ref_hit* out) const = 0;
};
template <int32_t N> struct GridT : GridBase {
    std::unique_ptr<Grid3D<N, N, N>> g{new Grid3D<N, N, N>()};   // 8 B/cell: 1 GiB at N=512
    void set(uint32_t x, uint32_t y, uint32_t z) override { g->setCell(Cell::Solid, x, y, z); }
    void cast(const float* o, const float* d, uint64_t n, vrt_ref_hit* out) const override {
        for (uint64_t i = 0; i < n; ++i)
            store(g->castRay(glm::vec3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), glm::vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2])), out + i);
    }
};

struct SvoBase {
    virtual ~SvoBase() {}
    virtual void set(uint32_t x, uint32_t y, uint32_t z) = 0;
    virtual void cast(const float* o, const float* d, uint32_t max_iter, uint64_t n, vrt_ref_hit* out) const = 0;
};
template <uint8_t N> struct SvoT : SvoBase {
    SVO<N> svo;
    void set(uint32_t x, uint32_t y, uint32_t z) override { svo.setCell(Cell::Solid, Cell::Grass, x, y, z); }
    void cast(const float* o, const float* d, uint32_t max_iter, uint64_t n, vrt_ref_hit* out) const override {
        for (uint64_t i = 0; i < n; ++i)
            store(svo.castRay(glm::vec3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), glm::vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), max_iter), out + i);
    }
};

}  // namespace

extern "C" {

// cubic grids of edge 2^log2_edge, 3 <= log2_edge <= 9; occ[(x*N+y)*N+z] != 0 → Cell::Solid
void* vrt_ref_grid_create(int log2_edge, const uint8_t* occ) {
    GridBase* g = nullptr;
    switch (log2_edge) {
        case 3: g = new GridT<8>(); break;
        case 4: g = new GridT<16>(); break;
        case 5: g = new GridT<32>(); break;
        case 6: g = new GridT<64>(); break;
        case 7: g = new GridT<128>(); break;
        case 8: g = new GridT<256>(); break;
        case 9: g = new GridT<512>(); break;
        default: return nullptr;
    }
    const uint32_t N = 1u << log2_edge;
    for (uint32_t x = 0; x < N; ++x)
        for (uint32_t y = 0; y < N; ++y)
            for (uint32_t z = 0; z < N; ++z)
                if (occ[(size_t(x) * N + y) * N + z]) g->set(x, y, z);
    return g;
}
void vrt_ref_grid_destroy(void* g) { delete static_cast<GridBase*>(g); }
void vrt_ref_grid_cast(void* g, const float* o, const float* d, uint64_t n, vrt_ref_hit* out) {
    static_cast<GridBase*>(g)->cast(o, d, n, out);
}

void* vrt_ref_svo_create(int depth, const uint8_t* occ) {
    SvoBase* s = nullptr;
    switch (depth) {
        case 2: s = new SvoT<2>(); break;
        case 3: s = new SvoT<3>(); break;
        case 4: s = new SvoT<4>(); break;
        case 5: s = new SvoT<5>(); break;
        case 6: s = new SvoT<6>(); break;
        case 7: s = new SvoT<7>(); break;
        case 8: s = new SvoT<8>(); break;
        case 9: s = new SvoT<9>(); break;
        default: return nullptr;
    }
    const uint32_t N = 1u << depth;
    for (uint32_t x = 0; x < N; ++x)
        for (uint32_t y = 0; y < N; ++y)
            for (uint32_t z = 0; z < N; ++z)
                if (occ[(size_t(x) * N + y) * N + z]) s->set(x, y, z);
    return s;
}
void vrt_ref_svo_destroy(void* s) { delete static_cast<SvoBase*>(s); }
void vrt_ref_svo_cast(void* s, const float* o, const float* d, uint32_t max_iter, uint64_t n, vrt_ref_hit* out) {
    static_cast<SvoBase*>(s)->cast(o, d, max_iter, n, out);
}

}  // extern "C"

// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI harness around the UNMODIFIED reference sources.
//
// This file contains no traversal or shading arithmetic of its own.  It #includes the reference
// headers where they lie under /root/reference (include/lsvo.hpp, svo.hpp, raycaster.hpp,
// camera_controller.hpp, lib/fastnoise/FastNoise.h, lib/swarm/swarm.hpp) and forwards C calls to
// them, so that tests can pin the restatement in oracle/port.c and the CUDA path against the
// reference's own code, and bench.py can time the reference's CPU path ("kind": "reference").
// Built by oracle/Makefile into oracle/_ref/libvrt_ref.so (git-ignored, travels to the GPU box).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the result.  The product (cpuvoxelraycaster_b200/) never does.
//
// VRT_REF_DEPTH: the reference hard-codes SVO_DEPTH = 9 (raycaster.hpp:42) and 1/512
// (raycaster.hpp:171, camera_controller.hpp:36,58).  The default build (=9) compiles those headers
// verbatim; oracle/Makefile builds extra depth variants from sed-patched temporary copies.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <SFML/Graphics.hpp>
#include <glm/glm.hpp>

#include "svo.hpp"
#include "lsvo.hpp"
#include "utils.hpp"
#include "raycaster.hpp"
#include "camera_controller.hpp"
#include "fastnoise/FastNoise.h"
#include "swarm/swarm.hpp"

#ifndef VRT_REF_DEPTH
#define VRT_REF_DEPTH 9
#endif

// ---- SFML stub statics -------------------------------------------------------------------------
namespace sf {
const Color Color::Black(0, 0, 0), Color::White(255, 255, 255), Color::Red(255, 0, 0), Color::Green(0, 255, 0),
    Color::Blue(0, 0, 255), Color::Yellow(255, 255, 0), Color::Magenta(255, 0, 255), Color::Cyan(0, 255, 255),
    Color::Transparent(0, 0, 0, 0);
std::map<std::string, std::vector<Uint8>>& shim_texture_registry() {
    static std::map<std::string, std::vector<Uint8>> reg;
    return reg;
}
}  // namespace sf

extern "C" {

// Mirrors HitPoint (volumetric.hpp:7-22) field by field; `hit` replaces the borrowed Cell pointer.
// On a miss the reference leaves position/normal/voxel_coord/distance uninitialised: they are
// reported as zero here and only `hit`/`complexity` are meaningful.
struct vrt_ref_hit {
    float position[3];
    float normal[3];
    float voxel_coord[2];
    float distance;
    uint32_t complexity;
    uint32_t hit;
    uint32_t pad;
};

struct vrt_ref_render_params {
    int32_t width, height;
    float cam_position[3];   // voxel units, Camera::position (main.cpp:51)
    float view_angle[2];     // Camera::setViewAngle
    float fov, aperture, focal_length;
    float light_position[3];  // already normalised: light*scale + 1 (main.cpp:126)
    int32_t use_gi, use_samples;
    int32_t spp;              // number of renderRay passes per pixel
    int32_t threads;          // 1 = deterministic single thread, >1 = swarm tiles (racy RNG, as the reference)
    int32_t row_begin, row_end;  // rows [begin,end) only (bounded samples for the CPU baseline)
    int32_t tile_step, tile_index;  // > 1: only 4-row tiles t (from row_begin) with t % tile_step == tile_index
    int32_t checker;          // 0 = every pixel; 1 / 2 = checker_board_offset 0 / 1 in the FIRST frame (main.cpp:137,143)
    int32_t checker_area_height;  // RENDER_HEIGHT / area_count (main.cpp:132); 0 = a single area starting at row 0
    int32_t frames;           // > 1: that many consecutive frames on the same RayCaster (the temporal blend of
                              // raycaster.hpp:79-85 accumulates), the checkerboard offset flipping each frame (main.cpp:137)
};

// castRay invocations seen by the counting subclass below, split by cone coefficient:
// coef == 0 → primary + sun-shadow calls (raycaster.hpp:131,153), coef != 0 → GI + GI-shadow (:194,:198)
struct vrt_ref_ray_counts {
    uint64_t cone0_calls;
    uint64_t cone_gi_calls;
};

}  // extern "C"

namespace {

struct SceneBase {
    int depth = 0;
    virtual ~SceneBase() {}
    virtual const std::vector<LNode>& nodes() const = 0;
    virtual void cast(const float* o, const float* d, float coef, float bias, uint64_t begin, uint64_t end,
                      vrt_ref_hit* out) const = 0;
    virtual const void* lsvo_ptr() const = 0;
};

template <uint8_t D> struct SceneT : SceneBase {
    LSVO<D>* lsvo = nullptr;
    SceneT(const SVO<D>& svo) {
        depth = D;
        lsvo = new LSVO<D>(svo);
    }
    ~SceneT() override { delete lsvo; }  // (the shared Cell allocated at lsvo.hpp:21 leaks, as in the reference)
    const std::vector<LNode>& nodes() const override { return lsvo->data; }
    const void* lsvo_ptr() const override { return lsvo; }
    void cast(const float* o, const float* d, float coef, float bias, uint64_t begin, uint64_t end,
              vrt_ref_hit* out) const override {
        for (uint64_t i = begin; i < end; ++i) {
            const glm::vec3 p(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
            const glm::vec3 dir(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
            const HitPoint h = lsvo->castRay(p, dir, coef, bias);
            vrt_ref_hit r;
            std::memset(&r, 0, sizeof(r));
            r.complexity = h.complexity;
            if (h.cell) {
                r.hit = 1;
                r.position[0] = h.position.x; r.position[1] = h.position.y; r.position[2] = h.position.z;
                r.normal[0] = h.normal.x; r.normal[1] = h.normal.y; r.normal[2] = h.normal.z;
                r.voxel_coord[0] = h.voxel_coord.x; r.voxel_coord[1] = h.voxel_coord.y;
                r.distance = h.distance;
            }
            out[i] = r;
        }
    }
};

template <uint8_t D> SceneBase* make_from_voxels(const uint32_t* xyz, uint64_t n) {
    SVO<D>* svo = new SVO<D>();
    for (uint64_t i = 0; i < n; ++i) svo->setCell(Cell::Solid, Cell::Grass, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    SceneBase* s = new SceneT<D>(*svo);
    delete svo;
    return s;
}

// main.cpp:59-76 with the literal 256 (= size/2 at depth 9) written as size/2.
template <uint8_t D> SceneBase* make_terrain() {
    const int32_t size = 1 << D;
    SVO<D>* svo = new SVO<D>();
    FastNoise noise;
    noise.SetNoiseType(FastNoise::SimplexFractal);
    for (uint32_t x = 0; x < uint32_t(size); x++) {
        for (uint32_t z = 0; z < uint32_t(size); z++) {
            const int32_t max_height = size;
            const int32_t height = int32_t(64.0f * noise.GetNoise(float(0.75f * x), float(0.75f * z)) + 32);
            const int32_t ground_level = 16;
            for (int y(1); y < std::max(ground_level, std::min(max_height, height)); ++y)
                svo->setCell(Cell::Solid, Cell::Grass, x, y + size / 2, z);
        }
    }
    SceneBase* s = new SceneT<D>(*svo);
    delete svo;
    return s;
}

template <uint8_t D> SceneBase* make_from_nodes(const LNode* nodes, uint64_t n) {
    SVO<D>* svo = new SVO<D>();            // empty tree: compileSVO yields the single root slot
    SceneT<D>* s = new SceneT<D>(*svo);
    delete svo;
    s->lsvo->data.assign(nodes, nodes + n);   // LSVO::data / raw_data are public (lsvo.hpp:287-288)
    s->lsvo->raw_data = &(s->lsvo->data[0]);
    return s;
}

// LSVO<D>::castRay is virtual (Volumetric, volumetric.hpp:58) and RayCaster calls it through a
// `const LSVO<SVO_DEPTH>&` (raycaster.hpp:265), so a subclass that forwards to the unmodified
// implementation can count invocations for the Mrays/s metric without touching reference code.
struct alignas(64) PaddedCounts { uint64_t cone0 = 0, cone_gi = 0; };
static PaddedCounts g_counts[256];
static std::atomic<uint32_t> g_next_thread_slot(0);
static thread_local int t_slot = -1;

struct CountingLSVO : LSVO<VRT_REF_DEPTH> {
    CountingLSVO(const SVO<VRT_REF_DEPTH>& svo) : LSVO<VRT_REF_DEPTH>(svo) {}
    HitPoint castRay(const glm::vec3& position, glm::vec3 d, const float ray_size_coef = 0.0f,
                     const float ray_size_bias = 0.0f) const override {
        if (t_slot < 0) t_slot = int(g_next_thread_slot.fetch_add(1) % 256u);
        if (ray_size_coef == 0.0f) ++g_counts[t_slot].cone0; else ++g_counts[t_slot].cone_gi;
        return LSVO<VRT_REF_DEPTH>::castRay(position, d, ray_size_coef, ray_size_bias);
    }
};

#define VRT_DISPATCH_DEPTH(depth, CALL)                                                                     \
    switch (depth) {                                                                                        \
        case 1: return CALL(1); case 2: return CALL(2); case 3: return CALL(3); case 4: return CALL(4);     \
        case 5: return CALL(5); case 6: return CALL(6); case 7: return CALL(7); case 8: return CALL(8);     \
        case 9: return CALL(9); case 10: return CALL(10); case 11: return CALL(11); case 12: return CALL(12); \
        default: return nullptr;                                                                            \
    }

// Run job(worker, n_workers) on n threads through the reference's own pool.  swrm::Swarm silently
// drops an execute() when its workers have not re-registered (swarm.hpp:221-223), so every run is
// verified with a counter and retried on a fresh pool; std::thread is the last resort.
template <typename F> int run_parallel(uint32_t n, F job) {
    if (n <= 1) { job(0u, 1u); return 0; }
    for (int attempt = 0; attempt < 3; ++attempt) {
        std::atomic<uint32_t> started(0);
        {
            swrm::Swarm swarm(n);
            std::this_thread::sleep_for(std::chrono::milliseconds(20 + 50 * attempt));
            swrm::WorkGroup g = swarm.execute([&](uint32_t id, uint32_t total) {
                started.fetch_add(1);
                job(id, total);
            });
            g.waitExecutionDone();
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
        if (started.load() == n) return 1;      // ran on the swarm
        if (started.load() != 0) return -1;     // partial run: results are not trustworthy
    }
    std::vector<std::thread> ts;
    for (uint32_t i = 0; i < n; ++i) ts.emplace_back([&, i] { job(i, n); });
    for (auto& t : ts) t.join();
    return 2;                                    // ran on std::thread
}

}  // namespace

extern "C" {

int vrt_ref_compiled_depth() { return VRT_REF_DEPTH; }
unsigned vrt_ref_sizeof_lnode() { return unsigned(sizeof(LNode)); }
unsigned vrt_ref_sizeof_hitpoint() { return unsigned(sizeof(HitPoint)); }

void vrt_ref_register_texture(const char* name, int w, int h, const uint8_t* rgb_top_down) {
    std::vector<sf::Uint8> v;
    v.push_back(sf::Uint8(w));
    v.push_back(sf::Uint8(h));
    v.insert(v.end(), rgb_top_down, rgb_top_down + size_t(w) * h * 3);
    sf::shim_texture_registry()[name] = v;
}

#define CALL_TERRAIN(D) make_terrain<D>()
void* vrt_ref_scene_terrain(int depth) { VRT_DISPATCH_DEPTH(depth, CALL_TERRAIN) }
#define CALL_VOX(D) make_from_voxels<D>(xyz, n)
void* vrt_ref_scene_from_voxels(int depth, const uint32_t* xyz, uint64_t n) { VRT_DISPATCH_DEPTH(depth, CALL_VOX) }
#define CALL_NODES(D) make_from_nodes<D>(static_cast<const LNode*>(nodes), n)
void* vrt_ref_scene_from_nodes(int depth, const void* nodes, uint64_t n) { VRT_DISPATCH_DEPTH(depth, CALL_NODES) }

void vrt_ref_scene_destroy(void* s) { delete static_cast<SceneBase*>(s); }
uint64_t vrt_ref_scene_node_count(void* s) { return static_cast<SceneBase*>(s)->nodes().size(); }
void vrt_ref_scene_copy_nodes(void* s, void* out) {
    const auto& v = static_cast<SceneBase*>(s)->nodes();
    std::memcpy(out, v.data(), v.size() * sizeof(LNode));
}

// FastNoise heights exactly as main.cpp:68 computes them; out[x*S+z].
void vrt_ref_terrain_heights(int size, int32_t* out) {
    FastNoise noise;
    noise.SetNoiseType(FastNoise::SimplexFractal);
    for (uint32_t x = 0; x < uint32_t(size); x++)
        for (uint32_t z = 0; z < uint32_t(size); z++)
            out[size_t(x) * size + z] = int32_t(64.0f * noise.GetNoise(float(0.75f * x), float(0.75f * z)) + 32);
}

void vrt_ref_noise2d(const float* x, const float* y, uint64_t n, float* out) {
    FastNoise noise;
    noise.SetNoiseType(FastNoise::SimplexFractal);
    for (uint64_t i = 0; i < n; ++i) out[i] = noise.GetNoise(x[i], y[i]);
}

// LSVO<D>::castRay over a ray buffer.  threads>1 → interleaved blocks of 4096 rays per worker.
// Returns 0 single thread, 1 swarm, 2 std::thread, -1 swarm failure.
int vrt_ref_lsvo_cast(void* scene, const float* org, const float* dir, float coef, float bias, uint64_t n,
                      vrt_ref_hit* out, int threads) {
    const SceneBase* s = static_cast<SceneBase*>(scene);
    const uint64_t block = 4096;
    return run_parallel(uint32_t(threads < 1 ? 1 : threads), [&](uint32_t id, uint32_t total) {
        for (uint64_t b = uint64_t(id) * block; b < n; b += uint64_t(total) * block)
            s->cast(org, dir, coef, bias, b, std::min(n, b + block), out);
    });
}

// Camera::setViewAngle (camera_controller.hpp:27-32) → rot_mat (column major, 9 floats) and camera_vec.
void vrt_ref_camera_basis(const float view_angle[2], float rot_mat[9], float camera_vec[3]) {
    Camera cam;
    cam.setViewAngle(glm::vec2(view_angle[0], view_angle[1]));
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) rot_mat[3 * c + r] = cam.rot_mat[c][r];
    camera_vec[0] = cam.camera_vec.x; camera_vec[1] = cam.camera_vec.y; camera_vec[2] = cam.camera_vec.z;
}

// Camera::getRay + the lens mapping of main.cpp:145-149 for every pixel of a W×H frame, x outer / y inner
// like the reference loop; origin/dir are written at pixel index y*W+x.  Consumes the global RNG.
void vrt_ref_camera_rays(const vrt_ref_render_params* p, float* origin, float* dir) {
    Camera cam;
    cam.position = glm::vec3(p->cam_position[0], p->cam_position[1], p->cam_position[2]);
    cam.fov = p->fov; cam.aperture = p->aperture; cam.focal_length = p->focal_length;
    cam.setViewAngle(glm::vec2(p->view_angle[0], p->view_angle[1]));
    const float scale = 1.0f / float(1 << VRT_REF_DEPTH);
    const float aspect_ratio = float(p->width) / float(p->height);
    for (int32_t x = 0; x < p->width; ++x)
        for (int32_t y = 0; y < p->height; ++y) {
            const float lens_x = float(x) / float(p->height) - aspect_ratio * 0.5f;
            const float lens_y = float(y) / float(p->height) - 0.5f;
            const CameraRay cr = cam.getRay(glm::vec2(lens_x, lens_y));
            const glm::vec3 start = (cam.position + cr.world_rand_offset) * scale + glm::vec3(1.0f);
            const size_t i = size_t(y) * p->width + x;
            origin[3 * i] = start.x; origin[3 * i + 1] = start.y; origin[3 * i + 2] = start.z;
            dir[3 * i] = cr.ray.x; dir[3 * i + 1] = cr.ray.y; dir[3 * i + 2] = cr.ray.z;
        }
}

// Camera::getClosestPoint (camera_controller.hpp:56-60) → focal length rule of main.cpp:115-121.
float vrt_ref_autofocus(void* scene, const vrt_ref_render_params* p) {
    const SceneBase* s = static_cast<SceneBase*>(scene);
    if (s->depth != VRT_REF_DEPTH) return -1.0f;
    Camera cam;
    cam.position = glm::vec3(p->cam_position[0], p->cam_position[1], p->cam_position[2]);
    cam.setViewAngle(glm::vec2(p->view_angle[0], p->view_angle[1]));
    const LSVO<VRT_REF_DEPTH>& lsvo = *static_cast<const LSVO<VRT_REF_DEPTH>*>(s->lsvo_ptr());
    const HitPoint closest = cam.getClosestPoint(lsvo);
    return closest.cell ? closest.distance * float(1 << VRT_REF_DEPTH) : 100.0f;
}

// The swarm lambda of main.cpp:139-152 without the checkerboard: every pixel, `spp` passes.
//  rgba_raw   [H*W*4] u8   colour RayCaster::castRay returned for the LAST pass (pre temporal blend), or NULL
//  rgba_image [H*W*4] u8   RayCaster::render_image after the passes (+ samples_to_image when use_samples)
//  samples    [H*W*4] f64  r,g,b,count accumulators (use_samples), or NULL
//  ray_counts [4]          distinct castRay invocations: primary, shadow, gi, gi-shadow (recomputed by
//                          replaying the same control flow is impossible with a racy RNG, so the harness
//                          counts inside a single-threaded replica only when threads==1; else zeros)
// Returns the run_parallel code.
int vrt_ref_render(void* scene, const vrt_ref_render_params* p, uint8_t* rgba_raw, uint8_t* rgba_image,
                   double* samples, double* seconds, vrt_ref_ray_counts* counts) {
    const SceneBase* s = static_cast<SceneBase*>(scene);
    if (s->depth != VRT_REF_DEPTH) return -2;
    // same node array behind a counting subclass (the scene's own LSVO is left untouched)
    SVO<VRT_REF_DEPTH>* empty = new SVO<VRT_REF_DEPTH>();
    CountingLSVO counting(*empty);
    delete empty;
    counting.data.clear();
    counting.raw_data = &(s->nodes()[0]);
    for (auto& c : g_counts) c = PaddedCounts();
    const LSVO<VRT_REF_DEPTH>& lsvo = counting;
    RayCaster raycaster(lsvo, sf::Vector2i(p->width, p->height));
    raycaster.use_gi = p->use_gi != 0;
    raycaster.use_samples = p->use_samples != 0;
    raycaster.setLightPosition(glm::vec3(p->light_position[0], p->light_position[1], p->light_position[2]));
    Camera cam;
    cam.position = glm::vec3(p->cam_position[0], p->cam_position[1], p->cam_position[2]);
    cam.fov = p->fov; cam.aperture = p->aperture; cam.focal_length = p->focal_length;
    cam.setViewAngle(glm::vec2(p->view_angle[0], p->view_angle[1]));

    const float scale = 1.0f / float(1 << VRT_REF_DEPTH);
    const uint32_t W = p->width, H = p->height;
    const float aspect_ratio = float(W) / float(H);
    const uint32_t r0 = p->row_begin, r1 = p->row_end > p->row_begin ? uint32_t(p->row_end) : H;
    const uint32_t threads = p->threads < 1 ? 1u : uint32_t(p->threads);
    const float time = 0.0f;

    const auto t0 = std::chrono::steady_clock::now();
    // Workers take interleaved columns; the reference's 4x4 tiles (main.cpp:140-143) are the special
    // case threads == 16 up to the assignment of pixels to workers, which does not change the work.
    const int32_t frames = p->frames > 1 ? p->frames : 1;
    int code = 0;
    for (int32_t frame = 0; frame < frames && code >= 0; ++frame) {
    const int32_t checker_offset = p->checker ? ((p->checker - 1 + frame) & 1) : 0;
    code = run_parallel(threads, [&](uint32_t id, uint32_t total) {
        for (int32_t pass = 0; pass < p->spp; ++pass)
            for (uint32_t x = id; x < W; x += total)
                for (uint32_t y = r0; y < r1; ++y) {
                    if (p->tile_step > 1 && int32_t(((y - r0) >> 2) % uint32_t(p->tile_step)) != p->tile_index) continue;
                    if (p->checker) {   // main.cpp:143: y = area_start + (x + offset) % 2, step 2
                        const uint32_t rel = p->checker_area_height > 0 ? y % uint32_t(p->checker_area_height) : y;
                        if ((rel & 1u) != ((x + uint32_t(checker_offset)) & 1u)) continue;
                    }
                    const float lens_x = float(x) / float(H) - aspect_ratio * 0.5f;
                    const float lens_y = float(y) / float(H) - 0.5f;
                    const CameraRay cr = cam.getRay(glm::vec2(lens_x, lens_y));
                    const glm::vec3 start = (cam.position + cr.world_rand_offset) * scale + glm::vec3(1.0f);
                    if (rgba_raw && pass == p->spp - 1 && threads == 1) {
                        // same call renderRay makes (raycaster.hpp:75) — evaluated on a COPY of the RNG
                        // stream would diverge, so raw output is only offered without GI/aperture noise
                        RayContext ctx;
                        const ColorResult cr_raw = raycaster.castRay(start, cr.ray, 1.5f * time, ctx);
                        uint8_t* q = rgba_raw + 4 * (size_t(y) * W + x);
                        q[0] = cr_raw.color.r; q[1] = cr_raw.color.g; q[2] = cr_raw.color.b; q[3] = 255;
                    }
                    raycaster.renderRay(sf::Vector2i(x, y), start, cr.ray, time);
                }
    });
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (counts) {
        counts->cone0_calls = counts->cone_gi_calls = 0;
        for (const auto& c : g_counts) { counts->cone0_calls += c.cone0; counts->cone_gi_calls += c.cone_gi; }
    }

    if (raycaster.use_samples) {
        if (samples)
            for (uint32_t y = 0; y < H; ++y)
                for (uint32_t x = 0; x < W; ++x) {
                    const Sample& sm = raycaster.colors[x][y];
                    double* q = samples + 4 * (size_t(y) * W + x);
                    q[0] = sm.r; q[1] = sm.g; q[2] = sm.b; q[3] = sm.update_count;
                }
        // samples_to_image divides by update_count; untouched rows would be 0/0 → only resolve when full
        if (r0 == 0 && r1 == H && p->tile_step <= 1) raycaster.samples_to_image();
    }
    if (rgba_image)
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) {
                const sf::Color c = raycaster.render_image.getPixel(x, y);
                uint8_t* q = rgba_image + 4 * (size_t(y) * W + x);
                q[0] = c.r; q[1] = c.g; q[2] = c.b; q[3] = 255;
            }
    return code;
}

// getRand() stream (utils.cpp:77-81), for distribution tests of the 100-level lattice.
void vrt_ref_getrand(uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = getRand();
}

}  // extern "C"

// A miniature of the reference's src/main.cpp written against the drop-in headers: same class names and calls
// (SVO::setCell loop, LSVO(const SVO&), Camera, RayCaster), with the swarm lambda replaced by RayCaster::render.
// Prints known answers and an FNV-1a hash of the frame for tests/test_cpp_dropin.py to compare with the Python API.
//   usage: dropin_test <textures.bin (top 768 B + side 768 B)> [compile-only]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <vrt/vrt.hpp>

static uint64_t fnv1a(const uint8_t* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s textures.bin\n", argv[0]); return 2; }
    std::vector<uint8_t> tex(1536);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(tex.data(), 1, 1536, f) != 1536) { std::fprintf(stderr, "cannot read textures\n"); return 2; }
    std::fclose(f);
    try {
        // --- single voxel known answer: setCell(100,200,300) is seen at (411,311,211)
        {
            SVO<9> one;
            one.setCell(Cell::Solid, Cell::Grass, 100, 200, 300);
            LSVO<9> l(one);
            const float S = 512.0f;
            const HitPoint h = l.castRay(glm::vec3(1 + 411.5f / S, 1 + 311.5f / S, 1.0f), glm::vec3(0, 0, 1));
            std::printf("kat nodes=%zu hit=%d complexity=%u distance=%.8f normal=(%g,%g,%g)\n", l.data.size(), h.cell != nullptr,
                        h.complexity, h.distance, h.normal.x, h.normal.y, h.normal.z);
            std::vector<glm::vec3> o(3, glm::vec3(1 + 411.5f / S, 1 + 311.5f / S, 1.0f)), d(3, glm::vec3(0, 0, 1));
            o[1] = glm::vec3(1 + 100.5f / S, 1 + 200.5f / S, 1.0f);
            const std::vector<HitPoint> hs = l.castRays(o, d);
            std::printf("batch hits=%d%d%d\n", hs[0].cell != nullptr, hs[1].cell != nullptr, hs[2].cell != nullptr);
        }
        // --- the demo scene, built like main.cpp:59-83
        constexpr uint8_t max_depth = 9;
        constexpr int32_t size = 1 << max_depth;
        std::vector<int32_t> heights(size_t(size) * size);
        vrt::check(vrt_host_terrain_heights(size, heights.data()));      // = int32_t(64*noise.GetNoise(.75x,.75z)+32)
        SVO<max_depth>* volume_raw = new SVO<max_depth>();
        for (uint32_t x = 0; x < uint32_t(size); x++)
            for (uint32_t z = 0; z < uint32_t(size); z++) {
                const int32_t height = heights[size_t(x) * size + z];
                for (int y(1); y < std::max(16, std::min(size, height)); ++y) volume_raw->setCell(Cell::Solid, Cell::Grass, x, y + 256, z);
            }
        LSVO<max_depth> lsvo(*volume_raw);
        delete volume_raw;
        std::printf("terrain nodes=%zu\n", lsvo.data.size());

        Camera camera;
        camera.position = glm::vec3(256, 200, 256);
        camera.fov = 1.0f;
        camera.setViewAngle(glm::vec2(0.3f, -0.35f));
        std::printf("autofocus=%.6f\n", camera.autofocus(lsvo));
        camera.focal_length = 60.0f;
        camera.aperture = 0.5f;

        RayCaster raycaster(lsvo, vrt::Vector2i(256, 144), tex.data(), tex.data() + 768);
        const float scale = 1.0f / size;
        raycaster.setLightPosition(glm::vec3(-200, -1000, -300) * scale + glm::vec3(1.0f));   // main.cpp:124-126
        raycaster.use_samples = true;
        raycaster.use_gi = true;
        raycaster.render(camera, 2);
        raycaster.render(camera, 1);                       // progressive: 3 samples in total
        std::printf("frame hash=%016llx samples=%u rays=%llu\n", (unsigned long long)fnv1a(raycaster.render_image.data(), raycaster.render_image.size()),
                    raycaster.sample_count, (unsigned long long)(raycaster.last_stats.rays[0] + raycaster.last_stats.rays[1]));
        // --- the interactive loop: checkerboard halves, 0.4/0.6 temporal blend, median + persistence (main.cpp:137-172)
        RayCaster live(lsvo, vrt::Vector2i(256, 144), tex.data(), tex.data() + 768);
        live.setLightPosition(glm::vec3(-200, -1000, -300) * scale + glm::vec3(1.0f));
        live.checker_area_height = 36;                     // 144 / 4 areas (main.cpp:132)
        camera.aperture = 0.0f;
        for (int frame = 0; frame < 3; ++frame) {
            live.checker_board_offset = 1 - (frame & 1);   // main.cpp:137 flips before rendering, starting from 0
            live.render(camera);
            live.present(3);
        }
        std::printf("live hash=%016llx display=%016llx\n", (unsigned long long)fnv1a(live.render_image.data(), live.render_image.size()),
                    (unsigned long long)fnv1a(live.display.data(), live.display.size()));
        // --- Grid3D / MipmapGrid3D agree
        Grid3D<32, 32, 32>* g = new Grid3D<32, 32, 32>();
        MipmapGrid3D<32, 32, 32, 3>* m = new MipmapGrid3D<32, 32, 32, 3>();
        for (uint32_t x = 0; x < 32; ++x) for (uint32_t z = 0; z < 32; ++z) { g->setCell(Cell::Solid, x, 2, z); m->setCell(Cell::Solid, x, 2, z); }
        const HitPoint a = g->castRay(glm::vec3(5.5f, 20.2f, 7.1f), glm::vec3(0.1f, -1.0f, 0.2f));
        const HitPoint b = m->castRay(glm::vec3(5.5f, 20.2f, 7.1f), glm::vec3(0.1f, -1.0f, 0.2f));
        std::printf("grid hit=%d complexity=%u same=%d\n", a.cell != nullptr, a.complexity, a.distance == b.distance && a.complexity == b.complexity);
        delete g; delete m;
    } catch (const vrt::Error& e) {
        std::printf("vrt::Error %d: %s\n", e.code, e.what());
        return e.code == VRT_ERR_CUDA ? 3 : 1;             // 3 = no CUDA device (expected on the CPU box)
    }
    return 0;
}

// compat.hpp — stand-ins for the non-hot-path pieces the reference's main.cpp names around its per-pixel loop, so that the
// loop itself (src/main.cpp:40-158) compiles unchanged against the drop-in headers:
//
//   swrm::Swarm / WorkGroup   lib/swarm/swarm.hpp:196-215    the thread pool is replaced by kernel launches: execute() runs the
//                                                            job for every "thread" id on the calling thread (the job only queues
//                                                            rays, RayCaster::renderRay) and waitExecutionDone() launches the queue
//   FastNoise                 lib/fastnoise/FastNoise.h      SetNoiseType(SimplexFractal) + GetNoise(x, y) on libvrt's bit-exact
//                                                            host restatement of that one noise (vrt_host_noise2d)
//   sf::Vector2i, sf::Color   SFML                           only when SFML is absent (define VRT_NO_SFML_SHIM to keep them out)
//
// Nothing here is on the hot path.
#pragma once
#include <cstdint>

#include "vrt.hpp"

namespace swrm {
struct WorkGroup {
    void waitExecutionDone() { vrt::flush_all(); }      // main.cpp:156: the queued rays are shaded here, in one launch
};
class Swarm {
public:
    explicit Swarm(uint32_t thread_count) : m_thread_count(thread_count) {}
    template <typename Job>
    WorkGroup execute(Job job, uint32_t group_size = 0) {                 // swarm.hpp:215
        const uint32_t n = group_size ? group_size : m_thread_count;
        for (uint32_t id = 0; id < n; ++id) job(id, n);
        return WorkGroup();
    }
private:
    uint32_t m_thread_count;
};
}  // namespace swrm

class FastNoise {
public:
    enum NoiseType { Value, ValueFractal, Perlin, PerlinFractal, Simplex, SimplexFractal, Cellular, WhiteNoise, Cubic, CubicFractal };
    void SetNoiseType(NoiseType t) {
        if (t != SimplexFractal) throw vrt::Error(VRT_ERR_UNSUPPORTED, "FastNoise stand-in: only SimplexFractal (the demo's terrain, main.cpp:62)");
    }
    float GetNoise(float x, float y) const { return vrt_host_noise2d(x, y); }
};

#if !defined(VRT_NO_SFML_SHIM) && !defined(SFML_GRAPHICS_HPP)
namespace sf {
using Vector2i = vrt::Vector2i;
using Color = vrt::Color;
}  // namespace sf
#endif

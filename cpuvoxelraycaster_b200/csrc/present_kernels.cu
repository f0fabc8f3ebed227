// Presentation step of the reference's main loop (src/main.cpp:159-177) as one kernel: optional per-channel median
// (what res/median_3.frag and res/median.frag compute: 3x3 and 5x5 windows, vec3 min/max exchanges = a median per
// channel; edge texels repeat, the default of an sf::Texture) and the persistence blend
//   display = min(255, mul8(display, c1) + mul8(frame, c2)),  mul8(a, c) = (a * c + 127) / 255
// which is what drawing the frame, a BlendMultiply rectangle and a BlendAdd sprite leave in RGBA8 render textures.
// HBM-bound byte work: 4 B read + 4 B read + 4 B written per pixel; the window comes from a shared-memory tile.
#include "kernels.h"

namespace vrt {
namespace {

constexpr int kTileW = 32, kTileH = 8, kHalo = 2;

__device__ __forceinline__ uint32_t mul8x4(uint32_t v, uint32_t c) {       // per byte (a * c + 127) / 255, alpha forced to 255
    const uint32_t r = ((v & 0xffu) * c + 127u) / 255u;
    const uint32_t g = (((v >> 8) & 0xffu) * c + 127u) / 255u;
    const uint32_t b = (((v >> 16) & 0xffu) * c + 127u) / 255u;
    return r | (g << 8) | (b << 16);
}

// Median of N packed RGBA8 values, all four bytes at once: the largest m with #{v >= m} >= (N + 1) / 2, found bit by
// bit from the top with the byte-wise SIMD compare.
template <int N>
__device__ __forceinline__ uint32_t median_bytes(const uint32_t (&v)[N]) {
    constexpr uint32_t kNeed = uint32_t((N + 1) / 2) * 0x01010101u;
    uint32_t med = 0u;
#pragma unroll
    for (int bit = 7; bit >= 0; --bit) {
        const uint32_t cand = med | (0x01010101u << bit);
        uint32_t count = 0u;                                                  // per-byte counters, at most 25
#pragma unroll
        for (int i = 0; i < N; ++i) count += __vcmpgeu4(v[i], cand) & 0x01010101u;
        med |= (0x01010101u << bit) & __vcmpgeu4(count, kNeed);
    }
    return med;
}

template <int R>   // window radius: 0 (no filter), 1 (3x3), 2 (5x5)
__global__ void __launch_bounds__(kTileW * kTileH) present_kernel(const uint32_t* __restrict__ frame, uint32_t* __restrict__ display,
                                                                 int width, int height, uint32_t c1, uint32_t c2) {
    __shared__ uint32_t tile[kTileH + 2 * kHalo][kTileW + 2 * kHalo];
    const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
    if (R > 0) {
        for (int i = threadIdx.y * kTileW + threadIdx.x; i < (kTileH + 2 * R) * (kTileW + 2 * R); i += kTileW * kTileH) {
            const int ty = i / (kTileW + 2 * R), tx = i - ty * (kTileW + 2 * R);
            const int sx = min(max(x0 + tx - R, 0), width - 1), sy = min(max(y0 + ty - R, 0), height - 1);
            tile[ty][tx] = __ldg(frame + size_t(sy) * width + sx);
        }
        __syncthreads();
    }
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= width || y >= height) return;
    const size_t i = size_t(y) * width + x;
    uint32_t f;
    if (R == 0) {
        f = __ldg(frame + i);
    } else {
        uint32_t v[(2 * R + 1) * (2 * R + 1)];
#pragma unroll
        for (int dy = 0; dy <= 2 * R; ++dy)
#pragma unroll
            for (int dx = 0; dx <= 2 * R; ++dx) v[dy * (2 * R + 1) + dx] = tile[threadIdx.y + dy][threadIdx.x + dx];
        f = median_bytes(v);
    }
    const uint32_t keep = mul8x4(display[i], c1), add = mul8x4(f, c2);
    display[i] = __vaddus4(keep, add) | 0xff000000u;                          // saturating byte add = sf::BlendAdd
}

}  // namespace

cudaError_t launch_present(const uint8_t* d_frame, uint8_t* d_display, int width, int height, int median, uint32_t c1, uint32_t c2,
                           cudaStream_t stream) {
    if (width <= 0 || height <= 0) return cudaSuccess;
    const dim3 block(kTileW, kTileH), grid((width + kTileW - 1) / kTileW, (height + kTileH - 1) / kTileH);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(d_frame);
    uint32_t* d = reinterpret_cast<uint32_t*>(d_display);
    if (median == 3) present_kernel<1><<<grid, block, 0, stream>>>(f, d, width, height, c1, c2);
    else if (median == 5) present_kernel<2><<<grid, block, 0, stream>>>(f, d, width, height, c1, c2);
    else present_kernel<0><<<grid, block, 0, stream>>>(f, d, width, height, c1, c2);
    return cudaGetLastError();
}

}  // namespace vrt

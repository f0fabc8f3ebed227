// Host-only helpers shared by the C ABI and the scene builders (no CUDA types).
#pragma once
#include <cstdint>
#include "../../include/vrt.h"

namespace vrt {
void host_terrain_heights(int32_t size, int32_t* out);
float host_noise2d(float x, float y);   // FastNoise SimplexFractal, default settings (seed 1337, frequency 0.01, 3 octaves)
uint64_t host_build_terrain_lsvo(uint32_t depth, const int32_t* heights, vrt_lnode* out, uint64_t cap);
uint64_t host_build_lsvo_from_voxels(uint32_t depth, const uint32_t* xyz, uint64_t n_voxels, vrt_lnode* out, uint64_t cap);
void host_simplex_tables(uint8_t perm[512], uint8_t perm12[512], float* bounding);
void host_camera_rotation(const float view_angle[2], float rot_mat[9], float camera_vec[3]);
}  // namespace vrt

// gcgen kernels (sm_100a): terrain of freshly created columns written straight into the device window.
//   flat, grid, void, dungeon   (Code/Generation.cpp:772-836, :371-410, :838-852, Code/Dungeon.cpp:17-31) use no noise: reproduced exactly
//   forest, islands, snow, desert, volcanic   (:83-226, :228-369, :412-556, :558-645, :647-770) in the form the input producer gamecraft_b200/worldgen/gcworld.c restates them:
//                      the reference's FastNoiseSIMD is not vendored, so the noise is this project's own seeded simplex;
//                      the layering, tree and island fall-off rules are the reference's
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcmesh.h"

namespace gc
{
struct GenArgs
{
    uint16_t* blocks;       // the window's brick arena (cleared by the caller)
    uint8_t* surface;       // the window's surface map (cleared by the caller): filled as ComputeSurface would (forest, islands, flat)
    int16_t* height;        // scratch: terrain height per (column, z, x)
    int32_t* max_y;         // scratch: per column, the highest y the layering loop visits
    int32_t size_x, size_z;
    int32_t origin_cx, origin_cz;   // world-space column of window column (0, 0): noise is sampled in world space
    uint32_t seed;
    int32_t biome;          // 0 forest, 1 islands, 2 snow, 3 desert, 4 volcanic rules
    int32_t radius, falloff;
};

cudaError_t launch_generate_terrain(const GenArgs& a, cudaStream_t s);   // heights -> layers -> trees
cudaError_t launch_generate_flat(const GenArgs& a, cudaStream_t s);
cudaError_t launch_generate_dungeon(const GenArgs& a, cudaStream_t s);   // writes the surface map as well
cudaError_t launch_generate_grid(const GenArgs& a, cudaStream_t s);
cudaError_t launch_generate_void(const GenArgs& a, cudaStream_t s);
}  // namespace gc

// gclight kernels (sm_100a): the column pre-processing that produces the mesher's light inputs on the device —
// surface map (ComputeSurface, Code/Lighting.cpp:5-26) and the sunlight / blockLight flood fill (SetLightNodes,
// ScatterSunlight, ScatterBlockLight, Lighting.cpp:28-281) as the least fixpoint of the reference's push rule.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcmesh.h"

namespace gc
{
struct LightRules
{
    uint8_t step[256];      // lightStep (Block.h:95); 255 = opaque (INT_MAX, Block.cpp:65-68)
    uint8_t emitted[256];   // lightEmitted
};

struct LightWorld
{
    const uint16_t* blocks;
    uint8_t* sun;
    uint8_t* light;
    uint8_t* surface;
    int32_t size_x, size_z;
    const LightRules* rules;
};

// counters written by the scan kernels (device memory, kLightCounterWords x u32); from LC_FLAGS on: the relaxation sweeps'
// "something changed" words, one preset word + one per sweep for sunlight, then the same for blockLight
enum { LC_SUN_CAND = 0, LC_LIGHT_CAND = 1, LC_ACTIVE_BRICKS = 2, LC_SEEDS = 4, LC_FLAGS = 8 };
constexpr int kLightSweeps = 14;
constexpr int kLightCounterWords = LC_FLAGS + 2 * (kLightSweeps + 1);

// brick_top: scratch, 1024 bytes per brick of the window
cudaError_t launch_surface(const LightWorld& w, uint8_t* brick_top, cudaStream_t s);
// pass over every column cell up to its maxY: appends the sunlight candidates (non-opaque cells below their column's
// surface; entries beyond `cap` are only counted), seeds the emitters and flags their bricks
cudaError_t launch_light_scan(const LightWorld& w, uint32_t* sun_list, uint32_t cap, uint8_t* brick_flags, uint32_t* counters, cudaStream_t s);
// bricks within one brick of a flagged brick -> active list
cudaError_t launch_active_bricks(const LightWorld& w, const uint8_t* brick_flags, uint32_t* active_list, uint32_t* counters, cudaStream_t s);
// non-opaque cells of the active bricks: append (entries beyond `cap` are only counted)
cudaError_t launch_light_candidates(const LightWorld& w, const uint32_t* active_list, int n_active, uint32_t* light_list, uint32_t cap, uint32_t* counters, cudaStream_t s);
// one in-place relaxation sweep over a candidate list (sun = true: sunlight rules with the sky sources); `changed`: this
// sweep's flag word, the previous sweep's at changed[-1]
cudaError_t launch_relax(const LightWorld& w, bool sun, const uint32_t* list, uint32_t n, uint32_t* changed, cudaStream_t s);
// one sweep of both fills in one launch; the lists' lengths are read from device memory, `cap` is their capacity
cudaError_t launch_relax_pair(const LightWorld& w, const uint32_t* sun_list, const uint32_t* light_list, uint32_t cap, const uint32_t* n_sun,
                              const uint32_t* n_light, uint32_t* changed_sun, uint32_t* changed_light, cudaStream_t s);
}  // namespace gc

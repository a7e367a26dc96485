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

// counters written by the scan kernels (device memory, 8 x u32)
enum { LC_SUN_CAND = 0, LC_LIGHT_CAND = 1, LC_ACTIVE_BRICKS = 2, LC_CHANGED = 3, LC_SEEDS = 4 };

cudaError_t launch_surface(const LightWorld& w, cudaStream_t s);
// pass over every column cell up to its maxY: counts (fill = false) or appends (fill = true) the sunlight candidates
// (non-opaque cells below their column's surface); the filling pass also seeds the emitters and flags their bricks
cudaError_t launch_light_scan(const LightWorld& w, bool fill, uint32_t* sun_list, uint8_t* brick_flags, uint32_t* counters, cudaStream_t s);
// bricks within one brick of a flagged brick -> active list
cudaError_t launch_active_bricks(const LightWorld& w, const uint8_t* brick_flags, uint32_t* active_list, uint32_t* counters, cudaStream_t s);
// non-opaque cells of the active bricks: count / append
cudaError_t launch_light_candidates(const LightWorld& w, const uint32_t* active_list, int n_active, bool fill, uint32_t* light_list, uint32_t* counters, cudaStream_t s);
// one in-place relaxation sweep over a candidate list (sun = true: sunlight rules with the sky sources)
cudaError_t launch_relax(const LightWorld& w, bool sun, const uint32_t* list, uint32_t n, uint32_t* counters, cudaStream_t s);
}  // namespace gc

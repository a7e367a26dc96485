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

// statistics of a fill (device memory, kLightCounterWords x u32, cleared by the caller); from LC_FLAGS on: the edit path's
// "something changed" words of its list sweeps, one preset word + one per sweep for sunlight, then the same for blockLight
constexpr int kLightSweeps = 14;
enum { LC_SUN_CAND = 0, LC_SWEEPS = 1, LC_SEEDS = 4, LC_FLAGS = 8, LC_VISITS = LC_FLAGS + 2 * (kLightSweeps + 1) };   // LC_VISITS: brick visits per sweep
constexpr int kLightCounterWords = LC_VISITS + kLightSweeps;

// brick_top: scratch, 1024 bytes per brick of the window
cudaError_t launch_surface(const LightWorld& w, uint8_t* brick_top, cudaStream_t s);
// pass over every column cell up to its maxY: the candidate bit maps (one word per x-row of a brick, 2048 words per brick;
// cleared by the caller), per brick "holds a sunlight candidate" / "holds a seeded emitter", per unit (256 rows; 8 per brick)
// "holds a sunlight candidate", emitters seeded
cudaError_t launch_light_scan(const LightWorld& w, uint32_t* sun_bits, uint32_t* opq_bits, uint8_t* sun_flag, uint8_t* sun_unit, uint8_t* seed_flag, uint32_t* counters, cudaStream_t s);
// the relaxation of both fills, all sweeps in one cooperative launch: sweep k visits the bricks that, or whose face neighbours,
// changed in sweep k - 1 (flag arrays [k] -> [k + 1], `flag_stride` bytes apart, cleared by the caller except [0] from the scan);
// lists: 2 x (2 x 8 x bricks) words of scratch (visits are units of 256 rows); row_maps: 3 x (2 x 64 x bricks) words — per sweep
// and fill one bit per row "changed" (map 1 cleared by the caller, the kernel clears the others as it goes); ends at the first
// sweep that has nothing to visit
cudaError_t launch_relax_all(const LightWorld& w, const uint32_t* sun_bits, const uint32_t* opq_bits, const uint8_t* sun_unit, uint8_t* sun_flags, uint8_t* light_flags,
                             size_t flag_stride, uint32_t* lists, uint32_t* row_maps, uint32_t* counters, int num_sms, int ctas_per_sm, cudaStream_t s);
// the edit path's sweep over its two candidate lists (gcedit_kernel.cu): both fills in one launch; the lists' lengths are read
// from device memory, `cap` is their capacity; changed_*: this sweep's flag word, the previous sweep's at changed_*[-1]
cudaError_t launch_relax_pair(const LightWorld& w, const uint32_t* sun_list, const uint32_t* light_list, uint32_t cap, const uint32_t* n_sun,
                              const uint32_t* n_light, uint32_t* changed_sun, uint32_t* changed_light, cudaStream_t s);
}  // namespace gc

// gcedit kernels (sm_100a): the interactive caller of the mesher — SetBlock (Code/World.cpp:548-582) with its
// ChunkOverflowed pre-check (WorldRender.cpp:412-439), the relight around the edit (RecomputeLight,
// Lighting.cpp:283-449) and the list of chunks to rebuild (FlagChunkForUpdate, World.cpp:507-546) — on a window
// whose arrays live on the device.  Nothing crosses PCIe but the edit, its status and the rebuild list.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcmesh.h"
#include "gclight_kernel.cuh"
#include "gcmesh_kernel.cuh"

namespace gc
{
// Light travels fewer than 15 voxels and an edit changes the sky of one column only, so every cell whose light can
// change lies in the full-height box of this radius around the edited column (one more than the reach, so the
// outermost cells that can change still have their fixed neighbours inside the window arrays).
constexpr int kEditRadius = 16;
constexpr int kEditSpan = 2 * kEditRadius + 1;
constexpr uint32_t kEditBoxCells = (uint32_t)kEditSpan * kEditSpan * GC_WORLD_BLOCK_HEIGHT;

struct EditBox
{
    int32_t x0, z0;       // window voxel coordinates of the box's low corner (may lie outside the window: cells are clipped)
    int32_t nx, nz;       // kEditSpan each
};

// device words of one edit's progress; the relaxation flag words follow the layout of LC_FLAGS
enum { ES_SUN_CAND = 0, ES_LIGHT_CAND = 1, ES_SKIP = 2, ES_FLAGS = 8 };
constexpr int kEditStateWords = ES_FLAGS + 2 * (kLightSweeps + 1);

struct EditArgs
{
    LightWorld w;                 // blocks are written by the apply kernel
    uint16_t* blocks_rw;
    const DeviceTables* tables;   // cls8: class byte per block id (cull class = popcount of its low four bits); kill_zone
    const int32_t* vert_table;    // per brick: vertices of its last build, -1 = never built
    const GcBlockEdit* edits;     // device copy of the caller's list
    int32_t* status;              // per edit: GC_EDIT_*
    uint8_t* dirty;               // per brick: 1 = rebuild
    uint32_t* state;              // kEditStateWords
    uint8_t* save;                // 2 x kEditBoxCells: effective sunlight, blockLight before the edit
    uint32_t* sun_list;           // kEditBoxCells each
    uint32_t* light_list;
};

// the five kernels of one edit, enqueued back to back on `s` (no host round trip in between)
cudaError_t launch_edit(const EditArgs& a, int edit_index, const EditBox& box, cudaStream_t s);
// dirty bricks of the pre-processed (non-edge) columns -> coords (unordered), count in *n_out; clears the flags
cudaError_t launch_edit_list(const LightWorld& w, uint8_t* dirty, GcChunkCoord* out, uint32_t* n_out, cudaStream_t s);
}  // namespace gc

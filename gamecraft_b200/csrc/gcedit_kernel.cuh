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
#include "relight_core.cuh"

namespace gc
{
// Light travels fewer than 15 voxels and an edit changes the sky of one column only, so every cell whose light can
// change lies in the full-height box of this radius around the edited column (one more than the reach, so the
// outermost cells that can change still have their fixed neighbours inside the window arrays).
constexpr int kEditRadius = 16;
constexpr int kEditSpan = 2 * kEditRadius + 1;
constexpr uint32_t kEditBoxCells = (uint32_t)kEditSpan * kEditSpan * GC_WORLD_BLOCK_HEIGHT;

// device words of one edit's progress (cleared before each edit); the relaxation flag words follow the layout of LC_FLAGS
enum { ES_SUN_CAND = 0, ES_LIGHT_CAND = 1, ES_SKIP = 2, ES_FLAGS = 8 };
constexpr int kEditStateWords = ES_FLAGS + 2 * (kLightSweeps + 1);
// ... followed by two words that live across the edits of one call: the rebuild count and the index of the current edit
enum { EP_REBUILD_COUNT = kEditStateWords, EP_CURSOR = kEditStateWords + 1, kEditWords = kEditStateWords + 2 };

struct EditArgs
{
    LightWorld w;                 // blocks are written by the apply kernel
    uint16_t* blocks_rw;
    const DeviceTables* tables;   // cls8: class byte per block id (cull class = popcount of its low four bits); kill_zone
    const int32_t* vert_table;    // per brick: vertices of its last build, -1 = never built
    const GcBlockEdit* edits;     // device copy of the caller's list
    int32_t* status;              // per edit: GC_EDIT_*
    uint8_t* dirty;               // per brick: 1 = rebuild
    uint32_t* state;              // kEditWords
    uint8_t* save;                // 2 x kEditBoxCells: effective sunlight, blockLight before the edit
    uint32_t* sun_list;           // kEditBoxCells each
    uint32_t* light_list;
    // GC_WORLD_EDIT_REPLAY: the reference's own queues replayed node for node (relight_core.cuh) instead of the box refill
    int32_t replay;
    uint64_t* queue_a;            // kRelightQueueNodes entries each
    uint64_t* queue_b;
};

// The kernels of one edit — the one state[EP_CURSOR] points at, which the last kernel advances — enqueued back to back
// on `s`.  Nothing in the argument list depends on the edit, so the sequence can be captured once into a CUDA graph and
// replayed per edit.
cudaError_t launch_edit(const EditArgs& a, cudaStream_t s);
// The same edit with RecomputeLight (Lighting.cpp:283-449) replayed as the reference runs it: apply (decisions, store, overflow,
// the 27 flags — the surface entry is left to the replay, which needs the old one) -> one thread walking the queues -> cursor step.
cudaError_t launch_edit_replay(const EditArgs& a, cudaStream_t s);
// dirty bricks of the pre-processed (non-edge) columns -> coords (unordered), count in *n_out; clears the flags
cudaError_t launch_edit_list(const LightWorld& w, uint8_t* dirty, GcChunkCoord* out, uint32_t* n_out, cudaStream_t s);
}  // namespace gc

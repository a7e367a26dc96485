// gcmesh kernels (sm_100a): declarations shared by the kernel TU and the C-ABI TU.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gcmesh.h"
#include "mesh_core.cuh"
#include "gchalo_kernel.cuh"

namespace gc
{
// Device-side block tables, built by gcmesh_set_block_table.
struct DeviceTables
{
    LutEntry lut[256];   // class bits spread for the bit-plane accumulation
    MatEntry mat[256];   // texture layers + alpha
    uint8_t cls8[256];   // class byte
    int32_t kill_zone;
};

struct MeshParams
{
    const uint16_t* blocks;
    const uint8_t* sun;
    const uint8_t* light;
    const uint8_t* surface;
    int32_t size_x, size_z;
    const GcChunkCoord* coords;
    const int32_t* order;              // optional: the k-th claimed work item is chunk order[k] (heaviest first), null = k
    int32_t n_chunks;
    GcChunkOut* out;
    int32_t* vert_table;               // optional, per brick of the window: vertices of its last build (chunk->totalVertices), -1 = never built
    uint8_t* vertex_arena;
    uint16_t* index_arena;
    unsigned long long max_quads;
    unsigned long long* quad_cursor;   // arena reservation cursor (quads)
    int32_t* work_counter;             // dynamic chunk scheduler
    uint32_t* done_counter;            // CTAs that have left the work loop; the last one resets the scheduler for the next launch
    unsigned long long* total_out;     // ... and leaves the quad cursor's final value here
    unsigned long long* total_out_host; // optional: and in mapped host memory (gcmesh_mesh_host: a copy on the stream would queue behind the read-backs)
    int32_t keep_cursor;               // 1: the cursor keeps running into the next launch (strips of one gcmesh_mesh_host call)
    const DeviceTables* tables;
    int32_t* error_flag;               // set by the mbarrier watchdog
    int32_t use_tma;                   // 0: the producer warp stages slabs with plain loads (debug aid)
    unsigned long long* prof;          // optional: 512 cycle counters (phase / wait attribution), null = off
    // slab halo exchange fused into this kernel (gcmesh_submit_exchange): work items [0, n_push) store this rank's boundary
    // planes into the neighbours' windows, item n_push + k is chunk order[k]
    int32_t n_push;
    HaloFused halo;
    int32_t trust_surface;             // GC_WORLD_SURFACE_BOUNDS_MESH: bricks at or above their column's highest surface entry are AIR, not read
    int32_t wide;                      // uncapped batch (GC_BATCH_UNCAPPED): no quad cap per chunk, 32-bit indices (24 bytes per quad in the index arena)
};

// The block-id arena viewed as (256, 8, z = 32, slot): one box (256, 8, 1, 1) is one whole z-slab of a brick
// (32 x 64 voxels, 4 KB contiguous).  The light arenas need no map: light is gathered per quad, never staged.
struct MeshTensorMaps
{
    CUtensorMap blocks_slab;
};

#ifndef GCMESH_THREADS
#define GCMESH_THREADS 448
#endif
constexpr int kMeshThreads = GCMESH_THREADS;   // two CTAs per SM
size_t mesh_smem_bytes();
cudaError_t launch_mesh(const MeshParams& p, const MeshTensorMaps& maps, int grid, cudaStream_t stream);
cudaError_t mesh_kernel_attributes(int* regs, int* static_smem, int* max_dyn_smem);
int mesh_ctas_per_sm();   // resident mesh CTAs per SM (occupancy query), 0 on error

// Sparse upload: copy the listed brick arrays from (pinned, device-mapped) host memory into their device slots.
// item = slot * 4 + array (0 blocks, 1 sunlight, 2 blockLight); host pointers are the brick-major arrays.
cudaError_t launch_upload_items(const uint32_t* items, int n_items, const uint16_t* h_blocks, const uint8_t* h_sun, const uint8_t* h_light,
                                uint16_t* d_blocks, uint8_t* d_sun, uint8_t* d_light, cudaStream_t stream);

// Slab-halo pack / unpack (multi-GPU): see gcmesh_world_pack_halo in gcmesh.h.
cudaError_t launch_halo_copy(const uint16_t* blocks, const uint8_t* sun, const uint8_t* light, const uint8_t* surface,
                             int size_x, int cz, int z, uint8_t* buf, bool unpack, cudaStream_t stream);
size_t halo_bytes(int size_x);
}  // namespace gc

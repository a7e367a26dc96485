// gcmesh relight core: the reference's incremental relight after one block edit — RecomputeLight (Code/Lighting.cpp:283-449)
// with ScatterSunlight / RemoveSunlightNodes / ScatterBlockLight / RemoveLightNodes (:109-229, :344-388) and ComputeSurfaceAt
// (:5-17) — replayed node for node: the same FIFO queues, the same neighbour order (DIRS_3, Utils.h:106-111: up, down, front,
// back, left, right), the same comparisons, so that the arrays end up byte for byte as the reference leaves them, including
// where its removal queues stop short of a from-scratch fill and the surface entry it leaves stale when a column empties.
// GC_WORLD_EDIT_REPLAY selects it in gcmesh_world_set_blocks; the default relight is the order-independent fill (gcedit_kernel.cu).
//
// Small __host__ __device__ functions over plain pointers (like mesh_core.cuh): the device kernel runs them on one thread — an
// edit touches a few thousand nodes and the replay is sequential by definition — and tests/emu runs the same code on the host
// against the reference compiled verbatim.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GC_RL_HD __host__ __device__ __forceinline__
#else
#define GC_RL_HD static inline
#endif

namespace gc
{
constexpr int kRelightQueueNodes = 262144;   // MAX_LIGHT_NODES, Lighting.h:8 (a power of two: the queue is a ring)

// a window of columns as the device holds it (brick-major, see gcmesh.h)
struct RelightWorld
{
    const uint16_t* blocks;
    uint8_t* sun;
    uint8_t* light;
    uint8_t* surface;
    const uint8_t* step;        // lightStep per block id, 255 = opaque (INT_MAX in the reference)
    const uint8_t* emitted;     // lightEmitted per block id
    const int32_t* vert_table;  // per brick: vertices of its last build, < 0 = never built (chunk->state < CHUNK_BUILDING)
    uint8_t* dirty;             // per brick: pendingUpdate
    int32_t size_x, size_z;
};

// Queue<ivec3> (Containers.h:52-94): a ring of window voxel positions packed x | z << 20 | y << 40
struct RelightQueue
{
    uint64_t* items;
    int32_t read, write, size;
};

GC_RL_HD uint64_t rl_pack(int x, int y, int z) { return (uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)z << 20) | ((uint64_t)(uint32_t)y << 40); }
GC_RL_HD void rl_unpack(uint64_t n, int& x, int& y, int& z) { x = (int)(n & 0xFFFFFu); z = (int)((n >> 20) & 0xFFFFFu); y = (int)(n >> 40); }
GC_RL_HD void rl_clear(RelightQueue& q) { q.read = 0; q.write = 0; q.size = 0; }
GC_RL_HD bool rl_empty(const RelightQueue& q) { return q.size == 0; }
GC_RL_HD void rl_push(RelightQueue& q, int x, int y, int z)
{
    q.items[q.write] = rl_pack(x, y, z);
    q.write = (q.write + 1) & (kRelightQueueNodes - 1);
    q.size++;
}
GC_RL_HD void rl_pop(RelightQueue& q, int& x, int& y, int& z)
{
    rl_unpack(q.items[q.read], x, y, z);
    q.read = (q.read + 1) & (kRelightQueueNodes - 1);
    q.size--;
}

GC_RL_HD size_t rl_voxel(const RelightWorld& w, int x, int y, int z)
{
    return (((size_t)(z >> 5) * w.size_x + (x >> 5)) * 4 + (y >> 6)) * 65536 + (size_t)((x & 31) + 32 * ((y & 63) + 64 * (z & 31)));
}
GC_RL_HD size_t rl_surface_index(const RelightWorld& w, int x, int z) { return ((size_t)(z >> 5) * w.size_x + (x >> 5)) * 1024 + (size_t)((z & 31) * 32 + (x & 31)); }
GC_RL_HD bool rl_opaque(const RelightWorld& w, uint32_t block) { return w.step[block & 255u] == 255; }
// GetSunlight (Lighting.cpp:69-79)
GC_RL_HD int rl_sunlight(const RelightWorld& w, int x, int y, int z, size_t v)
{
    if (y >= (int)w.surface[rl_surface_index(w, x, z)]) return 15;
    const int s = w.sun[v];
    return s == 0 ? 1 : s;
}
// next->state >= CHUNK_BUILDING -> next->pendingUpdate = true
GC_RL_HD void rl_touch(const RelightWorld& w, size_t v)
{
    const size_t slot = v >> 16;
    if (w.vert_table[slot] >= 0) w.dirty[slot] = 1;
}

GC_RL_HD void rl_neighbour(int d, int x, int y, int z, int& nx, int& ny, int& nz)
{
    // DIRS_3 (Utils.h:106-111): up, down, front (+z), back (-z), left (-x), right (+x)
    nx = x + (d == 5 ? 1 : (d == 4 ? -1 : 0));
    ny = y + (d == 0 ? 1 : (d == 1 ? -1 : 0));
    nz = z + (d == 2 ? 1 : (d == 3 ? -1 : 0));
}

// ScatterSunlight (Lighting.cpp:109-143)
GC_RL_HD void rl_scatter_sun(const RelightWorld& w, RelightQueue& q, bool update)
{
    while (!rl_empty(q))
    {
        int x, y, z;
        rl_pop(q, x, y, z);
        const size_t v = rl_voxel(w, x, y, z);
        const uint32_t block = w.blocks[v];
        if (rl_opaque(w, block)) continue;                               // value - INT_MAX <= MIN_LIGHT
        const int light = rl_sunlight(w, x, y, z, v) - (int)w.step[block & 255u];
        if (light <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            int nx, ny, nz;
            rl_neighbour(d, x, y, z, nx, ny, nz);
            if (ny < 0 || ny >= 256) continue;
            const size_t nv = rl_voxel(w, nx, ny, nz);
            if (update) rl_touch(w, nv);
            if (!rl_opaque(w, w.blocks[nv]) && rl_sunlight(w, nx, ny, nz, nv) < light)      // SetMaxSunlight (:81-93)
            {
                w.sun[nv] = (uint8_t)light;
                rl_push(q, nx, ny, nz);
            }
        }
    }
}

// RemoveSunlightNodes (Lighting.cpp:145-193); `fresh` is the queue it builds for the scatter that follows
GC_RL_HD void rl_remove_sun(const RelightWorld& w, RelightQueue& q, RelightQueue& fresh)
{
    rl_clear(fresh);
    while (!rl_empty(q))
    {
        int x, y, z;
        rl_pop(q, x, y, z);
        if (y >= (int)w.surface[rl_surface_index(w, x, z)]) { rl_push(fresh, x, y, z); continue; }
        const size_t v = rl_voxel(w, x, y, z);
        const int light = rl_sunlight(w, x, y, z, v) - 1;
        w.sun[v] = 1;
        if (light <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            int nx, ny, nz;
            rl_neighbour(d, x, y, z, nx, ny, nz);
            if (ny < 0 || ny >= 256) continue;
            const size_t nv = rl_voxel(w, nx, ny, nz);
            if (!rl_opaque(w, w.blocks[nv]))
            {
                if (rl_sunlight(w, nx, ny, nz, nv) <= light) rl_push(q, nx, ny, nz);
                else rl_push(fresh, nx, ny, nz);
            }
            rl_touch(w, nv);
        }
    }
    rl_scatter_sun(w, fresh, true);
}

// ScatterBlockLight (Lighting.cpp:195-229)
GC_RL_HD void rl_scatter_block(const RelightWorld& w, RelightQueue& q, bool update)
{
    while (!rl_empty(q))
    {
        int x, y, z;
        rl_pop(q, x, y, z);
        const size_t v = rl_voxel(w, x, y, z);
        const uint32_t block = w.blocks[v];
        if (rl_opaque(w, block)) continue;
        const int light = (int)w.light[v] - (int)w.step[block & 255u];
        if (light <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            int nx, ny, nz;
            rl_neighbour(d, x, y, z, nx, ny, nz);
            if (ny < 0 || ny >= 256) continue;
            const size_t nv = rl_voxel(w, nx, ny, nz);
            if (update) rl_touch(w, nv);
            if (!rl_opaque(w, w.blocks[nv]) && (int)w.light[nv] < light)                    // SetMaxBlockLight (:95-107)
            {
                w.light[nv] = (uint8_t)light;
                rl_push(q, nx, ny, nz);
            }
        }
    }
}

// RemoveLightNodes (Lighting.cpp:344-388)
GC_RL_HD void rl_remove_block(const RelightWorld& w, RelightQueue& q, RelightQueue& fresh)
{
    rl_clear(fresh);
    while (!rl_empty(q))
    {
        int x, y, z;
        rl_pop(q, x, y, z);
        const size_t v = rl_voxel(w, x, y, z);
        const int light = (int)w.light[v] - 1;
        w.light[v] = 0;
        if (light <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            int nx, ny, nz;
            rl_neighbour(d, x, y, z, nx, ny, nz);
            if (ny < 0 || ny >= 256) continue;
            const size_t nv = rl_voxel(w, nx, ny, nz);
            const uint32_t adj = w.blocks[nv];
            if (!rl_opaque(w, adj))
            {
                if ((int)w.light[nv] <= light) rl_push(q, nx, ny, nz);
                else rl_push(fresh, nx, ny, nz);
            }
            if (w.emitted[adj & 255u] > 1) rl_push(fresh, nx, ny, nz);
            rl_touch(w, nv);
        }
    }
    rl_scatter_block(w, fresh, true);
}

// the six neighbours of the edited cell, then a scatter (UpdateSunlight :283-296, UpdateBlockLight :390-403)
GC_RL_HD void rl_push_neighbours(RelightQueue& q, int x, int y, int z)
{
    for (int d = 0; d < 6; d++)
    {
        int nx, ny, nz;
        rl_neighbour(d, x, y, z, nx, ny, nz);
        if (ny < 0 || ny >= 256) continue;
        rl_push(q, nx, ny, nz);
    }
}

// RecomputeLight (Lighting.cpp:441-449) for the cell (x, y, z) whose block has just been stored.  qa, qb: two queues of
// kRelightQueueNodes entries each.
GC_RL_HD void rl_recompute_light(const RelightWorld& w, int x, int y, int z, RelightQueue& qa, RelightQueue& qb)
{
    // ---- RecomputeSunlight (:298-342) ----
    rl_clear(qa);
    const size_t si = rl_surface_index(w, x, z);
    const int old_surface = w.surface[si];
    // ComputeSurfaceAt (:5-17): the highest block in 0 .. 254, plus one — and no store at all when the column holds none
    for (int yy = 254; yy >= 0; yy--)
        if (w.blocks[rl_voxel(w, x, yy, z)] != 0) { w.surface[si] = (uint8_t)(yy + 1); break; }
    const int new_surface = w.surface[si];
    if (new_surface < old_surface)
    {
        for (int yy = new_surface; yy <= old_surface; yy++) rl_push(qa, x, yy, z);
        rl_scatter_sun(w, qa, true);
    }
    else if (new_surface > old_surface)
    {
        for (int yy = old_surface; yy <= new_surface; yy++)
        {
            w.sun[rl_voxel(w, x, yy, z)] = 15;
            rl_push(qa, x, yy, z);
        }
        rl_remove_sun(w, qa, qb);
    }
    else if (rl_opaque(w, w.blocks[rl_voxel(w, x, y, z)]))
    {
        rl_push(qa, x, y, z);
        rl_remove_sun(w, qa, qb);
    }
    else
    {
        rl_push_neighbours(qa, x, y, z);
        rl_scatter_sun(w, qa, true);
    }
    // ---- RecomputeBlockLight (:405-439) ----
    rl_clear(qa);
    const size_t v = rl_voxel(w, x, y, z);
    const int old_light = w.light[v];
    const int emission = w.emitted[w.blocks[v] & 255u];
    if (emission < old_light)
    {
        w.light[v] = 15;
        rl_push(qa, x, y, z);
        rl_remove_block(w, qa, qb);
    }
    if (emission > 1)
    {
        w.light[v] = (uint8_t)emission;
        rl_push(qa, x, y, z);
        rl_scatter_block(w, qa, true);
    }
    else
    {
        rl_push_neighbours(qa, x, y, z);
        rl_scatter_block(w, qa, true);
    }
}
}  // namespace gc

// gcedit — one block edit on a device-resident window (see gcedit_kernel.cuh).
//
// Five steps per edit, each a small kernel over the full-height box of radius kEditRadius around the edited column
// (33 x 33 x 256 cells, ~280 k threads: launch-latency bound, no host round trip in between).  The kernels find the edit
// through a device-side cursor, so their arguments never change and the host replays the sequence as one CUDA graph:
//   save     effective sunlight (GetSunlight, Lighting.cpp:69-79: 15 at or above the surface, a stored 0 reads 1) and
//            blockLight of every cell, as the mesher would read them before the edit
//   apply    one warp: the decisions of SetBlock (World.cpp:548-582), the ChunkOverflowed count
//            (WorldRender.cpp:412-439), the store, the column's surface entry (ComputeSurfaceAt, Lighting.cpp:5-17),
//            the 27 chunks of FlagChunkForUpdate (World.cpp:507-546)
//   prepare  resets the box to the state a from-scratch fill starts from (zero, emitters seeded) and lists its
//            candidate cells
//   relax    the fill's own sweeps (relax_cell, both fills per launch) over those lists; cells outside the box are read, never written:
//            they hold the fixpoint already and light does not reach further than the box
//   diff     every cell whose block light or effective sunlight differs from the saved value flags the chunks that
//            have it inside their one-voxel halo
// A refused edit sets the skip word: the later kernels of that edit return at once.
#include "gcedit_kernel.cuh"

namespace gc
{
namespace
{
constexpr int kVox = GC_CHUNK_SIZE_3;

struct Cell
{
    bool inside;
    int x, y, z;          // window voxel coordinates
    int col;
    uint32_t v;           // index into the brick arenas
};

// cell i of the box around the current edit (x fastest, then y, then z); cells outside the window are skipped
__device__ __forceinline__ Cell box_cell(const EditArgs& a, uint32_t i)
{
    const LightWorld& w = a.w;
    const GcBlockEdit e = a.edits[a.state[EP_CURSOR]];
    Cell c;
    const int lx = (int)(i % (uint32_t)kEditSpan);
    const uint32_t r = i / (uint32_t)kEditSpan;
    c.y = (int)(r % GC_WORLD_BLOCK_HEIGHT);
    const int lz = (int)(r / GC_WORLD_BLOCK_HEIGHT);
    c.x = e.wx - kEditRadius + lx; c.z = e.wz - kEditRadius + lz;
    c.inside = c.x >= 0 && c.z >= 0 && c.x < w.size_x * 32 && c.z < w.size_z * 32;
    c.col = 0; c.v = 0;
    if (c.inside)
    {
        c.col = (c.z >> 5) * w.size_x + (c.x >> 5);
        c.v = (uint32_t)(((size_t)c.col * 4 + (c.y >> 6)) * kVox + (c.x & 31) + 32 * ((c.y & 63) + 64 * (c.z & 31)));
    }
    return c;
}

__device__ __forceinline__ int surface_at(const LightWorld& w, int x, int z)
{
    return w.surface[(size_t)((z >> 5) * w.size_x + (x >> 5)) * 1024 + (z & 31) * 32 + (x & 31)];
}

__device__ __forceinline__ uint32_t effective_sun(const LightWorld& w, const Cell& c)
{
    if (c.y >= surface_at(w, c.x, c.z)) return 15u;
    const uint32_t s = w.sun[c.v];
    return s == 0 ? 1u : s;
}

__device__ __forceinline__ bool inner_column(const LightWorld& w, int x, int z)
{
    const int cx = x >> 5, cz = z >> 5;
    return cx >= 1 && cx < w.size_x - 1 && cz >= 1 && cz < w.size_z - 1;
}

// the chunks that have window voxel (x, y, z) inside their one-voxel halo
__device__ void flag_around(const LightWorld& w, uint8_t* dirty, int x, int y, int z)
{
    const int cx0 = max((x - 1) >> 5, 0), cx1 = min((x + 1) >> 5, w.size_x - 1);
    const int cz0 = max((z - 1) >> 5, 0), cz1 = min((z + 1) >> 5, w.size_z - 1);
    const int cy0 = max((y - 1) >> 6, 0), cy1 = min((y + 1) >> 6, 3);
    for (int cz = cz0; cz <= cz1; cz++)
        for (int cx = cx0; cx <= cx1; cx++)
            for (int cy = cy0; cy <= cy1; cy++)
                dirty[((size_t)cz * w.size_x + cx) * 4 + cy] = 1;
}

__global__ void __launch_bounds__(256) gc_edit_save_kernel(EditArgs a)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kEditBoxCells) return;
    const Cell c = box_cell(a, i);
    if (!c.inside) return;
    a.save[i] = (uint8_t)effective_sun(a.w, c);
    a.save[kEditBoxCells + i] = a.w.light[c.v];
}

__device__ __forceinline__ int cull_class(const EditArgs& a, uint32_t block) { return __popc(a.tables->cls8[block & 255u] & 15u); }

// block at window voxel (x, y, z) as GetBlockSafe reads a neighbour (World.cpp:422-431): KILL_ZONE below the world, AIR above
__device__ __forceinline__ uint32_t block_safe(const EditArgs& a, int x, int y, int z)
{
    if (y < 0) return (uint32_t)a.tables->kill_zone;
    if (y >= GC_WORLD_BLOCK_HEIGHT) return 0u;
    const int col = (z >> 5) * a.w.size_x + (x >> 5);
    return a.blocks_rw[((size_t)col * 4 + (y >> 6)) * kVox + (x & 31) + 32 * ((y & 63) + 64 * (z & 31))];
}

__global__ void __launch_bounds__(32) gc_edit_apply_kernel(EditArgs a)
{
    const int lane = threadIdx.x;
    const uint32_t edit_index = a.state[EP_CURSOR];
    const GcBlockEdit e = a.edits[edit_index];
    const int x = e.wx, y = e.wy, z = e.wz;
    int status = GC_EDIT_APPLIED;
    const int col = (z >> 5) * a.w.size_x + (x >> 5);
    size_t slot = 0, index = 0;
    if (y < 0 || y >= GC_WORLD_BLOCK_HEIGHT) status = GC_EDIT_IGNORED;                      // World.cpp:550
    else
    {
        slot = (size_t)col * 4 + (y >> 6);
        index = slot * kVox + (x & 31) + 32 * ((y & 63) + 64 * (z & 31));
        // lane 0 alone reads the old block (it is about to overwrite it) and the warp takes its verdict: every lane leaves
        // this block with the same status whatever the order of its load and lane 0's store
        if (lane == 0)
        {
            if (a.vert_table[slot] < 0) status = GC_EDIT_NOT_BUILT;                        // World.cpp:562-565
            else if (a.blocks_rw[index] == e.block) status = GC_EDIT_UNCHANGED;            // World.cpp:570
            else a.blocks_rw[index] = e.block;
        }
    }
    status = __shfl_sync(0xffffffffu, status, 0);
    if (status == GC_EDIT_APPLIED)
    {
        __syncwarp();
        if (e.block != 0)
        {
            // ChunkOverflowed (WorldRender.cpp:412-439): four vertices per face of the new block that CanDrawFace lets through
            const int cull = cull_class(a, e.block);
            int added = 0;
            const int dx[6] = { 0, 0, 0, 0, 1, -1 }, dy[6] = { 1, -1, 0, 0, 0, 0 }, dz[6] = { 0, 0, 1, -1, 0, 0 };
            for (int f = 0; f < 6; f++)
            {
                const uint32_t adj = block_safe(a, x + dx[f], y + dy[f], z + dz[f]);
                if (cull < cull_class(a, adj) && adj != e.block) added += 4;
            }
            if (a.vert_table[slot] + added > GC_MAX_VERTICES)
            {
                if (lane == 0) a.blocks_rw[index] = 0;                                      // World.cpp:573
                status = GC_EDIT_OVERFLOW;
            }
        }
    }
    if (status == GC_EDIT_APPLIED)
    {
        __syncwarp();
        // ComputeSurfaceAt (Lighting.cpp:5-17): highest non-AIR y in 0..254, plus one; the warp looks at all 255 cells at once
        // (replay mode: the replay computes it itself — RecomputeSunlight compares the old entry with the new one)
        int top = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const int yy = lane + 32 * k;
            if (yy <= GC_WORLD_BLOCK_HEIGHT - 2 && block_safe(a, x, yy, z) != 0) top = yy + 1;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, d));
        if (lane == 0)
        {
            if (!a.replay) a.w.surface[(size_t)col * 1024 + (z & 31) * 32 + (x & 31)] = (uint8_t)top;
            flag_around(a.w, a.dirty, x, y, z);                                             // FlagChunkForUpdate (World.cpp:507-546)
        }
    }
    if (lane == 0)
    {
        a.status[edit_index] = status;
        a.state[ES_SKIP] = status != GC_EDIT_APPLIED;
        // the sweeps read the word before their own: zero ends the relaxation before it starts
        a.state[ES_FLAGS] = status == GC_EDIT_APPLIED;
        a.state[ES_FLAGS + kLightSweeps + 1] = status == GC_EDIT_APPLIED;
    }
}

__global__ void __launch_bounds__(256) gc_edit_prepare_kernel(EditArgs a)
{
    if (a.state[ES_SKIP]) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    Cell c; c.inside = false;
    if (i < kEditBoxCells) c = box_cell(a, i);
    bool sun_cand = false, light_cand = false;
    if (c.inside)
    {
        const LightWorld& w = a.w;
        const uint32_t id = w.blocks[c.v] & 255u;
        const int st = w.rules->step[id], em = w.rules->emitted[id];
        const int s = surface_at(w, c.x, c.z);
        uint8_t seed = 0;
        if (em > 1 && inner_column(w, c.x, c.z))
        {
            // CheckForBlockLight up to ComputeMaxY (Lighting.cpp:28-67, 231-281); the neighbours of an inner column exist
            int my = s;
            my = max(my, surface_at(w, c.x, c.z + 1)); my = max(my, surface_at(w, c.x, c.z - 1));
            my = max(my, surface_at(w, c.x - 1, c.z)); my = max(my, surface_at(w, c.x + 1, c.z));
            if (my > GC_WORLD_BLOCK_HEIGHT - 1) my = GC_WORLD_BLOCK_HEIGHT - 1;
            if (c.y <= my) seed = (uint8_t)em;
        }
        w.sun[c.v] = 0;
        w.light[c.v] = seed;
        sun_cand = st != 255 && c.y < s;
        light_cand = st != 255;
    }
    const uint32_t ms = __ballot_sync(0xffffffffu, sun_cand), ml = __ballot_sync(0xffffffffu, light_cand);
    uint32_t bs = 0, bl = 0;
    if (lane == 0)
    {
        if (ms) bs = atomicAdd(&a.state[ES_SUN_CAND], (uint32_t)__popc(ms));
        if (ml) bl = atomicAdd(&a.state[ES_LIGHT_CAND], (uint32_t)__popc(ml));
    }
    bs = __shfl_sync(0xffffffffu, bs, 0); bl = __shfl_sync(0xffffffffu, bl, 0);
    const uint32_t below = (1u << lane) - 1u;
    if (sun_cand) a.sun_list[bs + __popc(ms & below)] = c.v;
    if (light_cand) a.light_list[bl + __popc(ml & below)] = c.v;
}

__global__ void __launch_bounds__(256) gc_edit_diff_kernel(EditArgs a)
{
    if (a.state[ES_SKIP]) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kEditBoxCells) return;
    const Cell c = box_cell(a, i);
    if (!c.inside) return;
    if (effective_sun(a.w, c) != a.save[i] || a.w.light[c.v] != a.save[kEditBoxCells + i]) flag_around(a.w, a.dirty, c.x, c.y, c.z);
}

__global__ void gc_edit_next_kernel(EditArgs a) { a.state[EP_CURSOR] += 1; }

// RecomputeLight as the reference runs it, on one thread: the queues are sequential by definition (relight_core.cuh)
__global__ void __launch_bounds__(32) gc_edit_replay_kernel(EditArgs a)
{
    if (threadIdx.x != 0 || a.state[ES_SKIP]) return;
    const GcBlockEdit e = a.edits[a.state[EP_CURSOR]];
    RelightWorld w;
    w.blocks = a.blocks_rw; w.sun = a.w.sun; w.light = a.w.light; w.surface = a.w.surface;
    w.step = a.w.rules->step; w.emitted = a.w.rules->emitted; w.vert_table = a.vert_table; w.dirty = a.dirty;
    w.size_x = a.w.size_x; w.size_z = a.w.size_z;
    RelightQueue qa = { a.queue_a, 0, 0, 0 }, qb = { a.queue_b, 0, 0, 0 };
    rl_recompute_light(w, e.wx, e.wy, e.wz, qa, qb);
}

__global__ void __launch_bounds__(256) gc_edit_list_kernel(LightWorld w, uint8_t* dirty, GcChunkCoord* out, uint32_t* n_out)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= w.size_x * w.size_z * 4) return;
    if (!dirty[slot]) return;
    dirty[slot] = 0;
    const int col = slot >> 2, cx = col % w.size_x, cz = col / w.size_x;
    if (cx < 1 || cx >= w.size_x - 1 || cz < 1 || cz >= w.size_z - 1) return;   // edge chunks are never meshed (WorldRender.cpp:235)
    GcChunkCoord c; c.cx = cx; c.cy = slot & 3; c.cz = cz;
    out[atomicAdd(n_out, 1u)] = c;
}
}  // namespace

cudaError_t launch_edit(const EditArgs& a, cudaStream_t s)
{
    const unsigned grid = (kEditBoxCells + 255u) / 256u;
    cudaError_t e = cudaMemsetAsync(a.state, 0, kEditStateWords * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    gc_edit_save_kernel<<<grid, 256, 0, s>>>(a);
    gc_edit_apply_kernel<<<1, 32, 0, s>>>(a);
    gc_edit_prepare_kernel<<<grid, 256, 0, s>>>(a);
    for (int sweep = 0; sweep < kLightSweeps; sweep++)
    {
        e = launch_relax_pair(a.w, a.sun_list, a.light_list, kEditBoxCells, a.state + ES_SUN_CAND, a.state + ES_LIGHT_CAND,
                              a.state + ES_FLAGS + 1 + sweep, a.state + ES_FLAGS + kLightSweeps + 2 + sweep, s);
        if (e != cudaSuccess) return e;
    }
    gc_edit_diff_kernel<<<grid, 256, 0, s>>>(a);
    gc_edit_next_kernel<<<1, 1, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_edit_replay(const EditArgs& a, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(a.state, 0, kEditStateWords * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    gc_edit_apply_kernel<<<1, 32, 0, s>>>(a);
    gc_edit_replay_kernel<<<1, 32, 0, s>>>(a);
    gc_edit_next_kernel<<<1, 1, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_edit_list(const LightWorld& w, uint8_t* dirty, GcChunkCoord* out, uint32_t* n_out, cudaStream_t s)
{
    const int n_slots = w.size_x * w.size_z * 4;
    cudaError_t e = cudaMemsetAsync(n_out, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    gc_edit_list_kernel<<<(n_slots + 255) / 256, 256, 0, s>>>(w, dirty, out, n_out);
    return cudaGetLastError();
}
}  // namespace gc

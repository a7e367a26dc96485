// gclight — surface map and light flood fill on the device (see gclight_kernel.cuh).
//
// The reference fills light with two FIFO queues per column (Lighting.cpp:109-143, 195-229).  What the queues compute
// is the least fixpoint of one push rule — a cell v offers  l = value(v) - lightStep[block(v)]  to each in-window,
// in-height neighbour whose block is not opaque, accepted when l > MIN_LIGHT and l > the neighbour's value — so the
// device runs it as an in-place relaxation over the cells that can change at all:
//   sunlight    candidates = non-opaque cells below their column's surface; sources = the sky cells (read as 15,
//               Lighting.cpp:69-79) of the pre-processed (non-edge) columns
//   blockLight  seeds = emitters of the pre-processed columns up to maxY (Lighting.cpp:231-281); candidates = the
//               non-opaque cells of the bricks within one brick of a seed (light travels < 15 voxels)
// A value falls by at least one per hop, so 14 sweeps reach the fixpoint whatever the order of the updates.
// The whole-window fill (gcmesh_world_light) keeps its candidates as bit maps (one word per x-row of a brick) and its active set
// per brick: nothing is sized on the host, so the call enqueues a fixed sequence of launches and never synchronises.
// Canonical form of the reference's column-order dependent block-light seeding: see oracle/gc_light_cpu.c.
#include <cooperative_groups.h>
#include <cstdlib>

#include "gclight_kernel.cuh"

namespace gc
{
namespace
{
constexpr int kVox = GC_CHUNK_SIZE_3;

__device__ __forceinline__ int surf_of(const LightWorld& w, int col, int x, int z) { return w.surface[(size_t)col * 1024 + z * 32 + x]; }

// ---- ComputeSurface in two steps.  (1) Every brick on its own: a thread owns eight x of one z row (one 16-byte load per
// layer), a warp eight rows of the brick; it walks down from the brick's top layer, four layers in flight per thread,
// until every voxel of the warp has hit a block, and leaves "world y + 1 of the highest block in this brick" (0: none) in
// a per-brick scratch map.  Bricks are independent, so the walk has four times the warps of a per-column walk and the air
// above the terrain streams at the memory rate.  (2) Per column cell, the entry of the highest brick that has one. ----
__global__ void __launch_bounds__(256) gc_surface_brick_kernel(LightWorld w, uint8_t* __restrict__ brick_top)
{
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int n_warps = w.size_x * w.size_z * 4 * 4;       // per brick: 32 rows, 8 per warp
    if (warp >= n_warps) return;
    const int slot = warp >> 2, cy = slot & 3, z = (warp & 3) * 8 + (lane >> 2), x0 = (lane & 3) * 8;
    const uint16_t* brick = w.blocks + (size_t)slot * kVox + x0 + 32 * 64 * z;
    uint32_t lo = 0, hi = 0;                               // entry of the eight voxels, one byte each
    uint32_t found = 0;                                    // bit i: voxel i has its entry
    for (int y0 = cy == 3 ? 62 : 63; y0 >= 0; y0 -= 4)     // world y = 255 is never looked at (Lighting.cpp:9)
    {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = *reinterpret_cast<const uint4*>(brick + 32 * max(y0 - k, 0));
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int y = y0 - k;
            if (y < 0) break;
            const uint32_t wds[4] = { v[k].x, v[k].y, v[k].z, v[k].w };
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                const uint32_t id = (wds[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
                if (id != 0 && !((found >> i) & 1u))
                {
                    found |= 1u << i;
                    const uint32_t top = (uint32_t)(cy * 64 + y + 1);
                    if (i < 4) lo |= top << (8 * i); else hi |= top << (8 * (i - 4));
                }
            }
        }
        if (__all_sync(0xffffffffu, found == 0xFFu)) break;
    }
    *reinterpret_cast<uint2*>(brick_top + (size_t)slot * 1024 + z * 32 + x0) = make_uint2(lo, hi);
}

__global__ void __launch_bounds__(256) gc_surface_max_kernel(LightWorld w, const uint8_t* __restrict__ brick_top)
{
    // a thread: four cells of one column (one 32-bit word per brick); entries grow with the brick, so the maximum per byte
    // is the entry of the highest brick that has one
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = w.size_x * w.size_z * 256;
    if (i >= n) return;
    const int col = i >> 8, q = i & 255;
    uint32_t best = 0;
#pragma unroll
    for (int cy = 0; cy < 4; cy++)
        best = __vmaxu4(best, reinterpret_cast<const uint32_t*>(brick_top + ((size_t)col * 4 + cy) * 1024)[q]);
    reinterpret_cast<uint32_t*>(w.surface + (size_t)col * 1024)[q] = best;
}

// ---- scan of every column cell up to maxY: candidate bit maps, emitters seeded, sunlight started at its straight-down value ----
// A warp owns one z-row of one column (lane = x) and walks down its cells, so a ballot is one word of a bit map: word
// (slot * 2048 + z * 64 + y) = voxel index >> 5.
//   sun_bits   bit x = the cell is a sunlight candidate (non-opaque, below its column's surface)
//   opq_bits   bit x = the cell is opaque (rows the walk never reaches are air: the maps are cleared beforehand)
//   sun_flag   per brick: it holds a sunlight candidate (the first sweep's active set)
//   sun_unit   per unit (four z-slices of a brick = 256 rows = 8192 cells): it holds a sunlight candidate (worth visiting at all)
//   seed_flag  per brick: it holds a seeded emitter (the first block-light sweep's active set)
__global__ void __launch_bounds__(256) gc_light_scan_kernel(LightWorld w, uint32_t* __restrict__ sun_bits, uint32_t* __restrict__ opq_bits,
                                                            uint8_t* sun_flag, uint8_t* sun_unit, uint8_t* seed_flag, uint32_t* counters)
{
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int n_rows = w.size_x * w.size_z * 32;
    if (row >= n_rows) return;
    const int col = row >> 5, z = row & 31, x = lane;
    const int cx = col % w.size_x, cz = col / w.size_x;
    const bool inner = cx >= 1 && cx < w.size_x - 1 && cz >= 1 && cz < w.size_z - 1;   // pre-processed columns (SetLightNodes)
    const int s = surf_of(w, col, x, z);
    // ComputeMaxY (Lighting.cpp:28-67); edge columns are never seeded, their cells only receive
    int my = s;
    if (inner)
    {
        const int sf = z < 31 ? surf_of(w, col, x, z + 1) : surf_of(w, col + w.size_x, x, 0);
        const int sb = z > 0 ? surf_of(w, col, x, z - 1) : surf_of(w, col - w.size_x, x, 31);
        const int sl = x > 0 ? surf_of(w, col, x - 1, z) : surf_of(w, col - 1, 31, z);
        const int sr = x < 31 ? surf_of(w, col, x + 1, z) : surf_of(w, col + 1, 0, z);
        my = max(max(max(my, sf), max(sb, sl)), sr);
        if (my > GC_WORLD_BLOCK_HEIGHT - 1) my = GC_WORLD_BLOCK_HEIGHT - 1;
    }
    // candidates and opaque cells lie below the own surface; seeds reach up to maxY
    int top_w = inner ? my : s - 1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) top_w = max(top_w, __shfl_xor_sync(0xffffffffu, top_w, d));
    uint32_t n_cand = 0, n_seed = 0;
    // The walk goes DOWN the column and carries what each cell hands to the one below it (its value less its own lightStep,
    // ScatterSunlight's rule; a sky cell of a pre-processed column reads 15): a candidate starts at the light that reaches it
    // straight from above.  That is a value some path really delivers — a lower bound of the fixpoint, so the relaxation ends
    // at the same values — and it is most of the fixpoint (open water, shafts, cave mouths): the sweeps that follow only move
    // light sideways.  (The cell above the walk's first cell is sky and AIR — every surface entry lies below it — unless the
    // walk starts at the world's top layer.)
    const int st_air = w.rules->step[0];
    int offer = (inner && st_air != 255 && (top_w | 7) < GC_WORLD_BLOCK_HEIGHT - 1) ? 15 - st_air : 0;
    for (int y0 = top_w & ~7; y0 >= 0; y0 -= 8)
    {
        uint32_t id[8];
#pragma unroll
        for (int k = 0; k < 8; k++)   // eight loads in flight per lane
            id[k] = w.blocks[((size_t)col * 4 + ((y0 + k) >> 6)) * kVox + x + 32 * (((y0 + k) & 63) + 64 * z)] & 255u;
#pragma unroll
        for (int k = 7; k >= 0; k--)
        {
            const int y = y0 + k;                                 // (<= 255: y0 is a multiple of eight below 256)
            const size_t slot = (size_t)col * 4 + (y >> 6);
            const uint32_t v = (uint32_t)(slot * kVox + x + 32 * ((y & 63) + 64 * z));
            const int st = w.rules->step[id[k]];
            const bool opaque = st == 255;
            const bool cand = y < s && !opaque;
            const uint32_t mo = __ballot_sync(0xffffffffu, opaque), mc = __ballot_sync(0xffffffffu, cand);
            if (lane == 0)
            {
                if (mo) opq_bits[v >> 5] = mo;
                if (mc) { sun_bits[v >> 5] = mc; sun_flag[slot] = 1; sun_unit[slot * 8 + (z >> 2)] = 1; n_cand += (uint32_t)__popc(mc); }
            }
            if (opaque) offer = 0;                                // neither holds nor passes light
            else if (y >= s) offer = inner ? 15 - st : 0;         // sky (only the pre-processed columns' sky cells were ever queued)
            else
            {
                const int val = offer > 1 ? offer : 0;            // (nothing at or below MIN_LIGHT is stored)
                if (val) w.sun[v] = (uint8_t)val;
                offer = val - st;
            }
            const uint32_t e = w.rules->emitted[id[k]];
            if (inner && y <= my && e > 1)   // CheckForBlockLight (Lighting.cpp:231-245); the arrays were cleared, so this is a max
            {
                w.light[v] = (uint8_t)e;
                seed_flag[slot] = 1;
                n_seed++;
            }
        }
    }
    if (top_w < GC_WORLD_BLOCK_HEIGHT - 1)
    {
        // the world's top layer lies outside ComputeSurface's walk (Lighting.cpp:9): a block there is above every surface entry
        const uint32_t v = (uint32_t)(((size_t)col * 4 + 3) * kVox + x + 32 * (63 + 64 * z));
        const uint32_t mo = __ballot_sync(0xffffffffu, w.rules->step[w.blocks[v] & 255u] == 255);
        if (lane == 0 && mo) opq_bits[v >> 5] = mo;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_seed += __shfl_xor_sync(0xffffffffu, n_seed, d);
    if (lane == 0)
    {
        if (n_cand) atomicAdd(&counters[LC_SUN_CAND], n_cand);
        if (n_seed) atomicAdd(&counters[LC_SEEDS], n_seed);
    }
}

// One candidate cell: the maximum over the six neighbours' offers, stored in place when it beats the cell's value.
template <bool kSun>
__device__ __forceinline__ bool relax_voxel(const LightWorld& w, uint32_t v)
{
    // (values are read through L2: in the persistent fill kernel another SM may have raised them since this SM last looked)
    uint8_t* arr = kSun ? w.sun : w.light;
    const int cur = __ldcg(arr + v);
    const uint32_t slot = v >> 16, in = v & 65535u;
    const int x = in & 31, y = (in >> 5) & 63, z = in >> 11;
    const int col = slot >> 2, cy = slot & 3, cx = col % w.size_x, cz = col / w.size_x;
    int best = cur;
#pragma unroll
    for (int d = 0; d < 6; d++)
    {
        // Utils.h:106-111 order is irrelevant here: the maximum over the six face neighbours
        int nx = x, ny = y, nz = z, ncx = cx, ncy = cy, ncz = cz;
        if (d == 0) { if (++ny > 63) { ny = 0; ncy++; } }
        else if (d == 1) { if (--ny < 0) { ny = 63; ncy--; } }
        else if (d == 2) { if (++nz > 31) { nz = 0; ncz++; } }
        else if (d == 3) { if (--nz < 0) { nz = 31; ncz--; } }
        else if (d == 4) { if (--nx < 0) { nx = 31; ncx--; } }
        else { if (++nx > 31) { nx = 0; ncx++; } }
        if (ncy < 0 || ncy > 3 || ncx < 0 || ncx >= w.size_x || ncz < 0 || ncz >= w.size_z) continue;   // world height / window edge
        const int ncol = ncz * w.size_x + ncx;
        const uint32_t nv = (uint32_t)(((size_t)ncol * 4 + ncy) * kVox + nx + 32 * (ny + 64 * nz));
        const int st = w.rules->step[w.blocks[nv] & 255u];
        if (st == 255) continue;                                  // opaque cells neither hold nor pass light
        int val;
        if (kSun && ncy * 64 + ny >= surf_of(w, ncol, nx, nz))
        {
            // a sky cell reads 15 (GetSunlight); only those of the pre-processed columns were ever queued (SetLightNodes)
            if (ncx < 1 || ncx >= w.size_x - 1 || ncz < 1 || ncz >= w.size_z - 1) continue;
            val = 15;
        }
        else val = __ldcg(arr + nv);                              // (a stored 0 reads as MIN_LIGHT = 1: offers nothing either way)
        best = max(best, val - st);
    }
    if (best > cur && best > 1)
    {
        arr[v] = (uint8_t)best;
        return true;
    }
    return false;
}
template <bool kSun>
__device__ __forceinline__ void relax_cell(const LightWorld& w, const uint32_t* __restrict__ list, uint32_t i, uint32_t* changed)
{
    if (relax_voxel<kSun>(w, list[i])) *changed = 1;
}

// ---- the whole-window relaxation: one persistent kernel, all sweeps ----
// Work unit: four z-slices of a brick (256 rows, 8192 cells).  Per sweep: (A) every (fill, unit) pair is tested — a brick is
// visited when it, or one of its six face neighbours (light moves by face steps), changed in the previous sweep; the first
// sweep's "previous" flags are the bricks that hold a sunlight candidate / a seeded emitter; sunlight units without a candidate
// are never visited — and the visits are appended to a list; grid barrier; (B) a CTA takes a visit off the list, turns the
// unit's 256 bit-map words into a list of cells in shared memory (popcount scan), and relaxes them one cell per thread:
// consecutive threads hold consecutive candidates, which are neighbours in x, so their loads share sectors; bricks that changed
// are flagged for the next sweep; grid barrier.  A sweep whose list is empty has reached the fixpoint and ends the kernel.
// Launched cooperatively (every CTA resident).
constexpr int kUnitRows = 256, kUnitsPerBrick = 2048 / kUnitRows;
// "did row r of brick `slot` change in the previous sweep?" — one bit per row, 64 words per brick
__device__ __forceinline__ uint32_t row_bit(const uint32_t* rows, int slot, int r) { return (__ldcg(rows + (size_t)slot * 64 + (r >> 5)) >> (r & 31)) & 1u; }

constexpr int kRelaxTeams = 4;                                   // teams of 256 threads per CTA: few, fat CTAs keep the grid barriers cheap
constexpr int kRelaxSmem = kRelaxTeams * (kUnitRows * 32 * 2 + 64);
// (barrier numbers as literals: with a number in a register ptxas reserves all sixteen hardware barriers for the CTA, and a second
// CTA no longer fits the SM)
__device__ __forceinline__ void team_sync(int team)
{
    if (team == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
    else if (team == 1) asm volatile("bar.sync 2, 256;" ::: "memory");
    else if (team == 2) asm volatile("bar.sync 3, 256;" ::: "memory");
    else asm volatile("bar.sync 4, 256;" ::: "memory");
}

__global__ void __launch_bounds__(256 * kRelaxTeams, 2) gc_relax_all_kernel(LightWorld w, const uint32_t* __restrict__ sun_bits, const uint32_t* __restrict__ opq_bits,
                                                                         const uint8_t* __restrict__ sun_unit, uint8_t* sun_flags, uint8_t* light_flags, size_t flag_stride,
                                                                         uint32_t* lists, uint32_t* row_maps, uint32_t* counters)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    extern __shared__ __align__(16) uint8_t relax_smem[];
    const int team = threadIdx.x >> 8, tid = threadIdx.x & 255, lane = tid & 31, warp = tid >> 5;
    uint16_t* s_cells = reinterpret_cast<uint16_t*>(relax_smem) + team * (kUnitRows * 32);
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(relax_smem + kRelaxTeams * kUnitRows * 32 * 2) + team * 16;
    uint32_t* s_warp_total = s_misc;                              // [8]
    uint32_t* s_row_changed = s_misc + 8;                         // [kUnitRows / 32 = 8] — word 0 .. 7; "any" is their OR
    const int n_slots = w.size_x * w.size_z * 4;
    const int n_units = n_slots * kUnitsPerBrick;
    const uint32_t vcta = blockIdx.x * kRelaxTeams + team, n_vcta = gridDim.x * kRelaxTeams;   // a team works like a CTA of its own
    for (int sweep = 0; sweep < kLightSweeps; sweep++)
    {
        const size_t prev_off = (size_t)sweep * flag_stride, now_off = prev_off + flag_stride;
        uint32_t* list = lists + (size_t)(sweep & 1) * 2 * n_units;
        uint32_t* count = counters + LC_VISITS + sweep;
        // rows that changed, one bit each, per fill (sunlight first): three maps in rotation — this sweep reads the previous
        // sweep's, writes its own, and clears the one the next sweep will write
        const size_t map_words = (size_t)2 * n_slots * 64;
        const uint32_t* prev_rows = row_maps + (size_t)(sweep % 3) * map_words;
        uint32_t* now_rows = row_maps + (size_t)((sweep + 1) % 3) * map_words;
        uint32_t* next_rows = row_maps + (size_t)((sweep + 2) % 3) * map_words;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < map_words; i += (size_t)gridDim.x * blockDim.x) next_rows[i] = 0;
        // (A) this sweep's visits
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n_units; i += gridDim.x * blockDim.x)
        {
            const int sun = i < n_units ? 1 : 0;
            const int unit = sun ? i : i - n_units;
            const int slot = unit / kUnitsPerBrick;
            if (sun && !sun_unit[unit]) continue;                 // no candidate in this unit: nothing can change here
            const uint8_t* pv = (sun ? sun_flags : light_flags) + prev_off;
            const int col = slot >> 2, cy = slot & 3, cx = col % w.size_x, cz = col / w.size_x;
            bool on = __ldcg(pv + slot) != 0;
            if (cy > 0) on |= __ldcg(pv + slot - 1) != 0;
            if (cy < 3) on |= __ldcg(pv + slot + 1) != 0;
            if (cx > 0) on |= __ldcg(pv + slot - 4) != 0;
            if (cx < w.size_x - 1) on |= __ldcg(pv + slot + 4) != 0;
            if (cz > 0) on |= __ldcg(pv + slot - 4 * w.size_x) != 0;
            if (cz < w.size_z - 1) on |= __ldcg(pv + slot + 4 * w.size_x) != 0;
            if (on) list[atomicAdd(count, 1u)] = (uint32_t)i;
        }
        grid.sync();
        const uint32_t n = __ldcg(count);
        if (n == 0) break;                                        // (the same for every thread of the grid)
        // (B) the visits
        uint32_t word = 0;
        int item = 0;
        if (vcta < n)
        {
            item = (int)__ldcg(list + vcta);
            word = ((item < n_units ? sun_bits : opq_bits))[(size_t)(item < n_units ? item : item - n_units) * kUnitRows + tid];
        }
        for (uint32_t j = vcta; j < n; j += n_vcta)
        {
            const bool sun = item < n_units;
            const int unit = sun ? item : item - n_units;
            const int slot = unit / kUnitsPerBrick;
            uint32_t m = sun ? word : ~word;                      // block light: every non-opaque cell
            if (sweep > 0 && m)
            {
                // a cell can only change when a neighbour's value has: this row, the rows above / below / in front / behind it,
                // or the same row of the bricks to the left / right changed in the previous sweep — else the row is left alone
                const uint32_t* pr = prev_rows + (sun ? (size_t)0 : (size_t)n_slots * 64);
                const int col = slot >> 2, cy = slot & 3, cx = col % w.size_x, cz = col / w.size_x;
                const int r = (unit % kUnitsPerBrick) * kUnitRows + tid, y = r & 63, z = r >> 6;
                uint32_t on = row_bit(pr, slot, r);
                if (y > 0) on |= row_bit(pr, slot, r - 1); else if (cy > 0) on |= row_bit(pr, slot - 1, r + 63);
                if (y < 63) on |= row_bit(pr, slot, r + 1); else if (cy < 3) on |= row_bit(pr, slot + 1, r - 63);
                if (z > 0) on |= row_bit(pr, slot, r - 64); else if (cz > 0) on |= row_bit(pr, slot - 4 * w.size_x, r + 31 * 64);
                if (z < 31) on |= row_bit(pr, slot, r + 64); else if (cz < w.size_z - 1) on |= row_bit(pr, slot + 4 * w.size_x, r - 31 * 64);
                if (cx > 0) on |= row_bit(pr, slot - 4, r);
                if (cx < w.size_x - 1) on |= row_bit(pr, slot + 4, r);
                if (!on) m = 0;
            }
            if (tid < kUnitRows / 32) s_row_changed[tid] = 0;
            // the next visit's bit-map word travels while this one is worked on
            if (j + n_vcta < n)
            {
                item = (int)__ldcg(list + j + n_vcta);
                word = ((item < n_units ? sun_bits : opq_bits))[(size_t)(item < n_units ? item : item - n_units) * kUnitRows + tid];
            }
            // cells of the unit -> shared list: thread = row, exclusive scan of the rows' candidate counts
            const uint32_t cnt = (uint32_t)__popc(m);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) s_warp_total[warp] = incl;
            team_sync(team);
            uint32_t base = incl - cnt, total = 0;
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                const uint32_t t = s_warp_total[k];
                if (k < warp) base += t;
                total += t;
            }
            while (m)
            {
                const uint32_t bit = (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1;
                s_cells[base++] = (uint16_t)((tid << 5) | bit);
            }
            team_sync(team);
            const uint32_t v0 = (uint32_t)unit * (uint32_t)(kUnitRows * 32);
            for (uint32_t k = tid; k < total; k += 256)
            {
                const uint32_t c = s_cells[k];
                if (sun ? relax_voxel<true>(w, v0 + c) : relax_voxel<false>(w, v0 + c))
                    atomicOr(&s_row_changed[c >> 10], 1u << ((c >> 5) & 31));
            }
            team_sync(team);
            if (tid < kUnitRows / 32 && s_row_changed[tid])
            {
                (sun ? sun_flags : light_flags)[now_off + slot] = 1;
                now_rows[(sun ? (size_t)0 : (size_t)n_slots * 64) + (size_t)slot * 64 + (unit % kUnitsPerBrick) * (kUnitRows / 32) + tid] = s_row_changed[tid];
            }
            team_sync(team);                                      // (the row words are cleared again at the top)
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) counters[LC_SWEEPS] = (uint32_t)(sweep + 1);
        grid.sync();
    }
}

// One sweep of both fills in one launch, list lengths read from device memory (the edit path: launch-latency bound, and
// only the device knows how many candidates the box holds).  The lower half of the grid strides over the sunlight list,
// the upper half over the blockLight list; each fill has its own chain of "changed" words.
__global__ void __launch_bounds__(256) gc_relax_pair_kernel(LightWorld w, const uint32_t* __restrict__ sun_list, const uint32_t* __restrict__ light_list,
                                                            const uint32_t* __restrict__ n_sun, const uint32_t* __restrict__ n_light,
                                                            uint32_t* changed_sun, uint32_t* changed_light)
{
    const uint32_t half = gridDim.x >> 1;
    const bool sun = blockIdx.x < half;
    const uint32_t* changed = sun ? changed_sun : changed_light;
    if (changed[-1] == 0) return;
    const uint32_t n = sun ? *n_sun : *n_light;
    const uint32_t stride = half * blockDim.x;
    for (uint32_t i = (sun ? blockIdx.x : blockIdx.x - half) * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        if (sun) relax_cell<true>(w, sun_list, i, changed_sun);
        else relax_cell<false>(w, light_list, i, changed_light);
    }
}
}  // namespace

static int rows_grid(const LightWorld& w) { return (int)(((size_t)w.size_x * w.size_z * 32 * 32 + 255) / 256); }

cudaError_t launch_surface(const LightWorld& w, uint8_t* brick_top, cudaStream_t s)
{
    const int grid = (int)(((size_t)w.size_x * w.size_z * 16 * 32 + 255) / 256);
    gc_surface_brick_kernel<<<grid, 256, 0, s>>>(w, brick_top);
    gc_surface_max_kernel<<<(w.size_x * w.size_z * 256 + 255) / 256, 256, 0, s>>>(w, brick_top);
    return cudaGetLastError();
}

cudaError_t launch_light_scan(const LightWorld& w, uint32_t* sun_bits, uint32_t* opq_bits, uint8_t* sun_flag, uint8_t* sun_unit, uint8_t* seed_flag, uint32_t* counters, cudaStream_t s)
{
    gc_light_scan_kernel<<<rows_grid(w), 256, 0, s>>>(w, sun_bits, opq_bits, sun_flag, sun_unit, seed_flag, counters);
    return cudaGetLastError();
}

cudaError_t launch_relax_all(const LightWorld& w, const uint32_t* sun_bits, const uint32_t* opq_bits, const uint8_t* sun_unit, uint8_t* sun_flags, uint8_t* light_flags,
                             size_t flag_stride, uint32_t* lists, uint32_t* row_maps, uint32_t* counters, int num_sms, int ctas_per_sm, cudaStream_t s)
{
    // every CTA has to be resident (grid barriers): the occupancy query bounds the grid, the cooperative launch enforces it.
    // (a barrier costs more the more CTAs take part: two CTAs of four 256-thread teams per SM — 2048 threads — instead of eight CTAs)
    static int per_sm = 0;
    if (!per_sm)
    {
        cudaError_t e = cudaFuncSetAttribute(gc_relax_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRelaxSmem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gc_relax_all_kernel, 256 * kRelaxTeams, kRelaxSmem);
        if (e != cudaSuccess) return e;
        if (per_sm > 2) per_sm = 2;
        if (const char* env = getenv("GCMESH_LIGHT_CTAS_PER_SM")) { const int v = atoi(env); if (v >= 1 && v < per_sm) per_sm = v; }
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    }
    LightWorld lw = w;
    void* args[] = { &lw, &sun_bits, &opq_bits, &sun_unit, &sun_flags, &light_flags, &flag_stride, &lists, &row_maps, &counters };
    // (a cooperative grid starts only when all of it fits: with every thread slot of every SM asked for, it waits for whatever
    // else is running — the PCIe pull kernels of gcmesh_mesh_host, say; half of that leaves them room)
    const int use = ctas_per_sm >= 1 && ctas_per_sm < per_sm ? ctas_per_sm : per_sm;
    return cudaLaunchCooperativeKernel((const void*)gc_relax_all_kernel, dim3(num_sms * use), dim3(256 * kRelaxTeams), args, kRelaxSmem, s);
}

cudaError_t launch_relax_pair(const LightWorld& w, const uint32_t* sun_list, const uint32_t* light_list, uint32_t cap, const uint32_t* n_sun,
                              const uint32_t* n_light, uint32_t* changed_sun, uint32_t* changed_light, cudaStream_t s)
{
    // enough CTAs to cover `cap` cells per fill in at most a few strides, few enough that a sweep with nothing to do is cheap
    unsigned half = (cap + 255u) / 256u;
    if (half > 296u) half = 296u;
    gc_relax_pair_kernel<<<2 * half, 256, 0, s>>>(w, sun_list, light_list, n_sun, n_light, changed_sun, changed_light);
    return cudaGetLastError();
}
}  // namespace gc

// gcmesh — the chunk-mesh kernel for sm_100a (B200).
//
// One persistent CTA per SM; each CTA pulls chunks from a work counter and, per chunk:
//   P0  neighbour columns' surface window -> shared memory
//   P1  34 z-slabs (32x64 voxels + 2 halo rows, three arrays) are staged by TMA tensor-map box loads
//       (cp.async.bulk.tensor.4d, 6-stage mbarrier ring, issued by one elected lane) and converted by
//       role-specialised warps into the resident on-chip form (class bit planes + packed light tile)
//   P1x the x = -1 / x = 32 halo cells that some face or light footprint will actually read (rows next to an
//       emitting boundary voxel) are gathered from the neighbour columns; air chunks gather nothing
//   P2  face masks per row (bitwise, 32 voxels per op), per-row counts, per-layer warp reductions, scan,
//       ONE atomicAdd reserves the chunk's contiguous segment in the batch arenas
//   P3  per batch of layers: (a) rows -> ordered quad list + index words, (b) one thread per quad ->
//       3x3 light neighbourhood from shared memory -> 4 vertices -> three 16-byte stores
// The per-lane work is in mesh_core.cuh (shared with the host emulator used by the CPU tests).
//
// Reference path replaced: BuildChunkAsync / BuildBlock / VertexLight, Code/WorldRender.cpp:6-27,329-383,445-560.
#include "gcmesh_kernel.cuh"

namespace gc
{
namespace
{
constexpr int NT = kMeshThreads;          // 20 warps
constexpr int kWarps = NT / 32;
constexpr int kStages = 6;
constexpr int kSlabs = kRowsZ;            // 34 z-slabs per chunk: tz = -1 .. 32
// P1 consumers: kTeams teams of kTeamWarps warps (half class, half light).  A team owns every kTeams-th slab, so
// several slabs are being converted at once and each lane carries a few independent tasks (ILP) — one warp per
// task per slab left the SM waiting on its own ALU latency.
constexpr int kTeams = 3, kTeamWarps = 6, kTeamRoleLanes = (kTeamWarps / 2) * 32;
constexpr int kConsumerWarps = kTeams * kTeamWarps;
constexpr int kTasksPerSlab = kRowsY * 4;                  // (row, octet) for class and again for light
constexpr int kTasksPerLane = (kTasksPerSlab + kTeamRoleLanes - 1) / kTeamRoleLanes;
static_assert(1 + kConsumerWarps <= kWarps, "warp roles");
constexpr int kHaloTasks = kHaloRows * 2;
constexpr int kHaloPerThread = (kHaloTasks + NT - 1) / NT;

// ---- one staging slot: a z-slab of the three arrays, each TMA destination 128-byte aligned ----
constexpr int kStBlocks = 0;                       // [64][32] u16
constexpr int kStBlocksLo = 4096;                  // row ty = -1 (64 B used)
constexpr int kStBlocksHi = 4096 + 128;            // row ty = 64
constexpr int kStSun = 4096 + 256;                 // [64][32] u8
constexpr int kStSunLo = kStSun + 2048;
constexpr int kStSunHi = kStSunLo + 128;
constexpr int kStLight = kStSunHi + 128;           // 6656
constexpr int kStLightLo = kStLight + 2048;
constexpr int kStLightHi = kStLightLo + 128;
constexpr int kStageBytes = kStLightHi + 128;      // 8960
static_assert(kStageBytes % 128 == 0, "stage alignment");
constexpr int kListCap = kStages * kStageBytes / 4;  // the quad list reuses the staging ring in P3
static_assert(kListCap >= kMaxQuads, "a whole chunk's quads must fit the list");

// ---- shared-memory carve-up (dynamic, 1024-aligned base) ----
constexpr int kOffStages = 0;
constexpr int kOffPlanes = kOffStages + kStages * kStageBytes;            // 53760
constexpr int kOffLt = kOffPlanes + kPlanes * kHaloRows * 4;              // +71808
constexpr int kOffHx = kOffLt + kLtBytes;                                 // +89760
constexpr int kOffBv = kOffHx + ((kHaloRows * 2 + 15) & ~15);             // [2][2256]: emitting voxel at x = 0 / x = 31 of the row
constexpr int kBvStride = (kHaloRows + 15) & ~15;
constexpr int kOffSurf = kOffBv + 2 * kBvStride;                          // [34][32] centre-x surface rows
constexpr int kOffSurfMax = kOffSurf + kRowsZ * 32;                       // [34][4] max surface per 8-voxel octet
constexpr int kOffLut = kOffSurfMax + ((kRowsZ * 4 + 15) & ~15);
constexpr int kOffMat = kOffLut + 256 * 8;
constexpr int kOffCls = kOffMat + 256 * 8;
constexpr int kOffLayerCnt = kOffCls + 256;                               // u32[64]
constexpr int kOffLayerTypA = kOffLayerCnt + 256;                         // u32[64]  types 1,2 (16 bit each)
constexpr int kOffLayerTypB = kOffLayerTypA + 256;                        // u32[64]  types 3,4
constexpr int kOffLayerBase = kOffLayerTypB + 256;                        // u32[65]
constexpr int kOffLayerTypBaseA = kOffLayerBase + 272;
constexpr int kOffLayerTypBaseB = kOffLayerTypBaseA + 256;
constexpr int kOffFaces = kOffLayerTypBaseB + 256;                        // FaceDesc[6], 16 B each
constexpr int kOffBars = kOffFaces + 96;                                  // full[6], empty[6]
constexpr int kOffCtx = kOffBars + 2 * kStages * 8;
constexpr int kSmemBytes = kOffCtx + 64;
static_assert(kOffPlanes % 16 == 0 && kOffLt % 16 == 0 && kOffHx % 16 == 0 && kOffBv % 16 == 0 && kOffSurf % 16 == 0, "align");
static_assert(kOffLut % 8 == 0 && kOffMat % 8 == 0 && kOffBars % 8 == 0 && kOffLayerCnt % 4 == 0 && kOffCtx % 4 == 0, "align");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct alignas(16) FaceDesc16 { FaceDesc f; uint8_t pad[3]; };
static_assert(sizeof(FaceDesc16) == 16, "FaceDesc16");

struct ChunkCtx
{
    int32_t chunk;        // index into coords, or -1 when the work is exhausted
    int32_t cx, cy, cz;
    int32_t any_emit;     // some voxel of the chunk emits quads (else P1x / P2 / P3 are skipped)
    int32_t any_bv;       // some emitting voxel sits at x = 0 or x = 31 (else no x-halo cell is ever read)
    int32_t emit;         // P3 runs
    uint32_t quad_base;   // arena segment (quads)
    uint32_t total;
    uint32_t type_off[kMeshTypes];
};

// ---- PTX helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)   // suspend-time hint (ns): sleep in hardware instead of spinning
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (and raises the error flag) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* error_flag)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
    {
        if (++spins > (1u << 22))
        {
            if (error_flag) atomicExch(error_flag, 1);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

#define GC_PROF_MARK(slot)                                                            \
    do {                                                                              \
        if (p.prof && tid == 0) { const long long t_ = clock64(); atomicAdd(p.prof + (slot), (unsigned long long)(t_ - prof_t)); prof_t = t_; } \
    } while (0)

// =====================================================================================================
__global__ void __launch_bounds__(NT, 1)
gc_mesh_kernel(const MeshParams p, const __grid_constant__ MeshTensorMaps maps)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint8_t* s_stage = smem + kOffStages;
    uint32_t* s_list = reinterpret_cast<uint32_t*>(smem + kOffStages);
    uint32_t* s_planes = reinterpret_cast<uint32_t*>(smem + kOffPlanes);
    uint8_t* s_planes_b = smem + kOffPlanes;
    uint8_t* s_lt = smem + kOffLt;
    uint16_t* s_hx = reinterpret_cast<uint16_t*>(smem + kOffHx);
    uint8_t* s_hx_b = smem + kOffHx;
    uint8_t* s_bv = smem + kOffBv;
    uint8_t* s_surf = smem + kOffSurf;
    uint8_t* s_surfmax = smem + kOffSurfMax;
    LutEntry* s_lut = reinterpret_cast<LutEntry*>(smem + kOffLut);
    MatEntry* s_mat = reinterpret_cast<MatEntry*>(smem + kOffMat);
    uint8_t* s_cls = smem + kOffCls;
    uint32_t* s_layer_cnt = reinterpret_cast<uint32_t*>(smem + kOffLayerCnt);
    uint32_t* s_layer_typa = reinterpret_cast<uint32_t*>(smem + kOffLayerTypA);
    uint32_t* s_layer_typb = reinterpret_cast<uint32_t*>(smem + kOffLayerTypB);
    uint32_t* s_layer_base = reinterpret_cast<uint32_t*>(smem + kOffLayerBase);
    uint32_t* s_layer_typbase_a = reinterpret_cast<uint32_t*>(smem + kOffLayerTypBaseA);
    uint32_t* s_layer_typbase_b = reinterpret_cast<uint32_t*>(smem + kOffLayerTypBaseB);
    FaceDesc16* s_faces = reinterpret_cast<FaceDesc16*>(smem + kOffFaces);
    uint64_t* s_bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    ChunkCtx* s_ctx = reinterpret_cast<ChunkCtx*>(smem + kOffCtx);

    // ---- one-time setup ----
    for (int i = tid; i < 256; i += NT)
    {
        s_lut[i] = p.tables->lut[i];
        s_mat[i] = p.tables->mat[i];
        s_cls[i] = p.tables->cls8[i];
    }
    if (tid < 6)
    {
        const FaceDesc table[6] = GC_FACE_TABLE;
        s_faces[tid].f = table[tid];
    }
    if (tid == 0)
    {
        for (int i = 0; i < kStages; i++)
        {
            mbar_init(smem_u32(&s_bars[i]), 1);                         // full: the producer's expect_tx arrive
            mbar_init(smem_u32(&s_bars[kStages + i]), kTeamWarps);      // empty: one arrive per warp of the owning team
        }
        fence_barrier_init();
        if (p.use_tma)
        {
            prefetch_tensormap(&maps.blocks_slab); prefetch_tensormap(&maps.blocks_row);
            prefetch_tensormap(&maps.sun_slab); prefetch_tensormap(&maps.sun_row);
            prefetch_tensormap(&maps.light_slab); prefetch_tensormap(&maps.light_row);
        }
    }
    const int kill_cls = p.tables->cls8[p.tables->kill_zone & 255];
    const int air_cls = p.tables->cls8[0];
    __syncthreads();

    uint32_t slab_seq = 0;  // slabs staged so far by this CTA (ring position / parity source)
    long long prof_t = clock64();

    for (;;)
    {
        // ================= P0: next chunk, surface window =================
        if (tid == 0)
        {
            int c = atomicAdd(p.work_counter, 1);
            if (c >= p.n_chunks) c = -1;
            s_ctx->chunk = c;
            s_ctx->any_emit = 0; s_ctx->any_bv = 0;
            if (c >= 0)
            {
                GcChunkCoord cc = p.coords[c];
                s_ctx->cx = cc.cx; s_ctx->cy = cc.cy; s_ctx->cz = cc.cz;
            }
        }
        __syncthreads();
        const int chunk = s_ctx->chunk;
        if (chunk < 0) break;
        const int cx = s_ctx->cx, cy = s_ctx->cy, cz = s_ctx->cz;
        const int lwy = cy * kTY;

        // centre-x surface rows of the three columns (cz-1, cz, cz+1): s_surf[(tz+1)*32 + x], and their per-octet maxima
        if (tid < kRowsZ * 4)
        {
            const int r = tid >> 2, q = tid & 3;
            const int tz = r - 1;
            const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
            const size_t col = (size_t)(cz + dcz) * p.size_x + cx;
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p.surface + col * 1024 + (tz & 31) * 32) + q);
            reinterpret_cast<uint2*>(s_surf)[tid] = v;
            const uint32_t m = max4(v.x, v.y);
            const uint32_t m2 = max4(m, m >> 16);
            s_surfmax[tid] = (uint8_t)max4(m2, m2 >> 8);
        }
        __syncthreads();

        GC_PROF_MARK(0);
        // ================= P1: stage + convert =================
        if (warp == 0)
        {
            // ---- producer ----
            for (int s = 0; s < kSlabs; s++)
            {
                const uint32_t g = slab_seq + s;
                const int stage = g % kStages;
                const uint32_t full = smem_u32(&s_bars[stage]), empty = smem_u32(&s_bars[kStages + stage]);
                const int tz = s - 1;
                const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
                const int zz = tz & 31;
                const int col = (cz + dcz) * p.size_x + cx;
                const int slot = col * 4 + cy;
                const bool has_lo = lwy > 0, has_hi = lwy + kTY < GC_WORLD_BLOCK_HEIGHT;
                uint8_t* st = s_stage + stage * kStageBytes;

                if (g >= kStages)
                {
                    const long long w0 = p.prof ? clock64() : 0;
                    mbar_wait(empty, ((g / kStages) + 1) & 1, p.error_flag);
                    if (p.prof && lane == 0) atomicAdd(p.prof + 8, (unsigned long long)(clock64() - w0));
                }

                if (p.use_tma)
                {
                    if (lane == 0)
                    {
                        const uint32_t bytes = 8192u + (has_lo ? 128u : 0u) + (has_hi ? 128u : 0u);
                        mbar_expect_tx(full, bytes);
                        const uint32_t d = smem_u32(st);
                        tma_load_4d(d + kStBlocks, &maps.blocks_slab, 0, 0, zz, slot, full);
                        tma_load_4d(d + kStSun, &maps.sun_slab, 0, 0, zz, slot, full);
                        tma_load_4d(d + kStLight, &maps.light_slab, 0, 0, zz, slot, full);
                        if (has_lo)
                        {
                            tma_load_4d(d + kStBlocksLo, &maps.blocks_row, 0, 63, zz, slot - 1, full);
                            tma_load_4d(d + kStSunLo, &maps.sun_row, 0, 63, zz, slot - 1, full);
                            tma_load_4d(d + kStLightLo, &maps.light_row, 0, 63, zz, slot - 1, full);
                        }
                        if (has_hi)
                        {
                            tma_load_4d(d + kStBlocksHi, &maps.blocks_row, 0, 0, zz, slot + 1, full);
                            tma_load_4d(d + kStSunHi, &maps.sun_row, 0, 0, zz, slot + 1, full);
                            tma_load_4d(d + kStLightHi, &maps.light_row, 0, 0, zz, slot + 1, full);
                        }
                    }
                }
                else
                {
                    // debug staging path: the producer warp copies the slab with plain 16-byte loads
                    const size_t vox = (size_t)slot * GC_CHUNK_SIZE_3 + (size_t)zz * 2048;
                    const uint4* gb = reinterpret_cast<const uint4*>(p.blocks + vox);
                    const uint4* gs = reinterpret_cast<const uint4*>(p.sun + vox);
                    const uint4* gl = reinterpret_cast<const uint4*>(p.light + vox);
                    for (int i = lane; i < 256; i += 32) reinterpret_cast<uint4*>(st + kStBlocks)[i] = gb[i];
                    for (int i = lane; i < 128; i += 32)
                    {
                        reinterpret_cast<uint4*>(st + kStSun)[i] = gs[i];
                        reinterpret_cast<uint4*>(st + kStLight)[i] = gl[i];
                    }
                    if (has_lo && lane < 8)
                    {
                        const size_t v = (size_t)(slot - 1) * GC_CHUNK_SIZE_3 + (size_t)zz * 2048 + 63 * 32;
                        if (lane < 4) reinterpret_cast<uint4*>(st + kStBlocksLo)[lane] = reinterpret_cast<const uint4*>(p.blocks + v)[lane];
                        else if (lane < 6) reinterpret_cast<uint4*>(st + kStSunLo)[lane - 4] = reinterpret_cast<const uint4*>(p.sun + v)[lane - 4];
                        else reinterpret_cast<uint4*>(st + kStLightLo)[lane - 6] = reinterpret_cast<const uint4*>(p.light + v)[lane - 6];
                    }
                    if (has_hi && lane >= 8 && lane < 16)
                    {
                        const int l = lane - 8;
                        const size_t v = (size_t)(slot + 1) * GC_CHUNK_SIZE_3 + (size_t)zz * 2048;
                        if (l < 4) reinterpret_cast<uint4*>(st + kStBlocksHi)[l] = reinterpret_cast<const uint4*>(p.blocks + v)[l];
                        else if (l < 6) reinterpret_cast<uint4*>(st + kStSunHi)[l - 4] = reinterpret_cast<const uint4*>(p.sun + v)[l - 4];
                        else reinterpret_cast<uint4*>(st + kStLightHi)[l - 6] = reinterpret_cast<const uint4*>(p.light + v)[l - 6];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full);
                }
            }
        }
        else if (warp <= kConsumerWarps)
        {
            const int cw = warp - 1;
            const int team = cw / kTeamWarps, wt = cw % kTeamWarps;
            const bool is_class = wt < kTeamWarps / 2;
            const int l96 = (wt % (kTeamWarps / 2)) * 32 + lane;       // lane index within the team's class / light half
            for (int s = team; s < kSlabs; s += kTeams)
            {
                const uint32_t g = slab_seq + s;
                const int stage = g % kStages;
                {
                    const long long w0 = p.prof ? clock64() : 0;
                    mbar_wait(smem_u32(&s_bars[stage]), (g / kStages) & 1, p.error_flag);
                    if (p.prof && lane == 0 && cw == 0) atomicAdd(p.prof + 9, (unsigned long long)(clock64() - w0));
                    if (p.prof && lane == 0 && cw == kTeamWarps / 2) atomicAdd(p.prof + 10, (unsigned long long)(clock64() - w0));
                }
                const uint8_t* st = s_stage + stage * kStageBytes;
                const bool centre_z = s >= 1 && s <= kTZ;
                if (is_class)
                {
                    // ---- block ids -> 8 class bit planes (+ "boundary voxel emits" flags for P1x) ----
#pragma unroll
                    for (int k = 0; k < kTasksPerLane; k++)
                    {
                        const int task = l96 + k * kTeamRoleLanes;        // (row, octet)
                        if (task < kTasksPerSlab)
                        {
                            const int ty = (task >> 2) - 1, o = task & 3;
                            const int wy = lwy + ty;
                            const int hr = hrow(ty, s - 1);
                            uint32_t vis = 0;
                            if (wy < 0) p1_class_const_octet(kill_cls, s_planes_b, hr, o);
                            else if (wy >= GC_WORLD_BLOCK_HEIGHT) p1_class_const_octet(air_cls, s_planes_b, hr, o);
                            else
                            {
                                const int row_off = ty < 0 ? kStBlocksLo : (ty >= kTY ? kStBlocksHi : kStBlocks + ty * 64);
                                const uint4 v = *reinterpret_cast<const uint4*>(st + row_off + o * 16);
                                vis = p1_class_octet(v.x, v.y, v.z, v.w, s_lut, s_planes_b, hr, o);
                            }
                            if (!(centre_z && ty >= 0 && ty < kTY)) vis = 0;
                            if (o == 0) s_bv[hr] = (uint8_t)(vis & 1u);
                            if (o == 3) s_bv[kBvStride + hr] = (uint8_t)(vis >> 7);
                            if (vis) s_ctx->any_emit = 1;
                            if ((o == 0 && (vis & 1u)) || (o == 3 && (vis >> 7))) s_ctx->any_bv = 1;
                        }
                    }
                }
                else
                {
                    // ---- sunlight + blockLight + surface -> packed light tile ----
#pragma unroll
                    for (int k = 0; k < kTasksPerLane; k++)
                    {
                        const int task = l96 + k * kTeamRoleLanes;        // (row, octet)
                        if (task < kTasksPerSlab)
                        {
                            const int ty = (task >> 2) - 1, o = task & 3;
                            const int wy = lwy + ty;
                            uint32_t* dst = reinterpret_cast<uint32_t*>(s_lt + hrow(ty, s - 1) * kLtStride + 4 + o * 8);
                            if (wy < 0) { dst[0] = 0u; dst[1] = 0u; }
                            else if (wy >= GC_WORLD_BLOCK_HEIGHT) { dst[0] = ~0u; dst[1] = ~0u; }
                            else
                            {
                                const int sun_off = (ty < 0 ? kStSunLo : (ty >= kTY ? kStSunHi : kStSun + ty * 32)) + o * 8;
                                const int light_off = (ty < 0 ? kStLightLo : (ty >= kTY ? kStLightHi : kStLight + ty * 32)) + o * 8;
                                const uint2 bl = *reinterpret_cast<const uint2*>(st + light_off);
                                if (wy >= (int)s_surfmax[s * 4 + o])
                                {
                                    dst[0] = p1_light4_above(bl.x);
                                    dst[1] = p1_light4_above(bl.y);
                                }
                                else
                                {
                                    const uint2 sn = *reinterpret_cast<const uint2*>(st + sun_off);
                                    const uint2 sf = *reinterpret_cast<const uint2*>(s_surf + s * 32 + o * 8);
                                    dst[0] = p1_light4(sn.x, bl.x, sf.x, wy);
                                    dst[1] = p1_light4(sn.y, bl.y, sf.y, wy);
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&s_bars[kStages + stage]));
            }
        }
        slab_seq += kSlabs;
        __syncthreads();
        GC_PROF_MARK(1);

        // ================= P1x: the x-halo cells that will be read =================
        const bool any_emit = s_ctx->any_emit != 0, any_bv = s_ctx->any_bv != 0;
        if (any_emit)
        {
            uint32_t id[kHaloPerThread], sn[kHaloPerThread], bl[kHaloPerThread], sf[kHaloPerThread];
            uint32_t need = 0;
#pragma unroll
            for (int k = 0; k < kHaloPerThread; k++)
            {
                const int task = tid + k * NT;
                if (task < kHaloTasks)
                {
                    const int hr = task >> 1, side = task & 1;
                    const int ty = hr / kRowsZ - 1, tz = hr % kRowsZ - 1;
                    const int wy = lwy + ty;
                    if (any_bv && wy >= 0 && wy < GC_WORLD_BLOCK_HEIGHT && xhalo_needed(s_bv + side * kBvStride, hr))
                    {
                        need |= 1u << k;
                        const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
                        const int zz = tz & 31, xx = side ? 0 : 31;
                        const size_t col = (size_t)(cz + dcz) * p.size_x + (cx + (side ? 1 : -1));
                        const size_t vox = (col * 4 + (wy >> 6)) * GC_CHUNK_SIZE_3 + xx + 32 * ((wy & 63) + 64 * zz);
                        id[k] = __ldg(p.blocks + vox);
                        sn[k] = __ldg(p.sun + vox);
                        bl[k] = __ldg(p.light + vox);
                        sf[k] = __ldg(p.surface + col * 1024 + zz * 32 + xx);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kHaloPerThread; k++)
            {
                const int task = tid + k * NT;
                if (task < kHaloTasks)
                {
                    const int hr = task >> 1, side = task & 1;
                    const int wy = lwy + hr / kRowsZ - 1;
                    if (wy < 0) p1_xhalo_cell(kill_cls, 0x00u, s_hx_b, s_lt, hr, side);
                    else if (wy >= GC_WORLD_BLOCK_HEIGHT) p1_xhalo_cell(air_cls, 0xFFu, s_hx_b, s_lt, hr, side);
                    else if ((need >> k) & 1u) p1_xhalo_cell(s_cls[id[k] & 255u], light_byte_of(sn[k], bl[k], sf[k], wy), s_hx_b, s_lt, hr, side);
                }
            }
        }
        __syncthreads();

        GC_PROF_MARK(2);
        // ================= P2: face masks, counts, scan, reservation =================
        if (!any_emit)
        {
            // nothing in the chunk can emit (all AIR / invisible): empty result, no scan, no emission
            if (tid == 0)
            {
                GcChunkOut o;
                o.quad_base = 0; o.vert_count = 0; o.quads = 0; o.flags = 0; o.reserved[0] = 0; o.reserved[1] = 0;
                for (int t = 0; t < kMeshTypes; t++) { o.index_count[t] = 0; o.index_offset[t] = 0; }
                p.out[chunk] = o;
            }
            __syncthreads();   // s_ctx is rewritten by thread 0 at the top of the loop
            GC_PROF_MARK(3);
            continue;
        }
        for (int layer = warp; layer < kTY; layer += kWarps)
        {
            RowFaces rf;
            p2_row_faces(s_planes, s_hx, layer, lane, rf);
            int c; uint32_t t4;
            p2_row_counts(rf, c, t4);
            uint32_t ta = (t4 & 255u) | (((t4 >> 8) & 255u) << 16);
            uint32_t tb = ((t4 >> 16) & 255u) | ((t4 >> 24) << 16);
            uint32_t cc = (uint32_t)c;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                cc += __shfl_xor_sync(0xffffffffu, cc, d);
                ta += __shfl_xor_sync(0xffffffffu, ta, d);
                tb += __shfl_xor_sync(0xffffffffu, tb, d);
            }
            if (lane == 0) { s_layer_cnt[layer] = cc; s_layer_typa[layer] = ta; s_layer_typb[layer] = tb; }
        }
        __syncthreads();
        if (warp == 0)
        {
            // exclusive scan over the 64 layers, two per lane
            const uint32_t c0 = s_layer_cnt[2 * lane], c1 = s_layer_cnt[2 * lane + 1];
            const uint32_t a0 = s_layer_typa[2 * lane], a1 = s_layer_typa[2 * lane + 1];
            const uint32_t b0 = s_layer_typb[2 * lane], b1 = s_layer_typb[2 * lane + 1];
            const uint32_t ci = warp_incl_scan(c0 + c1, lane);
            const uint32_t total = __shfl_sync(0xffffffffu, ci, 31);
            const uint32_t ce = ci - (c0 + c1);
            s_layer_base[2 * lane] = ce;
            s_layer_base[2 * lane + 1] = ce + c0;
            if (lane == 31) s_layer_base[64] = total;
            uint32_t ta_tot = 0, tb_tot = 0;
            if (total <= (uint32_t)kMaxQuads)   // 16-bit fields cannot overflow below the cap
            {
                const uint32_t ai = warp_incl_scan(a0 + a1, lane), bi = warp_incl_scan(b0 + b1, lane);
                ta_tot = __shfl_sync(0xffffffffu, ai, 31);
                tb_tot = __shfl_sync(0xffffffffu, bi, 31);
                const uint32_t ae = ai - (a0 + a1), be = bi - (b0 + b1);
                s_layer_typbase_a[2 * lane] = ae; s_layer_typbase_a[2 * lane + 1] = ae + a0;
                s_layer_typbase_b[2 * lane] = be; s_layer_typbase_b[2 * lane + 1] = be + b0;
            }
            if (lane == 0)
            {
                GcChunkOut o;
                o.quad_base = 0; o.vert_count = 0; o.quads = total; o.flags = 0; o.reserved[0] = 0; o.reserved[1] = 0;
                for (int t = 0; t < kMeshTypes; t++) { o.index_count[t] = 0; o.index_offset[t] = 0; }
                int emit = 0;
                if (total > (uint32_t)kMaxQuads) o.flags = GC_CHUNK_OVERFLOW;
                else
                {
                    uint32_t tq[kMeshTypes];
                    tq[1] = ta_tot & 0xFFFFu; tq[2] = ta_tot >> 16; tq[3] = tb_tot & 0xFFFFu; tq[4] = tb_tot >> 16;
                    tq[0] = total - tq[1] - tq[2] - tq[3] - tq[4];
                    unsigned long long base = total ? atomicAdd(p.quad_cursor, (unsigned long long)total) : 0ull;
                    if (base + total > p.max_quads) o.flags = GC_CHUNK_NO_SPACE;
                    else
                    {
                        o.quad_base = (uint32_t)base;
                        o.vert_count = total * 4;
                        uint32_t acc = 0;
                        for (int t = 0; t < kMeshTypes; t++)
                        {
                            s_ctx->type_off[t] = acc;
                            o.index_offset[t] = acc * 6; o.index_count[t] = tq[t] * 6;
                            acc += tq[t];
                        }
                        emit = total > 0;
                    }
                }
                s_ctx->emit = emit;
                s_ctx->quad_base = o.quad_base;
                s_ctx->total = total;
                p.out[chunk] = o;
            }
        }
        __syncthreads();

        GC_PROF_MARK(3);
        // ================= P3: emission =================
        if (s_ctx->emit)
        {
            const uint32_t quad_base = s_ctx->quad_base, total = s_ctx->total;
            uint32_t type_off[kMeshTypes];
#pragma unroll
            for (int t = 0; t < kMeshTypes; t++) type_off[t] = s_ctx->type_off[t];
            uint32_t* idx_words = reinterpret_cast<uint32_t*>(p.index_arena) + (size_t)quad_base * 3;
            uint4* vtx = reinterpret_cast<uint4*>(p.vertex_arena) + (size_t)quad_base * 3;
            const uint16_t* chunk_blocks = p.blocks + ((size_t)(cz * p.size_x + cx) * 4 + cy) * GC_CHUNK_SIZE_3;

            // ---- 3a: rows -> ordered quad list (whole chunk fits the staging ring) + index words ----
            for (int layer = warp; layer < kTY; layer += kWarps)
            {
                if (s_layer_base[layer + 1] == s_layer_base[layer]) continue;   // warp-uniform
                RowFaces rf;
                p2_row_faces(s_planes, s_hx, layer, lane, rf);
                int ci; uint32_t t4;
                p2_row_counts(rf, ci, t4);
                const uint32_t c = (uint32_t)ci;
                const uint32_t ta = (t4 & 255u) | (((t4 >> 8) & 255u) << 16);
                const uint32_t tb = ((t4 >> 16) & 255u) | ((t4 >> 24) << 16);
                const uint32_t ce = warp_incl_scan(c, lane) - c;
                const uint32_t ae = warp_incl_scan(ta, lane) - ta + s_layer_typbase_a[layer];
                const uint32_t be = warp_incl_scan(tb, lane) - tb + s_layer_typbase_b[layer];
                if (c)
                {
                    uint32_t trank[kMeshTypes];
                    const uint32_t rank0 = s_layer_base[layer] + ce;
                    trank[1] = ae & 0xFFFFu; trank[2] = ae >> 16; trank[3] = be & 0xFFFFu; trank[4] = be >> 16;
                    trank[0] = rank0 - trank[1] - trank[2] - trank[3] - trank[4];
                    p3a_row(rf, layer, lane, rank0, trank, 0u, s_list, idx_words, type_off);
                }
            }
            __syncthreads();
            GC_PROF_MARK(4);
            // ---- 3b: one thread per quad ----
            for (uint32_t i = tid; i < total; i += NT)
            {
                const uint32_t e = s_list[i];
                const int x = (e >> 3) & 31, tz = (e >> 8) & 31, ty = (e >> 13) & 63;
                const uint32_t id = __ldg(chunk_blocks + x + 32 * (ty + 64 * tz)) & 255u;
                uint32_t v[12];
                p3b_quad(e, s_faces[e & 7].f, s_lt, s_planes + P_OPAQUE * kHaloRows, s_hx, s_mat[id], v);
                uint4* dst = vtx + (size_t)i * 3;
                dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                dst[2] = make_uint4(v[8], v[9], v[10], v[11]);
            }
        }
        // the staging ring was used as the quad list by the generic proxy: order it before the next TMA writes
        fence_proxy_async();
        __syncthreads();
        GC_PROF_MARK(5);
    }
}

// ---- slab halo pack / unpack (multi-GPU) ----
// Buffer layout for one row of columns: per column cx, per chunk cy: blocks plane (2048 u16), sun plane (2048 B),
// light plane (2048 B); then per column the surface row z (32 B).
__global__ void gc_halo_copy_kernel(uint16_t* blocks, uint8_t* sun, uint8_t* light, uint8_t* surface,
                                    int size_x, int cz, int z, uint8_t* buf, int unpack)
{
    const int per_chunk = 8192 / 16;                      // uint4 per (column, cy)
    const int n_planes = size_x * 4;
    const int total = n_planes * per_chunk;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        const int plane = i / per_chunk, q = i % per_chunk;
        const int cx = plane >> 2, cy = plane & 3;
        const size_t slot = ((size_t)cz * size_x + cx) * 4 + cy;
        uint4* b = reinterpret_cast<uint4*>(buf + (size_t)plane * 8192) + q;
        uint4* g;
        if (q < 256) g = reinterpret_cast<uint4*>(blocks + slot * GC_CHUNK_SIZE_3 + (size_t)z * 2048) + q;
        else if (q < 384) g = reinterpret_cast<uint4*>(sun + slot * GC_CHUNK_SIZE_3 + (size_t)z * 2048) + (q - 256);
        else g = reinterpret_cast<uint4*>(light + slot * GC_CHUNK_SIZE_3 + (size_t)z * 2048) + (q - 384);
        if (unpack) *g = *b; else *b = *g;
    }
    const int surf_words = size_x * 8;                    // uint32 per row of columns
    uint32_t* sb = reinterpret_cast<uint32_t*>(buf + (size_t)n_planes * 8192);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < surf_words; i += gridDim.x * blockDim.x)
    {
        const int cx = i >> 3, q = i & 7;
        uint32_t* g = reinterpret_cast<uint32_t*>(surface + ((size_t)cz * size_x + cx) * 1024 + z * 32) + q;
        if (unpack) *g = sb[i]; else sb[i] = *g;
    }
}
}  // namespace

size_t mesh_smem_bytes() { return (size_t)kSmemBytes; }

cudaError_t launch_mesh(const MeshParams& p, const MeshTensorMaps& maps, int grid, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(gc_mesh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    gc_mesh_kernel<<<grid, NT, kSmemBytes, stream>>>(p, maps);
    return cudaGetLastError();
}

cudaError_t mesh_kernel_attributes(int* regs, int* static_smem, int* max_dyn_smem)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, gc_mesh_kernel);
    if (e != cudaSuccess) return e;
    if (regs) *regs = a.numRegs;
    if (static_smem) *static_smem = (int)a.sharedSizeBytes;
    if (max_dyn_smem) *max_dyn_smem = a.maxDynamicSharedSizeBytes;
    return cudaSuccess;
}

size_t halo_bytes(int size_x) { return (size_t)size_x * 4 * 8192 + (size_t)size_x * 32; }

cudaError_t launch_halo_copy(const uint16_t* blocks, const uint8_t* sun, const uint8_t* light, const uint8_t* surface,
                             int size_x, int cz, int z, uint8_t* buf, bool unpack, cudaStream_t stream)
{
    const int total = size_x * 4 * 512;
    int grid = (total + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    gc_halo_copy_kernel<<<grid, 256, 0, stream>>>(const_cast<uint16_t*>(blocks), const_cast<uint8_t*>(sun), const_cast<uint8_t*>(light),
                                                  const_cast<uint8_t*>(surface), size_x, cz, z, buf, unpack ? 1 : 0);
    return cudaGetLastError();
}
}  // namespace gc

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
#include <cstdlib>
#include "gcmesh_kernel.cuh"

namespace gc
{
namespace
{
constexpr int NT = kMeshThreads;          // 20 warps
constexpr int kWarps = NT / 32;
constexpr int kSlabs = kRowsZ;            // 34 z-slabs per chunk: tz = -1 .. 32
// P1: kTeams teams of 4 warps (2 class + 2 light, one lane per voxel row).  Team t converts slabs t, t+4, ... and
// stages them itself through its own two-slot ring: after a slab is converted the team syncs on a named barrier and
// one of its lanes issues the TMA loads of the slab after next into the slot just freed.  (A separate producer warp
// cost ~700 cycles per slab — try_wait ~140, arrive.expect_tx + UTMALDG ~360, measured — and paced the whole phase.)
constexpr int kTeams = 4, kTeamWarps = 4, kTeamStages = 2;
constexpr int kStages = kTeams * kTeamStages;              // 8
constexpr int kSlabWarps = kTeams * kTeamWarps;            // warps 0 .. 15
constexpr int kHaloWarp = kSlabWarps;                      // warp 16: the y-halo rows (ty = -1, 64), straight from global memory
constexpr int kFetchWarp = kHaloWarp + 1;                  // warp 17: prefetches the next chunk's id, position and surface rows
static_assert(kFetchWarp < kWarps, "warp roles");
constexpr int kYHaloTasks = 2 * kRowsZ;                    // 2 faces x 34 rows
constexpr int kHaloTasks = kHaloRows * 2;
constexpr int kHaloPerThread = (kHaloTasks + NT - 1) / NT;

// ---- one staging slot: a z-slab of the three arrays (TMA destinations 128-byte aligned) ----
constexpr int kStBlocks = 0;                       // [64][32] u16
constexpr int kStSun = 4096;                       // [64][32] u8
constexpr int kStLight = 6144;
constexpr int kStageBytes = 8192;
// P3 reuses the staging ring: quad list (u32 per quad) followed by the per-type ranks (u16 per quad)
constexpr int kOffTList = kMaxQuads * 4;
static_assert(kMaxQuads * 6 <= kStages * kStageBytes, "a whole chunk's quad list must fit the staging ring");

// ---- shared-memory carve-up (dynamic, 1024-aligned base) ----
constexpr int kOffStages = 0;
constexpr int kOffPlanes = kOffStages + kStages * kStageBytes;            // 65536
constexpr int kOffLt = kOffPlanes + kPlanes * kHaloRows * 4;              // centre bytes, then the x-halo bytes
constexpr int kOffHx = kOffLt + ((kLtBytes + kLtxBytes + 15) & ~15);
constexpr int kOffBv = kOffHx + ((kHaloRows * 2 + 15) & ~15);             // u32[2][132]: per layer, which rows' x = 0 / x = 31 voxel emits
constexpr int kSurfBytes = kRowsZ * 32 + 2 * ((kRowsZ * 4 + 15) & ~15);   // [34][32] centre-x surface rows + [34][4] per-octet maxima + minima
constexpr int kSurfMinOff = (kRowsZ * 4 + 15) & ~15;
constexpr int kOffSurf = kOffBv + ((2 * kBvWords * 4 + 15) & ~15);                          // two buffers: this chunk's and the prefetched next one's
constexpr int kOffLut = kOffSurf + 2 * kSurfBytes;
constexpr int kOffMat = kOffLut + 256 * 8;
constexpr int kOffCls = kOffMat + 256 * 8;
constexpr int kOffLayerCnt = kOffCls + 256;                               // u32[64]
constexpr int kOffLayerTypA = kOffLayerCnt + 256;                         // u32[64]  types 1,2 (16 bit each)
constexpr int kOffLayerTypB = kOffLayerTypA + 256;                        // u32[64]  types 3,4
constexpr int kOffLayerBase = kOffLayerTypB + 256;                        // u32[65]
constexpr int kOffLayerTypBaseA = kOffLayerBase + 272;
constexpr int kOffLayerTypBaseB = kOffLayerTypBaseA + 256;
constexpr int kOffFaces = kOffLayerTypBaseB + 256;                        // FaceDesc[6], 16 B each
constexpr int kOffBars = kOffFaces + 96;                                  // full[8]
constexpr int kOffCtx = kOffBars + kStages * 8;
constexpr int kOffNext = kOffCtx + 64;                                    // NextChunk[2]
constexpr int kSmemBytes = kOffNext + 32;
static_assert(kOffPlanes % 128 == 0 && kOffLt % 16 == 0 && kOffHx % 16 == 0 && kOffBv % 16 == 0 && kOffSurf % 16 == 0, "align");
static_assert(kOffLut % 8 == 0 && kOffMat % 8 == 0 && kOffBars % 8 == 0 && kOffLayerCnt % 4 == 0 && kOffCtx % 4 == 0, "align");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

static_assert(sizeof(FacePlan) == 16, "FacePlan");

struct NextChunk { int32_t chunk, cx, cy, cz; };   // chunk = index into coords, or -1 when the work is exhausted

struct ChunkCtx
{
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug raises the error flag and traps after ~4 s of wall time instead of hanging the GPU.
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* error_flag)
{
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity))
    {
        if ((++spins & 0xFFFu) == 0)
        {
            const unsigned long long t = global_timer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull)
            {
                if (error_flag) atomicExch(error_flag, 1);
                __trap();
            }
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
// TMA bulk copy (no tensor map): `bytes` contiguous bytes global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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

// In-kernel cycle attribution is compiled in only for the profiling instantiation (GCMESH_PROFILE=1).
#define GC_PROF_MARK(slot)                                                            \
    do {                                                                              \
        if (kProf && tid == 0) { const long long t_ = clock64(); atomicAdd(p.prof + (slot), (unsigned long long)(t_ - prof_t)); prof_t = t_; } \
    } while (0)

// =====================================================================================================
template <bool kProf>
__global__ void __launch_bounds__(NT, 1)
gc_mesh_kernel(const MeshParams p, const __grid_constant__ MeshTensorMaps maps)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint8_t* s_stage = smem + kOffStages;
    uint32_t* s_list = reinterpret_cast<uint32_t*>(smem + kOffStages);
    uint16_t* s_tlist = reinterpret_cast<uint16_t*>(smem + kOffStages + kOffTList);
    uint32_t* s_planes = reinterpret_cast<uint32_t*>(smem + kOffPlanes);
    uint8_t* s_planes_b = smem + kOffPlanes;
    uint8_t* s_lt = smem + kOffLt;
    uint16_t* s_hx = reinterpret_cast<uint16_t*>(smem + kOffHx);
    uint8_t* s_hx_b = smem + kOffHx;
    uint32_t* s_bvm = reinterpret_cast<uint32_t*>(smem + kOffBv);
    NextChunk* s_next = reinterpret_cast<NextChunk*>(smem + kOffNext);
    LutEntry* s_lut = reinterpret_cast<LutEntry*>(smem + kOffLut);
    MatEntry* s_mat = reinterpret_cast<MatEntry*>(smem + kOffMat);
    uint8_t* s_cls = smem + kOffCls;
    uint32_t* s_layer_cnt = reinterpret_cast<uint32_t*>(smem + kOffLayerCnt);
    uint32_t* s_layer_typa = reinterpret_cast<uint32_t*>(smem + kOffLayerTypA);
    uint32_t* s_layer_typb = reinterpret_cast<uint32_t*>(smem + kOffLayerTypB);
    uint32_t* s_layer_base = reinterpret_cast<uint32_t*>(smem + kOffLayerBase);
    uint32_t* s_layer_typbase_a = reinterpret_cast<uint32_t*>(smem + kOffLayerTypBaseA);
    uint32_t* s_layer_typbase_b = reinterpret_cast<uint32_t*>(smem + kOffLayerTypBaseB);
    FacePlan* s_faces = reinterpret_cast<FacePlan*>(smem + kOffFaces);
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
        s_faces[tid] = make_face_plan(table[tid]);
    }
    if (tid == 0)
    {
        for (int i = 0; i < kStages; i++) mbar_init(smem_u32(&s_bars[i]), 1);   // full: the issuing lane's expect_tx arrive
        fence_barrier_init();
        if (p.use_tma)
        {
            prefetch_tensormap(&maps.blocks_slab);
            prefetch_tensormap(&maps.sun_slab);
            prefetch_tensormap(&maps.light_slab);
        }
    }
    const int kill_cls = p.tables->cls8[p.tables->kill_zone & 255];
    const int air_cls = p.tables->cls8[0];
    __syncthreads();

    uint32_t team_seq = 0;  // slabs this warp's team has staged so far (ring slot / parity source)
    long long prof_t = clock64();

    // Claim the next chunk and stage what P1 needs before any voxel arrives: its position and the centre-x surface
    // rows of the three columns (cz-1, cz, cz+1) with their per-octet maxima.  Run by one warp into buffer `b`,
    // one chunk ahead of the rest of the CTA.
    auto fetch_chunk = [&](int b)
    {
        int c = 0, fx = 0, fy = 0, fz = 0;
        if (lane == 0)
        {
            c = atomicAdd(p.work_counter, 1);
            if (c >= p.n_chunks) c = -1;
            if (c >= 0) { const GcChunkCoord cc = p.coords[c]; fx = cc.cx; fy = cc.cy; fz = cc.cz; }
            s_next[b].chunk = c; s_next[b].cx = fx; s_next[b].cy = fy; s_next[b].cz = fz;
        }
        c = __shfl_sync(0xffffffffu, c, 0); fx = __shfl_sync(0xffffffffu, fx, 0); fz = __shfl_sync(0xffffffffu, fz, 0);
        if (c < 0) return;
        uint8_t* surf = smem + kOffSurf + b * kSurfBytes;
        for (int i = lane; i < kRowsZ * 4; i += 32)
        {
            const int r = i >> 2, q = i & 3;
            const int tz = r - 1;
            const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
            const size_t col = (size_t)(fz + dcz) * p.size_x + fx;
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p.surface + col * 1024 + (tz & 31) * 32) + q);
            reinterpret_cast<uint2*>(surf)[i] = v;
            const uint32_t m = max4(v.x, v.y);
            const uint32_t m2 = max4(m, m >> 16);
            surf[kRowsZ * 32 + i] = (uint8_t)max4(m2, m2 >> 8);
            const uint32_t n = max4(~v.x, ~v.y);              // per-byte min = ~max(~a, ~b)
            const uint32_t n2 = max4(n, n >> 16);
            surf[kRowsZ * 32 + kSurfMinOff + i] = (uint8_t)~max4(n2, n2 >> 8);
        }
    };

    int buf = 0;
    if (warp == kFetchWarp) fetch_chunk(0);
    if (tid == 0) { s_ctx->any_emit = 0; s_ctx->any_bv = 0; }
    if (tid < 2 * kBvWords) s_bvm[tid] = 0;

    for (;; buf ^= 1)
    {
        // ================= P0: the chunk prefetched during the previous chunk's P1 =================
        __syncthreads();
        const int chunk = s_next[buf].chunk;
        if (chunk < 0) break;
        const int cx = s_next[buf].cx, cy = s_next[buf].cy, cz = s_next[buf].cz;
        const int lwy = cy * kTY;
        const uint8_t* s_surf = smem + kOffSurf + buf * kSurfBytes;
        const uint8_t* s_surfmax = s_surf + kRowsZ * 32;

        GC_PROF_MARK(0);
        // ================= P1: stage + convert =================
        const bool has_lo = lwy > 0, has_hi = lwy + kTY < GC_WORLD_BLOCK_HEIGHT;
        const long long p1_t0 = kProf ? clock64() : 0;
        const int col0 = cz * p.size_x + cx;
        if (warp < kSlabWarps)
        {
            // ---- slab teams: one lane per voxel row (ty = 0..63); the team stages its own slabs ----
            // team = warp % 4 puts each team on its own scheduler partition (SMSP = warp % 4): two class + two light warps each
            const int team = warp % kTeams, wt = warp / kTeams;
            const bool is_class = wt < 2;
            const int ty = (wt & 1) * 32 + lane;
            const int wy = lwy + ty;
            const int rot4 = (lane >> 1) & 3, rot2 = (lane >> 2) & 1;   // shared-memory read rotations (bank-conflict free)
            const uint32_t unrot = rot_selector(rot4);
            const int boff0 = kStBlocks + ty * 64 + ((0 + rot4) & 3) * 16, boff1 = kStBlocks + ty * 64 + ((1 + rot4) & 3) * 16;
            const int boff2 = kStBlocks + ty * 64 + ((2 + rot4) & 3) * 16, boff3 = kStBlocks + ty * 64 + ((3 + rot4) & 3) * 16;
            const int n_mine = (kSlabs - team + kTeams - 1) / kTeams;  // slabs team, team + 4, ...

            // stage slab number k of this team (s = team + 4k) into the team's ring slot for sequence number `seq`
            auto issue = [&](int k, uint32_t seq)
            {
                const int s = team + k * kTeams;
                const int stage = team * kTeamStages + (int)(seq % kTeamStages);
                const int tz = s - 1;
                const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
                const int zz = tz & 31;
                const int slot = (col0 + dcz * p.size_x) * 4 + cy;
                const uint32_t full = smem_u32(&s_bars[stage]);
                const uint32_t d = smem_u32(s_stage + stage * kStageBytes);
                fence_proxy_async();   // the slot was read (and, as the quad list, written) by the generic proxy
                mbar_expect_tx(full, 8192u);
                tma_load_4d(d + kStBlocks, &maps.blocks_slab, 0, 0, zz, slot, full);
                tma_load_4d(d + kStSun, &maps.sun_slab, 0, 0, zz, slot, full);
                tma_load_4d(d + kStLight, &maps.light_slab, 0, 0, zz, slot, full);
                if (kProf && blockIdx.x == 0) p.prof[64 + s] = (unsigned long long)(clock64() - p1_t0);
            };
            auto stage_plain = [&](int k, uint32_t seq)   // debug staging path (GCMESH_NO_TMA=1): the whole warp copies
            {
                const int s = team + k * kTeams;
                const int stage = team * kTeamStages + (int)(seq % kTeamStages);
                const int tz = s - 1;
                const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
                const int slot = (col0 + dcz * p.size_x) * 4 + cy;
                const size_t vox = (size_t)slot * GC_CHUNK_SIZE_3 + (size_t)(tz & 31) * 2048;
                uint8_t* st = s_stage + stage * kStageBytes;
                for (int i = lane; i < 256; i += 32) reinterpret_cast<uint4*>(st + kStBlocks)[i] = reinterpret_cast<const uint4*>(p.blocks + vox)[i];
                for (int i = lane; i < 128; i += 32)
                {
                    reinterpret_cast<uint4*>(st + kStSun)[i] = reinterpret_cast<const uint4*>(p.sun + vox)[i];
                    reinterpret_cast<uint4*>(st + kStLight)[i] = reinterpret_cast<const uint4*>(p.light + vox)[i];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&s_bars[stage]));
            };

            // prologue: the first kTeamStages slabs (all slots are free: the previous chunk ended with a CTA-wide sync)
            if (wt == 0)
            {
                for (int k = 0; k < kTeamStages && k < n_mine; k++)
                {
                    if (p.use_tma) { if (lane == 0) issue(k, team_seq + k); }
                    else stage_plain(k, team_seq + k);
                }
            }
            for (int k = 0; k < n_mine; k++)
            {
                const int s = team + k * kTeams;
                const uint32_t seq = team_seq + k;
                const int stage = team * kTeamStages + (int)(seq % kTeamStages);
                {
                    const long long w0 = kProf ? clock64() : 0;
                    mbar_wait(smem_u32(&s_bars[stage]), (seq / kTeamStages) & 1, p.error_flag);
                    if (kProf && lane == 0 && warp == 0) atomicAdd(p.prof + 9, (unsigned long long)(clock64() - w0));
                    if (kProf && lane == 0 && warp == 2) atomicAdd(p.prof + 10, (unsigned long long)(clock64() - w0));
                    if (kProf && lane == 0 && warp == 0 && k == 0) atomicAdd(p.prof + 11, (unsigned long long)(clock64() - p1_t0));
                    if (kProf && lane == 0 && wt == 0 && blockIdx.x == 0) p.prof[128 + s] = (unsigned long long)(clock64() - p1_t0);
                }
                const long long c0 = kProf ? clock64() : 0;
                const uint8_t* st = s_stage + stage * kStageBytes;
                const int hr = hrow(ty, s - 1);
                if (is_class)
                {
                    Ids8 q[4];
                    {
                        const uint4 v0 = *reinterpret_cast<const uint4*>(st + boff0), v1 = *reinterpret_cast<const uint4*>(st + boff1);
                        const uint4 v2 = *reinterpret_cast<const uint4*>(st + boff2), v3 = *reinterpret_cast<const uint4*>(st + boff3);
                        q[0] = Ids8{ v0.x, v0.y, v0.z, v0.w }; q[1] = Ids8{ v1.x, v1.y, v1.z, v1.w };
                        q[2] = Ids8{ v2.x, v2.y, v2.z, v2.w }; q[3] = Ids8{ v3.x, v3.y, v3.z, v3.w };
                    }
                    uint32_t out[8];
                    p1_class_row(q, s_lut, out);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        out[j] = rotl_bytes(out[j], unrot);
                        s_planes[j * kHaloRows + hr] = out[j];
                    }
                    const uint32_t vis = (s >= 1 && s <= kTZ) ? row_emit_mask(out) : 0u;
                    if (vis) s_ctx->any_emit = 1;
                    if (vis & 0x80000001u)
                    {
                        s_ctx->any_bv = 1;
                        if (vis & 1u) atomicOr(&s_bvm[(ty + 1) * 2 + (s >> 5)], 1u << (s & 31));
                        if (vis >> 31) atomicOr(&s_bvm[kBvWords + (ty + 1) * 2 + (s >> 5)], 1u << (s & 31));
                    }
                }
                else
                {
                    const uint32_t smax = *reinterpret_cast<const uint32_t*>(s_surfmax + s * 4);
                    const uint32_t smin = *reinterpret_cast<const uint32_t*>(s_surfmax + kSurfMinOff + s * 4);
#pragma unroll
                    for (int j = 0; j < 2; j++)
                    {
                        const int h = j ^ rot2;
                        const uint4 sn = *reinterpret_cast<const uint4*>(st + kStSun + ty * 32 + h * 16);
                        const uint4 bl = *reinterpret_cast<const uint4*>(st + kStLight + ty * 32 + h * 16);
                        const uint4 sf = *reinterpret_cast<const uint4*>(s_surf + s * 32 + h * 16);
                        const Words4 o = p1_light_half(Words4{ sn.x, sn.y, sn.z, sn.w }, Words4{ bl.x, bl.y, bl.z, bl.w },
                                                       Words4{ sf.x, sf.y, sf.z, sf.w }, wy,
                                                       light_mode(wy, (smax >> (16 * h)) & 255u, (smin >> (16 * h)) & 255u),
                                                       light_mode(wy, (smax >> (16 * h + 8)) & 255u, (smin >> (16 * h + 8)) & 255u));
                        *reinterpret_cast<uint4*>(s_lt + hr * kLtStride + h * 16) = make_uint4(o.x, o.y, o.z, o.w);
                    }
                }
                if (kProf && lane == 0 && warp == 0) { atomicAdd(p.prof + 12, (unsigned long long)(clock64() - c0)); atomicAdd(p.prof + 13, 1ull); }
                if (kProf && lane == 0 && warp == 2) atomicAdd(p.prof + 14, (unsigned long long)(clock64() - c0));
                if (kProf && lane == 0 && wt == 0 && blockIdx.x == 0) p.prof[192 + s] = (unsigned long long)(clock64() - p1_t0);
                // refill: once all four warps are done with the slot, one of them stages the slab after next into it
                if (k + kTeamStages < n_mine)
                {
                    named_bar_sync(1 + team, kTeamWarps * 32);
                    if (wt == (k & 3))
                    {
                        if (p.use_tma) { if (lane == 0) issue(k + kTeamStages, seq + kTeamStages); }
                        else stage_plain(k + kTeamStages, seq + kTeamStages);
                    }
                }
            }
            team_seq += n_mine;
        }
        else if (warp == kHaloWarp)
        {
            // ---- y-halo rows: ty = -1 (face 0) and ty = 64 (face 1), tz = -1 .. 32, read straight from the neighbour chunks ----
            for (int t = lane; t < kYHaloTasks; t += 32)
            {
                const int face = t / kRowsZ, tz = t % kRowsZ - 1;
                const int ty = face ? kTY : -1;
                const int wy = lwy + ty;
                const int hr = hrow(ty, tz);
                uint32_t out[8];
                uint4 l0, l1;
                if (!(face ? has_hi : has_lo))
                {
                    p1_class_row_const(face ? air_cls : kill_cls, out);
                    const uint32_t c = face ? ~0u : 0u;
                    l0 = make_uint4(c, c, c, c); l1 = l0;
                }
                else
                {
                    const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
                    const int slot = (col0 + dcz * p.size_x) * 4 + cy + (face ? 1 : -1);
                    const size_t vox = (size_t)slot * GC_CHUNK_SIZE_3 + 32 * ((face ? 0 : 63) + 64 * (tz & 31));
                    const uint4* gb = reinterpret_cast<const uint4*>(p.blocks + vox);
                    const uint4* gs = reinterpret_cast<const uint4*>(p.sun + vox);
                    const uint4* gl = reinterpret_cast<const uint4*>(p.light + vox);
                    Ids8 q[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        const uint4 v = __ldg(gb + j);
                        q[j].x = v.x; q[j].y = v.y; q[j].z = v.z; q[j].w = v.w;
                    }
                    const uint4 sn0 = __ldg(gs), sn1 = __ldg(gs + 1), bl0 = __ldg(gl), bl1 = __ldg(gl + 1);
                    p1_class_row(q, s_lut, out);
                    const uint4 sf0 = *reinterpret_cast<const uint4*>(s_surf + (tz + 1) * 32), sf1 = *reinterpret_cast<const uint4*>(s_surf + (tz + 1) * 32 + 16);
                    const Words4 a = p1_light_half(Words4{ sn0.x, sn0.y, sn0.z, sn0.w }, Words4{ bl0.x, bl0.y, bl0.z, bl0.w },
                                                   Words4{ sf0.x, sf0.y, sf0.z, sf0.w }, wy, 0, 0);
                    const Words4 b2 = p1_light_half(Words4{ sn1.x, sn1.y, sn1.z, sn1.w }, Words4{ bl1.x, bl1.y, bl1.z, bl1.w },
                                                    Words4{ sf1.x, sf1.y, sf1.z, sf1.w }, wy, 0, 0);
                    l0 = make_uint4(a.x, a.y, a.z, a.w); l1 = make_uint4(b2.x, b2.y, b2.z, b2.w);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) s_planes[j * kHaloRows + hr] = out[j];
                *reinterpret_cast<uint4*>(s_lt + hr * kLtStride) = l0;
                *reinterpret_cast<uint4*>(s_lt + hr * kLtStride + 16) = l1;
            }
        }
        else if (warp == kFetchWarp) fetch_chunk(buf ^ 1);
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
                    // consecutive lanes walk up one (side, tz) column: neighbouring y rows share cache lines
                    const int side = task / kHaloRows, rem = task % kHaloRows;
                    const int tz = rem / kRowsY - 1, ty = rem % kRowsY - 1;
                    const int wy = lwy + ty;
                    if (any_bv && wy >= 0 && wy < GC_WORLD_BLOCK_HEIGHT && xhalo_needed(s_bvm + side * kBvWords, ty, tz))
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
                    const int side = task / kHaloRows, rem = task % kHaloRows;
                    const int tz = rem / kRowsY - 1, ty = rem % kRowsY - 1;
                    const int hr = hrow(ty, tz);
                    const int wy = lwy + ty;
                    if (wy < 0) p1_xhalo_cell(kill_cls, 0x00u, s_hx_b, s_lt, hr, side);
                    else if (wy >= GC_WORLD_BLOCK_HEIGHT) p1_xhalo_cell(air_cls, 0xFFu, s_hx_b, s_lt, hr, side);
                    else if ((need >> k) & 1u) p1_xhalo_cell(s_cls[id[k] & 255u], light_byte_of(sn[k], bl[k], sf[k], wy), s_hx_b, s_lt, hr, side);
                }
            }
        }
        __syncthreads();
        if (tid == 0) { s_ctx->any_emit = 0; s_ctx->any_bv = 0; }   // every thread has read them; P1 of the next chunk sets them again
        if (any_bv && tid < 2 * kBvWords) s_bvm[tid] = 0;

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
            uint32_t* idx_words = reinterpret_cast<uint32_t*>(p.index_arena) + (size_t)quad_base * 3;
            uint4* vtx = reinterpret_cast<uint4*>(p.vertex_arena) + (size_t)quad_base * 3;
            const uint16_t* chunk_blocks = p.blocks + ((size_t)(cz * p.size_x + cx) * 4 + cy) * GC_CHUNK_SIZE_3;

            // ---- 3a: rows -> ordered quad list + per-type ranks (the whole chunk fits the staging ring) ----
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
                    p3a_row(rf, layer, lane, rank0, trank, s_list, s_tlist);
                }
            }
            __syncthreads();
            GC_PROF_MARK(4);
            // ---- 3b: one thread per quad: its index words and its four vertices ----
            uint32_t id_next = 0;
            if ((uint32_t)tid < total)
            {
                const uint32_t e = s_list[tid];
                id_next = __ldg(chunk_blocks + ((e >> 3) & 31) + 32 * (((e >> 13) & 63) + 64 * ((e >> 8) & 31)));
            }
            for (uint32_t i = tid; i < total; i += NT)
            {
                const uint32_t e = s_list[i];
                const uint32_t id = id_next & 255u;
                if (i + NT < total)
                {
                    const uint32_t en = s_list[i + NT];   // block id of this thread's next quad: its latency hides behind this one
                    id_next = __ldg(chunk_blocks + ((en >> 3) & 31) + 32 * (((en >> 13) & 63) + 64 * ((en >> 8) & 31)));
                }
                uint32_t iw[3];
                quad_index_words(i, iw);
                uint32_t* ip = idx_words + (size_t)(s_ctx->type_off[(e >> 19) & 7] + s_tlist[i]) * 3;
                ip[0] = iw[0]; ip[1] = iw[1]; ip[2] = iw[2];
                uint32_t v[12];
                p3b_quad(e, s_faces[e & 7], s_lt, s_planes + P_OPAQUE * kHaloRows, s_hx, s_mat[id], v);
                uint4* dst = vtx + (size_t)i * 3;
                dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                dst[2] = make_uint4(v[8], v[9], v[10], v[11]);
            }
        }
        // the staging ring was used as the quad list by the generic proxy: order it before the next TMA writes
        fence_proxy_async();
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
// ---- sparse upload: the GPU pulls the non-zero brick arrays straight out of pinned host memory over PCIe ----
__global__ void __launch_bounds__(256)
gc_upload_kernel(const uint32_t* __restrict__ items, int n_items, const uint16_t* __restrict__ h_blocks, const uint8_t* __restrict__ h_sun,
                 const uint8_t* __restrict__ h_light, uint16_t* d_blocks, uint8_t* d_sun, uint8_t* d_light)
{
    // each CTA takes 16 KB pieces so that every SM keeps many 16-byte loads in flight on the link
    constexpr int kPiece = 16384;
    const long long total = (long long)n_items * 8;   // up to 8 pieces per array (blocks 128 KB = 8, light arrays 64 KB = 4)
    for (long long w = blockIdx.x; w < total; w += gridDim.x)
    {
        const uint32_t item = items[w >> 3];
        const int piece = (int)(w & 7);
        const uint32_t slot = item >> 2, arr = item & 3u;
        const size_t bytes = arr == 0 ? (size_t)GC_CHUNK_SIZE_3 * 2 : (size_t)GC_CHUNK_SIZE_3;
        if ((size_t)piece * kPiece >= bytes) continue;
        const uint8_t* src = arr == 0 ? reinterpret_cast<const uint8_t*>(h_blocks) + (size_t)slot * bytes
                                      : (arr == 1 ? h_sun : h_light) + (size_t)slot * bytes;
        uint8_t* dst = arr == 0 ? reinterpret_cast<uint8_t*>(d_blocks) + (size_t)slot * bytes : (arr == 1 ? d_sun : d_light) + (size_t)slot * bytes;
        const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)piece * kPiece);
        uint4* d4 = reinterpret_cast<uint4*>(dst + (size_t)piece * kPiece);
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = s4[threadIdx.x + i * 256];
#pragma unroll
        for (int i = 0; i < 4; i++) d4[threadIdx.x + i * 256] = v[i];
    }
}
}  // namespace

cudaError_t launch_upload_items(const uint32_t* items, int n_items, const uint16_t* h_blocks, const uint8_t* h_sun, const uint8_t* h_light,
                                uint16_t* d_blocks, uint8_t* d_sun, uint8_t* d_light, cudaStream_t stream)
{
    if (n_items <= 0) return cudaSuccess;
    // PCIe needs ~100 KB in flight (50 GB/s x 2 us); 16 KB per CTA, so a few dozen CTAs saturate the link and the other
    // SMs stay free for a mesh kernel running on another stream (its CTA needs a whole SM's registers)
    static const int ctas_env = getenv("GCMESH_PULL_CTAS") ? atoi(getenv("GCMESH_PULL_CTAS")) : 0;
    const int cap = ctas_env > 0 ? ctas_env : 32;
    long long pieces = (long long)n_items * 8;
    int grid = pieces < cap ? (int)pieces : cap;
    gc_upload_kernel<<<grid, 256, 0, stream>>>(items, n_items, h_blocks, h_sun, h_light, d_blocks, d_sun, d_light);
    return cudaGetLastError();
}

size_t mesh_smem_bytes() { return (size_t)kSmemBytes; }

cudaError_t launch_mesh(const MeshParams& p, const MeshTensorMaps& maps, int grid, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(gc_mesh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gc_mesh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (p.prof) gc_mesh_kernel<true><<<grid, NT, kSmemBytes, stream>>>(p, maps);
    else gc_mesh_kernel<false><<<grid, NT, kSmemBytes, stream>>>(p, maps);
    return cudaGetLastError();
}

cudaError_t mesh_kernel_attributes(int* regs, int* static_smem, int* max_dyn_smem)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, gc_mesh_kernel<false>);
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

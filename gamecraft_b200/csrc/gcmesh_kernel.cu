// gcmesh — the chunk-mesh kernel for sm_100a (B200).
//
// Two persistent CTAs per SM (448 threads = 14 warps, ~111 KB shared memory each); each CTA pulls chunks from a work counter
// and, per chunk:
//   P0  the next chunk's position and the surface rows of its three columns (one warp, a chunk ahead)
//   P1  the 34 z-slabs of BLOCK IDS (32x64 voxels, 4 KB, contiguous in the brick) are staged by TMA tensor-map box
//       loads (cp.async.bulk.tensor.4d) into warp-owned mbarrier slots and converted, two voxel rows per lane, into
//       the resident on-chip form: eight class bit planes over the chunk + 1-voxel halo
//   P1x the x = -1 / x = 32 halo classes that some face or corner test will actually read are gathered from the
//       neighbour columns; air chunks gather nothing and stop here — their light arrays are never touched
//   P2  face masks per row (bitwise, 32 voxels per op), per-row counts, per-layer warp reductions, scan,
//       ONE atomicAdd reserves the chunk's contiguous segment in the batch arenas
//   P3  per batch of layers: (a) rows -> ordered quad list + per-type ranks, (b) one thread per quad -> its 3x3
//       light cells gathered straight from the sunlight / blockLight arrays (only cells in front of drawn faces are
//       ever read, as in the reference's VertexLight) -> 4 vertices -> three 16-byte stores + its index words
// While one CTA of an SM computes (P2 / P3), the other streams (P1): the SM's issue slots and its share of HBM
// bandwidth are used at the same time.  The per-lane work is in mesh_core.cuh (shared with the host emulator used by
// the CPU tests).
//
// Reference path replaced: BuildChunkAsync / BuildBlock / VertexLight, Code/WorldRender.cpp:6-27,329-383,445-560.
#include <cstdlib>
#include "gcmesh_kernel.cuh"

namespace gc
{
namespace
{
constexpr int NT = kMeshThreads;          // 448 threads = 14 warps
constexpr int kWarps = NT / 32;
constexpr int kSlabs = kRowsZ;            // 34 z-slabs per chunk: tz = -1 .. 32
// P1: eight slab warps, each with its own ring slot and mbarrier.  Warp w converts slabs w, w + 8, ...: it waits for its
// slot, pulls the slab's 64 voxel rows into registers (two per lane), immediately re-arms the slot with the TMA load of
// its next slab, and converts the rows while that load is in flight.  (Measured: a lane's conversion is one long
// dependency chain, ~1400 cycles per row whatever the memory system does, so two rows per lane double the phase's
// throughput; a separate producer warp cost ~700 cycles per slab and paced the phase; a ring shared by independent
// producers breaks mbarrier parity waits.)
constexpr int kSlabWarps = 8;                              // warps 0 .. 7
constexpr int kSlabsPerWarp = kTZ / kSlabWarps;            // the 32 centre slabs; the two z-halo slabs go to the fetch warp
constexpr int kStages = kSlabWarps;                        // one slot per slab warp
constexpr int kHaloWarp = kSlabWarps;                      // warp 8: the y-halo rows (ty = -1, 64), straight from global memory
constexpr int kFetchWarp = kHaloWarp + 1;                  // warp 9: claims the next work item and reads the two z-halo slabs (warps 10 .. 13 join in from P1x on)
static_assert(kFetchWarp < kWarps, "warp roles");
constexpr int kYHaloTasks = 2 * kRowsZ;                    // 2 faces x 34 rows

constexpr int kStageBytes = 4096;                          // one z-slab of block ids: [64][32] u16

// ---- shared-memory carve-up (dynamic, 1024-aligned base) ----
// The staging ring is dead once P1 ends; P1x .. P3 reuse it for the x-halo classes and the quad list of a batch.
constexpr int kOffStages = 0;
constexpr int kOffHx = 0;                                                  // u16[2244]
constexpr int kOffList = (kHaloRows * 2 + 127) & ~127;                     // u32[kListCap] quad entries, then u16[kListCap] per-type ranks
constexpr int kListCap = ((kStages * kStageBytes - kOffList) / 6) & ~15;   // quads per P3 batch
constexpr int kOffTList = kOffList + kListCap * 4;
constexpr int kMaxLayerQuads = 4 * kTX * kTZ + 2 * (kTX + kTZ);            // a layer's faces: <= 2 per cell up/down + 1 per in-layer neighbour pair
static_assert(kListCap >= kMaxLayerQuads, "one layer must fit a batch");
constexpr int kOffPlanes = kOffStages + kStages * kStageBytes;            // 32768
constexpr int kOffBv = kOffPlanes + kPlanes * kHaloRows * 4;              // u32[2][132]: per layer, which rows' x = 0 / x = 31 voxel emits
constexpr int kSurfBytes = (kSurfTileBytes + 15) & ~15;                   // [34][34] surface heights under the halo tile
constexpr int kOffSurf = kOffBv + ((2 * kBvWords * 4 + 15) & ~15);        // staged at the start of P3, for chunks that emit
constexpr int kOffLut = kOffSurf + kSurfBytes;
constexpr int kOffMat = kOffLut + 256 * 8;
constexpr int kOffCls = kOffMat + 256 * 8;
constexpr int kOffLayerCnt = kOffCls + 256;                               // u32[64]
constexpr int kOffLayerTypA = kOffLayerCnt + 256;                         // u32[64]  types 1,2 (16 bit each)
constexpr int kOffLayerTypB = kOffLayerTypA + 256;                        // u32[64]  types 3,4
constexpr int kOffLayerBase = kOffLayerTypB + 256;                        // u32[65]
constexpr int kOffLayerTypBaseA = kOffLayerBase + 272;
constexpr int kOffLayerTypBaseB = kOffLayerTypBaseA + 256;
constexpr int kOffFaces = kOffLayerTypBaseB + 256;                        // FacePlan[6], 48 B each
constexpr int kOffBars = kOffFaces + 6 * 48;                              // full[8]
constexpr int kOffCtx = kOffBars + kStages * 8;
constexpr int kOffNext = kOffCtx + 128;                                    // NextChunk[2]
constexpr int kOffTyp32 = kOffNext + 32;                                   // u32[4][64]: uncapped mode, first quad of mesh types 1..4 per layer
constexpr int kSmemBytes = kOffTyp32 + 4 * 64 * 4;
static_assert(kOffPlanes % 128 == 0 && kOffBv % 16 == 0 && kOffSurf % 16 == 0 && kOffList % 16 == 0 && kOffTList % 2 == 0, "align");
static_assert(kOffLut % 8 == 0 && kOffMat % 8 == 0 && kOffBars % 8 == 0 && kOffLayerCnt % 4 == 0 && kOffCtx % 4 == 0 && kOffFaces % 16 == 0, "align");
static_assert(2 * (kSmemBytes + 1024) <= 228 * 1024, "two CTAs per SM");

static_assert(sizeof(FacePlan) == 48, "FacePlan");

struct NextChunk { int32_t chunk, cx, cy, cz; };   // chunk = index into coords, -1 when the work is exhausted, kPushItem: halo push task cx
constexpr int kPushItem = -2;

struct ChunkCtx
{
    int32_t any_emit;     // some voxel of the chunk emits quads (else P1x / P2 / P3 are skipped)
    int32_t any_bv;       // some emitting voxel sits at x = 0 or x = 31 (else no x-halo cell is ever read)
    int32_t emit;         // P3 runs
    uint32_t quad_base;   // arena segment (quads)
    uint32_t total;
    uint32_t type_off[kMeshTypes];
    uint8_t slab_cls[kSlabs + 2];   // per z-slab: 0xFF = its plane rows are written; else the class byte of the one non-emitting block
                                    // (air, as a rule) that fills it — such rows are only written if the chunk turns out to need them
    int32_t halo_ok;      // fused exchange: result of a wait for a neighbour, broadcast to the CTA
    uint32_t halo_seen, halo_failed;   // fused exchange, bit d: the plane from neighbour d is known to have arrived / was given up on
    uint32_t batch_base[kMeshTypes];   // uncapped mode: first quad of each mesh type in the current P3 batch (the list holds ranks relative to it)
};
static_assert(sizeof(ChunkCtx) <= 128, "ChunkCtx outgrew its slot");

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

// ---- fused slab halo exchange: one push task ----
// Task t of direction d (d = the t / tasks_per_dir-th connected neighbour) stores planes [part * P / T, (part + 1) * P / T) of
// this rank's boundary row — P = size_x * 4 planes of 8 KB: blocks | sunlight | blockLight of one chunk at z = 0 (down) or
// z = 31 (up) — and the surface rows of the same share of columns into the neighbour's halo row, with 16-byte stores over
// NVLink.  The task that completes a direction publishes "arrived" in the neighbour's flag block behind a system-scope fence.
__device__ void halo_push_task(const HaloFused& h, int task, int* s_flag)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    if (h.debug & 8) return;
    const int T = h.tasks_per_dir;
    int d = task / T;
    const int part = task % T;
    if (!h.peer[0].connected) d = 1;                       // one neighbour only: every task is for it
    const HaloPeer& peer = h.peer[d];
    // the neighbour must have entered this exchange: until then its previous kernel may still read its halo rows
    if (tid == 0) *s_flag = ((h.debug & 4) || halo_wait_at_least(h.flags, d == 0 ? HF_READY_FROM_DOWN : HF_READY_FROM_UP, h.serial, h.wait_ns, peer.flags)) ? 1 : 0;
    __syncthreads();
    const bool ready = *s_flag != 0;
    __syncthreads();
    if (!ready) return;                                    // (both flag blocks carry the error; no arrival is published)
    const int z = d == 0 ? 0 : 31;
    const int src_row = h.send_row[d];
    const int planes = h.size_x * 4;
    const int p0 = planes * part / T, p1 = planes * (part + 1) / T;      // (planes * T stays far below 2^31: size_x <= 16 384 columns)
    const int per_plane = 8192 / 16;                       // uint4 per plane: 256 blocks, 128 sunlight, 128 blockLight
    if (!(h.debug & 1))
    for (int i = p0 * per_plane + tid; i < p1 * per_plane; i += nt)
    {
        const int plane = i / per_plane, q = i % per_plane;
        const int cx = plane >> 2, cy = plane & 3;
        const size_t src = (((size_t)src_row * h.size_x + cx) * 4 + cy) * GC_CHUNK_SIZE_3 + (size_t)z * 2048;
        const size_t dst = (((size_t)peer.row * h.size_x + cx) * 4 + cy) * GC_CHUNK_SIZE_3 + (size_t)z * 2048;
        if (q < 256) reinterpret_cast<uint4*>(peer.blocks + dst)[q] = reinterpret_cast<const uint4*>(h.blocks + src)[q];
        else if (q < 384) reinterpret_cast<uint4*>(peer.sun + dst)[q - 256] = reinterpret_cast<const uint4*>(h.sun + src)[q - 256];
        else reinterpret_cast<uint4*>(peer.light + dst)[q - 384] = reinterpret_cast<const uint4*>(h.light + src)[q - 384];
    }
    const int c0 = (p0 + 3) >> 2, c1 = (p1 + 3) >> 2;      // columns whose first plane lies in this share
    if (!(h.debug & 1))
    for (int i = c0 * 8 + tid; i < c1 * 8; i += nt)
    {
        const int cx = i >> 3, q = i & 7;
        reinterpret_cast<uint32_t*>(peer.surface + ((size_t)peer.row * h.size_x + cx) * 1024 + z * 32)[q] =
            reinterpret_cast<const uint32_t*>(h.surface + ((size_t)src_row * h.size_x + cx) * 1024 + z * 32)[q];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0)
    {
        uint32_t* ticket = h.flags + (d == 0 ? HF_TICKET : HF_TICKET_UP);
        if (atomicAdd(ticket, 1u) == (uint32_t)T - 1)
        {
            *ticket = 0;
            __threadfence_system();
            halo_store_release_sys(peer.flags + (d == 0 ? HF_ARRIVED_FROM_UP : HF_ARRIVED_FROM_DOWN), h.serial);
        }
    }
}

// In-kernel cycle attribution is compiled in only for the profiling instantiation (GCMESH_PROFILE=1).
#define GC_PROF_MARK(slot)                                                            \
    do {                                                                              \
        if (kProf && tid == 0) { const long long t_ = clock64(); atomicAdd(p.prof + (slot), (unsigned long long)(t_ - prof_t)); prof_t = t_; } \
    } while (0)

// =====================================================================================================
// kWide: the uncapped ("stress") form — no 10 000-quad cap, 32-bit indices (not a reference behaviour)
template <bool kProf, bool kHalo, bool kWide>
__global__ void __launch_bounds__(NT, 2)
gc_mesh_kernel(const MeshParams p, const __grid_constant__ MeshTensorMaps maps)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint8_t* s_stage = smem + kOffStages;
    uint32_t* s_list = reinterpret_cast<uint32_t*>(smem + kOffList);
    uint16_t* s_tlist = reinterpret_cast<uint16_t*>(smem + kOffTList);
    uint32_t* s_planes = reinterpret_cast<uint32_t*>(smem + kOffPlanes);
    uint16_t* s_hx = reinterpret_cast<uint16_t*>(smem + kOffHx);
    uint8_t* s_hx_b = smem + kOffHx;
    uint32_t* s_bvm = reinterpret_cast<uint32_t*>(smem + kOffBv);
    uint8_t* s_surf = smem + kOffSurf;
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
    uint32_t* s_typ32 = reinterpret_cast<uint32_t*>(smem + kOffTyp32);

    // Programmatic dependent launch: the next launch on the stream may start its CTAs as soon as SM slots come free (this grid is
    // wholly resident from the start, so nothing of it waits for a slot) ...
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // ---- one-time setup (block tables never change while a launch is in flight: gcmesh_set_block_table synchronises) ----
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
        if (p.use_tma) prefetch_tensormap(&maps.blocks_slab);
    }
    const int kill_cls = p.tables->cls8[p.tables->kill_zone & 255];
    const int air_cls = p.tables->cls8[0];
    __syncthreads();
    // ... and the set-up above ran beside the previous launch's tail: from here on that launch is complete and its writes — the
    // control block its last CTA reset, meshes, whatever earlier work on the stream left in the window — are visible
    asm volatile("griddepcontrol.wait;" ::: "memory");

    uint32_t warp_seq = 0;  // slabs this warp has staged so far (mbarrier parity source)
    long long prof_t = clock64();

    // Claim the next chunk, a chunk ahead of the rest of the CTA (the fetch warp, into buffer `b`).
    // With p.trust_surface (the window's surface maps are ComputeSurface of its block ids, Lighting.cpp:5-26) a brick that lies
    // wholly at or above its column's highest surface entry is AIR by definition: its empty result is written right here and
    // the warp claims again, so such a chunk costs one 1 KB surface read instead of 128 KB of block ids and never reaches the
    // CTA's phases.  (Layer y = 255 is outside ComputeSurface's walk, Lighting.cpp:9: the top brick's last layer is looked at.)
    auto fetch_chunk = [&](int b)
    {
        int c = -1, fx = 0, fy = 0, fz = 0;
        for (;;)
        {
            if (lane == 0)
            {
                c = atomicAdd(p.work_counter, 1); fx = 0; fy = 0; fz = 0;
                if (kHalo && c < p.n_push) { fx = c; c = kPushItem; }          // a halo push task (fused exchange): its number travels in cx
                else
                {
                    if (kHalo) c -= p.n_push;
                    if (c >= p.n_chunks) c = -1;
                    else if (p.order) c = p.order[c];
                    if (c >= 0) { const GcChunkCoord cc = p.coords[c]; fx = cc.cx; fy = cc.cy; fz = cc.cz; }
                }
            }
            if (!p.trust_surface) break;
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c < 0) break;
            fx = __shfl_sync(0xffffffffu, fx, 0); fy = __shfl_sync(0xffffffffu, fy, 0); fz = __shfl_sync(0xffffffffu, fz, 0);
            const size_t col = (size_t)fz * p.size_x + fx;
            const uint4* sp = reinterpret_cast<const uint4*>(p.surface + col * 1024) + lane * 2;
            const uint4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
            uint32_t m = __vmaxu4(__vmaxu4(__vmaxu4(s0.x, s0.y), __vmaxu4(s0.z, s0.w)), __vmaxu4(__vmaxu4(s1.x, s1.y), __vmaxu4(s1.z, s1.w)));
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) m = __vmaxu4(m, __shfl_xor_sync(0xffffffffu, m, d));
            const int top = (int)max(max(m & 255u, (m >> 8) & 255u), max((m >> 16) & 255u, m >> 24));   // highest block of the column + 1
            bool air = fy * kTY >= top;
            if (air && fy == GC_WORLD_CHUNK_HEIGHT - 1)
            {
                const uint4* bp = reinterpret_cast<const uint4*>(p.blocks + (col * 4 + fy) * GC_CHUNK_SIZE_3 + 32 * (63 + 64 * lane));   // row z = lane of layer y = 255
                const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1), b2 = __ldg(bp + 2), b3 = __ldg(bp + 3);
                const uint32_t any = b0.x | b0.y | b0.z | b0.w | b1.x | b1.y | b1.z | b1.w | b2.x | b2.y | b2.z | b2.w | b3.x | b3.y | b3.z | b3.w;
                air = !__any_sync(0xffffffffu, any != 0u);
            }
            if (!air) break;
            if (lane == 0)
            {
                GcChunkOut o;
                o.quad_base = 0; o.vert_count = 0; o.quads = 0; o.flags = 0; o.reserved[0] = 0; o.reserved[1] = 0;
                for (int t = 0; t < kMeshTypes; t++) { o.index_count[t] = 0; o.index_offset[t] = 0; }
                p.out[c] = o;
                if (p.vert_table) p.vert_table[(col * 4 + fy)] = 0;
            }
        }
        if (lane == 0) { s_next[b].chunk = c; s_next[b].cx = fx; s_next[b].cy = fy; s_next[b].cz = fz; }
    };

    // fused exchange: tell both neighbours that this rank has entered exchange `serial` — its previous kernel, the last reader
    // of its halo rows, has finished (stream order), so they may be overwritten
    if (kHalo && blockIdx.x == 0 && tid < 2 && p.halo.peer[tid].connected && !(p.halo.debug & 4))
        halo_store_release_sys(p.halo.peer[tid].flags + (tid == 0 ? HF_READY_FROM_UP : HF_READY_FROM_DOWN), p.halo.serial);
    if (kHalo && tid == 0) { s_ctx->halo_seen = 0; s_ctx->halo_failed = 0; }   // (kept in shared memory: two registers less in the chunk loop)

    int buf = 0;
    if (warp == kFetchWarp) fetch_chunk(0);
    if (tid == 0) { s_ctx->any_emit = 0; s_ctx->any_bv = 0; }
    if (tid < 2 * kBvWords) s_bvm[tid] = 0;

    if (kHalo)
    {
        // ---- halo push tasks: a share of this rank's boundary planes goes into a neighbour's halo row over peer memory ----
        // They are the first items of the work counter, so a CTA's claims run through them before its first chunk: they are
        // worked off here, ahead of the chunk loop (whose registers they then do not compete for).
        for (;;)
        {
            __syncthreads();
            if (s_next[0].chunk != kPushItem) break;
            const int task = s_next[0].cx;
            __syncthreads();                                      // (everyone has read the item: the fetch warp may claim the next one)
            if (warp == kFetchWarp) fetch_chunk(0);
            halo_push_task(p.halo, task, &s_ctx->halo_ok);
        }
    }
    for (;; buf ^= 1)
    {
        // ================= P0: the chunk prefetched during the previous chunk's P1 =================
        __syncthreads();
        const int chunk = s_next[buf].chunk;
        if (chunk == -1) break;
        const int cx = s_next[buf].cx, cy = s_next[buf].cy, cz = s_next[buf].cz;
        const int lwy = cy * kTY;
        if (kHalo)
        {
            // a chunk of the first / last owned row reads the neighbour's plane: wait (once per CTA and side) until it has arrived
            uint32_t need = ((cz == p.halo.send_row[0] && p.halo.peer[0].connected) ? 1u : 0u) | ((cz == p.halo.send_row[1] && p.halo.peer[1].connected) ? 2u : 0u);
            need &= ~s_ctx->halo_seen;                      // (tid 0 updates it only behind the barrier below: every thread has read it by then)
            if (p.halo.debug & 2) need = 0;
            if (need)
            {
                int* s_halo_ok = &s_ctx->halo_ok;
                if (tid == 0)
                {
                    bool ok = (need & s_ctx->halo_failed) == 0;   // (a plane this CTA already gave up on is not waited for again)
                    if (ok && (need & 1u)) ok = halo_wait_at_least(p.halo.flags, HF_ARRIVED_FROM_DOWN, p.halo.serial, p.halo.wait_ns) && ok;
                    if (ok && (need & 2u)) ok = halo_wait_at_least(p.halo.flags, HF_ARRIVED_FROM_UP, p.halo.serial, p.halo.wait_ns);
                    // the planes were written by another GPU (generic proxy); this CTA reads them through TMA as well
                    asm volatile("fence.proxy.async;" ::: "memory");
                    *s_halo_ok = ok ? 1 : 0;
                }
                __syncthreads();
                const bool ok = *s_halo_ok != 0;
                if (tid == 0) { if (ok) s_ctx->halo_seen |= need; else s_ctx->halo_failed |= need; }
                __syncthreads();
                if (!ok)
                {
                    // the neighbour never delivered: the chunk is flagged, nothing is emitted
                    if (warp == kFetchWarp) fetch_chunk(buf ^ 1);
                    if (tid == 0)
                    {
                        GcChunkOut o;
                        o.quad_base = 0; o.vert_count = 0; o.quads = 0; o.flags = GC_CHUNK_HALO_MISSING; o.reserved[0] = 0; o.reserved[1] = 0;
                        for (int t = 0; t < kMeshTypes; t++) { o.index_count[t] = 0; o.index_offset[t] = 0; }
                        p.out[chunk] = o;
                        atomicExch(p.error_flag, 2);
                    }
                    continue;
                }
            }
        }

        GC_PROF_MARK(0);
        // ================= P1: stage + convert the block ids =================
        const bool has_lo = lwy > 0, has_hi = lwy + kTY < GC_WORLD_BLOCK_HEIGHT;
        const int col0 = cz * p.size_x + cx;
        // two voxel rows of slab s (ty = lane, lane + 32; the octets rotated as `unrot` undoes) -> their plane words
        // one voxel row (ty, slab s; the octets rotated as `unrot` undoes) -> its plane words; returns the row's emitting voxels
        auto convert_row = [&](const Ids8 (&q)[4], int u, uint32_t unrot, int ty, int s) -> uint32_t
        {
            uint32_t o[8];
            p1_class_row(q, u, s_lut, o);
            if (u != 2)   // (all-zero / all-one words need no un-rotation)
            {
#pragma unroll
                for (int j = 0; j < 8; j++) o[j] = rotl_bytes(o[j], unrot);
            }
            const int hr = hrow(ty, s - 1);
            store_row_planes(s_planes, hr, o);
            return row_emit_mask(o);
        };
        // two voxel rows of slab s (ty = lane, lane + 32)
        auto convert_rows = [&](const Ids8 (&qa)[4], const Ids8 (&qb)[4], int ua, int ub, uint32_t unrot, int s)
        {
            {
                // The whole slab one non-emitting block (the air above the terrain): nothing is written now.  Three quarters of a
                // terrain world's chunks are nothing but such slabs and are never looked at again; if the chunk does emit, the
                // rows are filled in after P1.
                const uint32_t id0 = __shfl_sync(0xffffffffu, qa[0].x, 0);
                const uint32_t cls = s_cls[id0 & 255u];
                const bool skip = __all_sync(0xffffffffu, ua == 2 && ub == 2 && qa[0].x == id0 && qb[0].x == id0) && (cls >> P_M0) == (uint32_t)kNoEmitType;
                if (lane == 0) s_ctx->slab_cls[s] = skip ? (uint8_t)cls : (uint8_t)0xFF;
                if (skip) return;
            }
            const uint32_t va = convert_row(qa, ua, unrot, lane, s), vb = convert_row(qb, ub, unrot, lane + 32, s);
            if (s >= 1 && s <= kTZ)
            {
                if (va | vb) s_ctx->any_emit = 1;
                if ((va | vb) & 0x80000001u)
                {
                    s_ctx->any_bv = 1;
                    if (va & 1u) atomicOr(&s_bvm[(lane + 1) * 2 + (s >> 5)], 1u << (s & 31));
                    if (va >> 31) atomicOr(&s_bvm[kBvWords + (lane + 1) * 2 + (s >> 5)], 1u << (s & 31));
                    if (vb & 1u) atomicOr(&s_bvm[(lane + 33) * 2 + (s >> 5)], 1u << (s & 31));
                    if (vb >> 31) atomicOr(&s_bvm[kBvWords + (lane + 33) * 2 + (s >> 5)], 1u << (s & 31));
                }
            }
        };
        if (warp < kSlabWarps)
        {
            // ---- slab warps: warp w converts the centre slabs tz = w, w + 8, w + 16, w + 24 out of its own ring slot; a lane
            // holds two voxel rows (ty = lane, lane + 32) so that two independent dependency chains are in flight ----
            const int rot4 = (lane >> 1) & 3;   // shared-memory read rotation (bank-conflict free)
            const uint32_t unrot = rot_selector(rot4);
            const int boff0 = lane * 64 + ((0 + rot4) & 3) * 16, boff1 = lane * 64 + ((1 + rot4) & 3) * 16;
            const int boff2 = lane * 64 + ((2 + rot4) & 3) * 16, boff3 = lane * 64 + ((3 + rot4) & 3) * 16;
            const uint32_t full = smem_u32(&s_bars[warp]);
            uint8_t* const st = s_stage + warp * kStageBytes;
            const int own_slot = col0 * 4 + cy;

            auto issue = [&](int k)   // stage this warp's slab number k into its slot
            {
                // (no proxy fence here: the slot's last generic accesses were reads that have completed — the refill follows them in
                // this lane's instruction stream — and the generic writes of P1x .. P3 were fenced by every thread at the end of
                // the previous chunk)
                mbar_expect_tx(full, (uint32_t)kStageBytes);
                tma_load_4d(smem_u32(st), &maps.blocks_slab, 0, 0, warp + k * kSlabWarps, own_slot, full);
            };
            auto stage_plain = [&](int k)   // debug staging path (GCMESH_NO_TMA=1): the whole warp copies
            {
                const size_t vox = (size_t)own_slot * GC_CHUNK_SIZE_3 + (size_t)(warp + k * kSlabWarps) * 2048;
                for (int i = lane; i < 256; i += 32) reinterpret_cast<uint4*>(st)[i] = reinterpret_cast<const uint4*>(p.blocks + vox)[i];
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
            };

            // the slot is free: the previous chunk ended with a CTA-wide sync
            if (p.use_tma) { if (lane == 0) issue(0); }
            else stage_plain(0);
            for (int k = 0; k < kSlabsPerWarp; k++)
            {
                const int s = 1 + warp + k * kSlabWarps;
                const long long w0 = kProf ? clock64() : 0;
                mbar_wait(full, (warp_seq + k) & 1, p.error_flag);
                const long long w1 = kProf ? clock64() : 0;
                if (kProf && warp == 0 && lane == 0) { atomicAdd(p.prof + 8, (unsigned long long)(w1 - w0)); atomicAdd(p.prof + 10, 1ull); }
                Ids8 qa[4], qb[4];
                {
                    const uint4 a0 = *reinterpret_cast<const uint4*>(st + boff0), a1 = *reinterpret_cast<const uint4*>(st + boff1);
                    const uint4 a2 = *reinterpret_cast<const uint4*>(st + boff2), a3 = *reinterpret_cast<const uint4*>(st + boff3);
                    const uint4 b0 = *reinterpret_cast<const uint4*>(st + 2048 + boff0), b1 = *reinterpret_cast<const uint4*>(st + 2048 + boff1);
                    const uint4 b2 = *reinterpret_cast<const uint4*>(st + 2048 + boff2), b3 = *reinterpret_cast<const uint4*>(st + 2048 + boff3);
                    qa[0] = Ids8{ a0.x, a0.y, a0.z, a0.w }; qa[1] = Ids8{ a1.x, a1.y, a1.z, a1.w };
                    qa[2] = Ids8{ a2.x, a2.y, a2.z, a2.w }; qa[3] = Ids8{ a3.x, a3.y, a3.z, a3.w };
                    qb[0] = Ids8{ b0.x, b0.y, b0.z, b0.w }; qb[1] = Ids8{ b1.x, b1.y, b1.z, b1.w };
                    qb[2] = Ids8{ b2.x, b2.y, b2.z, b2.w }; qb[3] = Ids8{ b3.x, b3.y, b3.z, b3.w };
                }
                int ua = row_uniformity(qa), ub = row_uniformity(qb);   // read all thirty-two words: both rows are in registers from here on
                // refill: the instructions above consumed every loaded register before this point of the (in-order) instruction
                // stream, so the slot is free — one lane stages the warp's next slab into it while the rows are converted
                asm volatile("" : "+r"(ua), "+r"(ub)::"memory");
                if (k + 1 < kSlabsPerWarp)
                {
                    __syncwarp();
                    if (p.use_tma) { if (lane == 0) issue(k + 1); }
                    else stage_plain(k + 1);
                }
                convert_rows(qa, qb, ua, ub, unrot, s);
                if (kProf && warp == 0 && lane == 0) atomicAdd(p.prof + 9, (unsigned long long)(clock64() - w1));
            }
            warp_seq += kSlabsPerWarp;
        }
        else if (warp == kHaloWarp)
        {
            // ---- y-halo rows: ty = -1 (face 0) and ty = 64 (face 1), tz = -1 .. 32, read straight from the neighbour chunks ----
            for (int t = lane; t < kYHaloTasks; t += 32)
            {
                const int face = t / kRowsZ, tz = t % kRowsZ - 1;
                const int ty = face ? kTY : -1;
                const int hr = hrow(ty, tz);
                uint32_t out[8];
                if (!(face ? has_hi : has_lo)) p1_class_row_const(face ? air_cls : kill_cls, out);
                else
                {
                    const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
                    const int slot = (col0 + dcz * p.size_x) * 4 + cy + (face ? 1 : -1);
                    const size_t vox = (size_t)slot * GC_CHUNK_SIZE_3 + 32 * ((face ? 0 : 63) + 64 * (tz & 31));
                    const uint4* gb = reinterpret_cast<const uint4*>(p.blocks + vox);
                    Ids8 q[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        const uint4 v = __ldg(gb + j);
                        q[j].x = v.x; q[j].y = v.y; q[j].z = v.z; q[j].w = v.w;
                    }
                    p1_class_row(q, row_uniformity(q), s_lut, out);
                }
                store_row_planes(s_planes, hr, out);
            }
        }
        else if (warp == kFetchWarp)
        {
            // ---- the two z-halo slabs (tz = -1, 32): contiguous 4 KB of the bricks in front / behind, two rows per lane straight
            // from global memory; then the next chunk's claim (which may pass over several bricks that are known to be air) ----
            if (!p.trust_surface) fetch_chunk(buf ^ 1);
#pragma unroll 1
            for (int side = 0; side < 2; side++)
            {
                const int slot = (col0 + (side ? p.size_x : -p.size_x)) * 4 + cy;
                const uint4* g = reinterpret_cast<const uint4*>(p.blocks + (size_t)slot * GC_CHUNK_SIZE_3 + (size_t)(side ? 0 : 31) * 2048) + lane * 4;
                Ids8 qa[4], qb[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    const uint4 a = __ldg(g + j), b = __ldg(g + 128 + j);
                    qa[j] = Ids8{ a.x, a.y, a.z, a.w }; qb[j] = Ids8{ b.x, b.y, b.z, b.w };
                }
                convert_rows(qa, qb, row_uniformity(qa), row_uniformity(qb), rot_selector(0), side ? kSlabs - 1 : 0);
            }
            if (p.trust_surface) fetch_chunk(buf ^ 1);
        }
        __syncthreads();
        GC_PROF_MARK(1);

        // ================= P1x: the x-halo classes that will be read (the staging ring is free from here on) =================
        const bool any_emit = s_ctx->any_emit != 0, any_bv = s_ctx->any_bv != 0;
        if (any_emit)
        {
            // the plane rows of the slabs P1 skipped (one non-emitting block throughout): constant words
            for (int t = tid; t < kSlabs * kTY; t += NT)
            {
                const int s = t / kTY, ty = t % kTY;
                const uint32_t cls = s_ctx->slab_cls[s];
                if (cls != 0xFFu)
                {
                    uint32_t o[8];
                    p1_class_row_const(cls, o);
                    store_row_planes(s_planes, hrow(ty, s - 1), o);
                }
            }
            // which halo cells can be read at all: per (side, layer) the rows next to an emitting boundary voxel
            uint64_t* s_need = reinterpret_cast<uint64_t*>(smem + kOffList);   // [2][66], the quad list's space (not yet in use)
            if (tid < 2 * kRowsY) s_need[tid] = any_bv ? xhalo_needed_mask(s_bvm + (tid / kRowsY) * kBvWords, tid % kRowsY - 1) : 0ull;
            __syncthreads();
            // a warp takes whole (side, tz) columns of halo cells: the column's brick row is computed once, lanes walk up
            // its 66 rows in three passes (three loads in flight per lane), neighbouring rows share cache lines
#pragma unroll 1
            for (int c = warp; c < 2 * kRowsZ; c += kWarps)
            {
                const int side = c / kRowsZ, tz = c % kRowsZ - 1;
                const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
                const size_t col = (size_t)(cz + dcz) * p.size_x + (cx + (side ? 1 : -1));
                const uint16_t* src = p.blocks + col * 4 * GC_CHUNK_SIZE_3 + (side ? 0 : 31) + 32 * 64 * (tz & 31);
                const uint64_t* need_rows = s_need + side * kRowsY;
                uint32_t id[3];
                uint32_t need = 0;
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const int r = lane + 32 * k, wy = lwy + r - 1;
                    if (r < kRowsY && wy >= 0 && wy < GC_WORLD_BLOCK_HEIGHT && ((need_rows[r] >> (tz + 1)) & 1ull))
                    {
                        need |= 1u << k;
                        id[k] = __ldg(src + (size_t)(wy >> 6) * GC_CHUNK_SIZE_3 + 32 * (wy & 63));
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const int r = lane + 32 * k, wy = lwy + r - 1;
                    if (r < kRowsY)
                    {
                        const int hr = r * kRowsZ + tz + 1;
                        if (wy < 0) p1_xhalo_cell(kill_cls, s_hx_b, hr, side);
                        else if (wy >= GC_WORLD_BLOCK_HEIGHT) p1_xhalo_cell(air_cls, s_hx_b, hr, side);
                        else if ((need >> k) & 1u) p1_xhalo_cell(s_cls[id[k] & 255u], s_hx_b, hr, side);
                    }
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
            // nothing in the chunk can emit (all AIR / invisible): empty result, no scan, no emission, no light read
            if (tid == 0)
            {
                GcChunkOut o;
                o.quad_base = 0; o.vert_count = 0; o.quads = 0; o.flags = 0; o.reserved[0] = 0; o.reserved[1] = 0;
                for (int t = 0; t < kMeshTypes; t++) { o.index_count[t] = 0; o.index_offset[t] = 0; }
                p.out[chunk] = o;
                if (p.vert_table) { const GcChunkCoord c = p.coords[chunk]; p.vert_table[(c.cz * p.size_x + c.cx) * 4 + c.cy] = 0; }
            }
            GC_PROF_MARK(3);
            continue;
        }
        for (int layer = warp; layer < kTY; layer += kWarps)
        {
            {
                // a layer without a single emitting voxel (air above the terrain) has no faces
                const Planes4 mv = planes_m(s_planes)[hrow(layer, lane)];
                if (!__any_sync(0xffffffffu, ~(mv.y & mv.z & mv.w) != 0u))
                {
                    if (lane == 0) { s_layer_cnt[layer] = 0; s_layer_typa[layer] = 0; s_layer_typb[layer] = 0; }
                    continue;
                }
            }
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
            uint32_t wide_tot[4] = { 0, 0, 0, 0 };
            if (kWide)
            {
                // a layer holds at most 4224 quads (16-bit fields are enough per layer), a chunk up to 196 608: 32-bit scans
#pragma unroll
                for (int t = 0; t < 4; t++)
                {
                    const uint32_t v0 = ((t < 2 ? a0 : b0) >> (16 * (t & 1))) & 0xFFFFu, v1 = ((t < 2 ? a1 : b1) >> (16 * (t & 1))) & 0xFFFFu;
                    const uint32_t vi = warp_incl_scan(v0 + v1, lane);
                    wide_tot[t] = __shfl_sync(0xffffffffu, vi, 31);
                    s_typ32[t * 64 + 2 * lane] = vi - (v0 + v1);
                    s_typ32[t * 64 + 2 * lane + 1] = vi - v1;
                }
            }
            else if (total <= (uint32_t)kMaxQuads)   // 16-bit fields cannot overflow below the cap
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
                if (!kWide && total > (uint32_t)kMaxQuads) o.flags = GC_CHUNK_OVERFLOW;
                else
                {
                    uint32_t tq[kMeshTypes];
                    tq[1] = ta_tot & 0xFFFFu; tq[2] = ta_tot >> 16; tq[3] = tb_tot & 0xFFFFu; tq[4] = tb_tot >> 16;
                    if (kWide) { tq[1] = wide_tot[0]; tq[2] = wide_tot[1]; tq[3] = wide_tot[2]; tq[4] = wide_tot[3]; }
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
                if (p.vert_table)   // chunk->totalVertices (WorldRender.cpp:559)
                {
                    const GcChunkCoord c = p.coords[chunk];
                    p.vert_table[(c.cz * p.size_x + c.cx) * 4 + c.cy] = (int32_t)(total * 4);
                }
            }
        }
        __syncthreads();

        GC_PROF_MARK(3);
        // ================= P3: emission, in batches of whole layers that fit the quad list =================
        if (s_ctx->emit)
        {
            const uint32_t quad_base = s_ctx->quad_base;
            uint32_t* idx_words = reinterpret_cast<uint32_t*>(p.index_arena) + (size_t)quad_base * 3;
            uint4* vtx = reinterpret_cast<uint4*>(p.vertex_arena) + (size_t)quad_base * 3;
            LightGather lg;
            lg.sun = p.sun; lg.light = p.light; lg.surf = s_surf;
            lg.own = ((long long)col0 * 4 + cy) * GC_CHUNK_SIZE_3;
            lg.size_x = p.size_x; lg.cx = cx; lg.cy = cy; lg.cz = cz;
            const uint16_t* chunk_blocks = p.blocks + lg.own;
            // the surface heights under the halo tile ([34][34] bytes out of the nine columns around the chunk); P3b reads
            // them after the barrier that follows P3a
            for (int i = tid; i < kSurfTileBytes; i += NT)
            {
                const int tz = i / kSurfStride - 1, x = i % kSurfStride - 1;
                const int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0), dcx = x < 0 ? -1 : (x > 31 ? 1 : 0);
                const size_t col = (size_t)(cz + dcz) * p.size_x + (cx + dcx);
                s_surf[i] = __ldg(p.surface + col * 1024 + (tz & 31) * 32 + (x & 31));
            }

            for (int l0 = 0; l0 < kTY;)
            {
                const uint32_t first = s_layer_base[l0];
                int l1 = kTY;                                               // nearly always: the rest of the chunk fits
                if (s_layer_base[kTY] - first > (uint32_t)kListCap)
                {
                    l1 = l0 + 1;
                    while (s_layer_base[l1 + 1] - first <= (uint32_t)kListCap) l1++;
                }
                const uint32_t count = s_layer_base[l1] - first;
                if (kWide && count)
                {
                    // the list's 16-bit per-type ranks are relative to the batch (a batch holds at most kListCap quads)
                    if (tid < kTY && tid >= l0 && tid < l1)
                    {
                        const uint32_t r1 = s_typ32[0 * 64 + tid] - s_typ32[0 * 64 + l0], r2 = s_typ32[1 * 64 + tid] - s_typ32[1 * 64 + l0];
                        const uint32_t r3 = s_typ32[2 * 64 + tid] - s_typ32[2 * 64 + l0], r4 = s_typ32[3 * 64 + tid] - s_typ32[3 * 64 + l0];
                        s_layer_typbase_a[tid] = r1 | (r2 << 16);
                        s_layer_typbase_b[tid] = r3 | (r4 << 16);
                    }
                    if (tid == 0)
                    {
                        const uint32_t b1 = s_typ32[0 * 64 + l0], b2 = s_typ32[1 * 64 + l0], b3 = s_typ32[2 * 64 + l0], b4 = s_typ32[3 * 64 + l0];
                        s_ctx->batch_base[0] = first - b1 - b2 - b3 - b4;
                        s_ctx->batch_base[1] = b1; s_ctx->batch_base[2] = b2; s_ctx->batch_base[3] = b3; s_ctx->batch_base[4] = b4;
                    }
                    __syncthreads();
                }
                if (count)
                {
                    // ---- 3a: rows -> ordered quad list + per-type ranks ----
                    for (int layer = l0 + warp; layer < l1; layer += kWarps)
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
                            trank[0] = (kWide ? rank0 - first : rank0) - trank[1] - trank[2] - trank[3] - trank[4];
                            p3a_row(rf, layer, lane, rank0 - first, trank, s_list, s_tlist);
                        }
                    }
                    __syncthreads();
                    GC_PROF_MARK(4);
                    // ---- 3b: one thread per quad: its index words and its four vertices ----
                    uint32_t id_next = 0;
                    if ((uint32_t)tid < count)
                    {
                        const uint32_t e = s_list[tid];
                        id_next = __ldg(chunk_blocks + ((e >> 3) & 31) + 32 * (((e >> 13) & 63) + 64 * ((e >> 8) & 31)));
                    }
                    for (uint32_t i = tid; i < count; i += NT)
                    {
                        const uint32_t e = s_list[i];
                        const uint32_t id = id_next & 255u;
                        if (i + NT < count)
                        {
                            const uint32_t en = s_list[i + NT];   // block id of this thread's next quad: its latency hides behind this one
                            id_next = __ldg(chunk_blocks + ((en >> 3) & 31) + 32 * (((en >> 13) & 63) + 64 * ((en >> 8) & 31)));
                        }
                        const uint32_t rank = first + i;
                        if (kWide)
                        {
                            // SetIndices' pattern with 32-bit indices: o + 2, o + 1, o, o + 3, o + 2, o
                            const uint32_t t = (e >> 19) & 7, o = rank * 4u;
                            uint2* ip = reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(p.index_arena) + (size_t)quad_base * 6) +
                                        (size_t)(s_ctx->type_off[t] + s_ctx->batch_base[t] + s_tlist[i]) * 3;
                            ip[0] = make_uint2(o + 2u, o + 1u); ip[1] = make_uint2(o, o + 3u); ip[2] = make_uint2(o + 2u, o);
                        }
                        else
                        {
                            uint32_t iw[3];
                            quad_index_words(rank, iw);
                            uint32_t* ip = idx_words + (size_t)(s_ctx->type_off[(e >> 19) & 7] + s_tlist[i]) * 3;
                            ip[0] = iw[0]; ip[1] = iw[1]; ip[2] = iw[2];
                        }
                        uint32_t v[12];
                        p3b_quad(e, s_faces[e & 7], lg, s_planes + 4 * kHaloRows, s_hx, s_mat[id], v);
                        uint4* dst = vtx + (size_t)rank * 3;
                        dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                        dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                        dst[2] = make_uint4(v[8], v[9], v[10], v[11]);
                    }
                    if (l1 < kTY) __syncthreads();   // the list is rewritten by the next batch
                    GC_PROF_MARK(5);
                }
                l0 = l1;
            }
        }
        // hx and the quad list were written into the staging ring by the generic proxy: order that before the next TMA writes
        fence_proxy_async();
    }
    // the last CTA out leaves the control block ready for the next launch (no memset between launches): every CTA's claims and
    // arena reservations precede its increment of the done counter, which wraps to zero by itself
    if (tid == 0)
    {
        __threadfence();
        if (atomicInc(p.done_counter, gridDim.x - 1) == gridDim.x - 1)
        {
            const unsigned long long total = p.keep_cursor ? atomicAdd(p.quad_cursor, 0ull) : atomicExch(p.quad_cursor, 0ull);
            *p.total_out = total;
            if (p.total_out_host)   // (mapped host memory: the host reads it once the launch's event has fired — no copy engine involved)
            {
                *p.total_out_host = total;
                __threadfence_system();
            }
            atomicExch(p.work_counter, 0);
        }
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
    // PCIe needs ~100 KB in flight (50 GB/s x 2 us); 16 KB per CTA, so a few dozen CTAs saturate the link (measured: 16 is too few, 64 .. 512 are alike) and the other
    // SMs stay free for a mesh kernel running on another stream (its CTA needs a whole SM's registers)
    static const int ctas_env = getenv("GCMESH_PULL_CTAS") ? atoi(getenv("GCMESH_PULL_CTAS")) : 0;
    const int cap = ctas_env > 0 ? ctas_env : 64;
    long long pieces = (long long)n_items * 8;
    int grid = pieces < cap ? (int)pieces : cap;
    gc_upload_kernel<<<grid, 256, 0, stream>>>(items, n_items, h_blocks, h_sun, h_light, d_blocks, d_sun, d_light);
    return cudaGetLastError();
}

size_t mesh_smem_bytes() { return (size_t)kSmemBytes; }

static cudaError_t configure_mesh_kernel()
{
    static bool configured = false;
    if (configured) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    const void* kernels[4] = { (const void*)gc_mesh_kernel<false, false, false>, (const void*)gc_mesh_kernel<true, false, false>,
                               (const void*)gc_mesh_kernel<false, true, false>, (const void*)gc_mesh_kernel<false, false, true> };
    for (int k = 0; k < 4 && e == cudaSuccess; k++)
    {
        e = cudaFuncSetAttribute(kernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        // the whole unified L1 / shared array as shared memory: two ~112 KB CTAs per SM
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernels[k], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (e == cudaSuccess) configured = true;
    return e;
}

int mesh_ctas_per_sm()
{
    if (configure_mesh_kernel() != cudaSuccess) return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gc_mesh_kernel<false, false, false>, NT, kSmemBytes) != cudaSuccess) return 0;
    int nh = 0;   // (the exchange instantiation must fit as many: its waiting CTAs rely on the grid being resident)
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nh, gc_mesh_kernel<false, true, false>, NT, kSmemBytes) != cudaSuccess) return 0;
    return n < nh ? n : nh;
}

cudaError_t launch_mesh(const MeshParams& p, const MeshTensorMaps& maps, int grid, cudaStream_t stream)
{
    cudaError_t e0 = configure_mesh_kernel();
    if (e0 != cudaSuccess) return e0;
    static const bool force_halo = getenv("GCMESH_FORCE_HALO_KERNEL") != nullptr;
    // launched with programmatic stream serialisation: back-to-back submissions overlap a launch's set-up with its predecessor's
    // tail (GCMESH_NO_PDL=1: plain stream order)
    static const bool no_pdl = getenv("GCMESH_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    if (p.wide) return cudaLaunchKernelEx(&cfg, gc_mesh_kernel<false, false, true>, p, maps);                        // (uncapped form: no exchange, no profiling)
    if (p.halo.enabled || force_halo) return cudaLaunchKernelEx(&cfg, gc_mesh_kernel<false, true, false>, p, maps);   // (no profiling instantiation of the exchange form)
    if (p.prof) return cudaLaunchKernelEx(&cfg, gc_mesh_kernel<true, false, false>, p, maps);
    return cudaLaunchKernelEx(&cfg, gc_mesh_kernel<false, false, false>, p, maps);
}

cudaError_t mesh_kernel_attributes(int* regs, int* static_smem, int* max_dyn_smem)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, gc_mesh_kernel<false, false, false>);
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

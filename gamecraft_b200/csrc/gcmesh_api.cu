// gcmesh — C ABI (include/gcmesh.h) over the sm_100a kernels.  C-style C++ host code: contexts, the device
// world (the window of columns), batch arenas, tensor-map encoding, submission and result hand-off.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <immintrin.h>
#include <string>
#include <thread>
#include <vector>

#include "gcmesh_kernel.cuh"
#include "gclight_kernel.cuh"
#include "gcrle_kernel.cuh"
#include "gcedit_kernel.cuh"
#include "gchalo_kernel.cuh"
#include "gcgen_kernel.cuh"
#include "gcfill_kernel.cuh"

using namespace gc;

namespace
{
thread_local std::string g_error;

int fail(int status, const char* what, const char* detail = nullptr)
{
    g_error = what;
    if (detail) { g_error += ": "; g_error += detail; }
    return status;
}

#define GC_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return fail(GC_ERR_CUDA, #call, cudaGetErrorString(e_)); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The reference's 24 block types (Code/Block.cpp:139-276; texture layers Code/Assets.h:7-34).
void solid(GcBlockInfo* b, int top, int bottom, int side)
{
    memset(b, 0, sizeof(*b));
    b->textures[0] = (uint8_t)top; b->textures[1] = (uint8_t)bottom;
    b->textures[2] = b->textures[3] = b->textures[4] = b->textures[5] = (uint8_t)side;
    b->alpha = 255; b->opaque = 1;                      // CreateDefaultBlock, Block.cpp:128-137
}

void reference_block_table(GcBlockInfo* t)
{
    solid(&t[0], 0, 0, 0); t[0].invisible = 1; t[0].cull = GC_CULL_INVISIBLE; t[0].opaque = 0;      // AIR
    solid(&t[1], 8, 6, 9);                                                                          // GRASS
    solid(&t[2], 6, 6, 6);                                                                          // DIRT
    solid(&t[3], 48, 48, 48);                                                                       // STONE
    solid(&t[4], 5, 5, 5);                                                                          // CRATE
    solid(&t[5], 44, 44, 44);                                                                       // SAND
    solid(&t[6], 49, 49, 49);                                                                       // STONE_BRICK
    solid(&t[7], 42, 42, 42);                                                                       // METAL_CRATE
    solid(&t[8], 51, 51, 51); t[8].mesh_type = GC_MESH_FLUID; t[8].cull = GC_CULL_FLUID; t[8].alpha = 127; t[8].opaque = 0;   // WATER
    solid(&t[9], 60, 60, 59);                                                                       // WOOD
    solid(&t[10], 21, 21, 21);                                                                      // LEAVES
    solid(&t[11], 3, 3, 3);                                                                         // CLAY
    solid(&t[12], 46, 6, 47);                                                                       // SNOW
    solid(&t[13], 10, 10, 10); t[13].mesh_type = GC_MESH_TRANSPARENT; t[13].cull = GC_CULL_TRANSPARENT; t[13].alpha = 190; t[13].opaque = 0;  // ICE
    solid(&t[14], 12, 12, 12); t[14].opaque = 0;                                                    // LANTERN
    solid(&t[15], 50, 50, 50);                                                                      // TRAMPOLINE
    solid(&t[16], 2, 0, 1);                                                                         // CACTUS
    solid(&t[17], 22, 22, 22); t[17].mesh_type = GC_MESH_MAGMA; t[17].opaque = 0;                   // MAGMA
    solid(&t[18], 4, 4, 4);                                                                         // COOLED_MAGMA
    solid(&t[19], 43, 43, 43);                                                                      // OBSIDIAN
    solid(&t[20], 0, 0, 0); t[20].invisible = 1; t[20].cull = GC_CULL_INVISIBLE; t[20].opaque = 0;  // KILL_ZONE
    solid(&t[21], 13, 13, 13); t[21].mesh_type = GC_MESH_FLUID; t[21].cull = GC_CULL_FLUID; t[21].alpha = 127; t[21].opaque = 0;  // LAVA
    solid(&t[22], 7, 7, 7); t[22].mesh_type = GC_MESH_CUTOUT; t[22].cull = GC_CULL_CUTOUT; t[22].opaque = 0;                      // GLASS
    solid(&t[23], 45, 45, 45);                                                                      // SHOCKER
}
}  // namespace

struct GcContext
{
    int device;
    cudaStream_t stream;
    bool own_stream;
    int num_sms;
    int mesh_ctas;            // persistent mesh CTAs that fit the device at once (SMs x resident CTAs per SM)
    EncodeTiledFn encode;
    DeviceTables* d_tables;
    int n_block_types;        // size of the block table in use (gcmesh_set_block_table)
    LightRules* d_light_rules;
    int use_tma;
    // gcmesh_mesh_host pipeline
    cudaStream_t upload_stream, upload_stream2, download_stream;   // (upload_stream2: the light arrays of a strip travel beside its block ids)
    cudaEvent_t ev_up2;
    cudaEvent_t ev_fork, ev_join, ev_uploaded[64], ev_meshed[64];
    unsigned long long* h_strip_cursor;   // pinned [64]
};

struct GcWorld
{
    GcContext* ctx;
    int size_x, size_z;
    int origin_cx, origin_cz;   // world-space column of window column (0, 0): gcmesh_world_set_origin / gcmesh_world_generate (region records)
    size_t n_cols, n_slots;
    uint16_t* d_blocks;
    uint8_t* d_sun;
    uint8_t* d_light;
    uint8_t* d_surface;
    MeshTensorMaps maps;
    // sparse upload state
    uint8_t* slot_dirty;      // host: per slot, bit a set = device array a of the slot may hold non-zero data
    uint8_t* scan_flags;      // host: per slot, bit a set = host array a is non-zero (result of the last scan)
    uint8_t* block_flags;     // host: per slot, 1 = the host block array is non-zero (first stage of the scan)
    uint32_t* h_items;        // pinned: work list of the upload kernel
    uint32_t* d_items;
    int launches;             // kernels launched by the last upload call
    unsigned long long upload_bytes;   // host -> device bytes the last sparse upload / gcmesh_mesh_host call moved
    bool surface_bounds_blocks;        // GC_WORLD_SURFACE_BOUNDS_BLOCKS
    bool surface_bounds_mesh;          // GC_WORLD_SURFACE_BOUNDS_MESH
    bool edit_replay;                  // GC_WORLD_EDIT_REPLAY
    uint64_t* d_edit_queues;           // two queues of kRelightQueueNodes positions (replay mode)
    int copy_run;                      // GC_WORLD_COPY_RUN (0: default)
    int16_t* col_smax;                 // host, per column: highest surface entry (-1: not computed in this scan)
    // light fill scratch (grown on demand)
    uint32_t* d_light_counters;   // 8 x u32
    uint32_t* h_light_counters;   // pinned
    uint8_t* d_brick_flags;       // per brick and sweep: changed (2 x (kLightSweeps + 1) arrays)
    uint8_t* d_group_sun; uint8_t* d_group_light; size_t group_light_slots;   // gcmesh_mesh_host without light arrays: the light of one group of rows
    uint32_t* d_light_bits;       // candidate bit maps: sunlight candidates, then opaque cells (2048 words per brick each)
    uint8_t* d_brick_top;         // compute_surface scratch: per brick, the highest block of every column cell
    int light_launches;
    // both light arenas are known to hold zeros (fresh window, gcmesh_world_generate): gcmesh_world_light need not clear them.
    // Every writer of the arenas resets it; handing out the raw pointers or connecting a halo peer disables it for good.
    bool light_zero, light_untracked;
    // the last writer of the light arenas was a sparse upload: arrays no mesh can depend on were skipped (and cleared), so the
    // arenas do not hold the whole light fixpoint that gcmesh_world_set_blocks builds on.  gcmesh_world_light and full uploads clear it.
    bool light_partial;
    // terrain generation scratch (gcmesh_world_generate)
    int16_t* d_gen_height; int32_t* d_gen_max_y;
    // peer-memory halo exchange (gcmesh_world_halo_*)
    uint32_t* d_halo_flags;
    HaloPeer halo_peer[2];        // [0]: the rank below, [1]: the rank above
    void* halo_ipc[2][5];         // pointers opened with cudaIpcOpenMemHandle (closed on destroy)
    uint32_t halo_serial;
    // block edits (gcmesh_world_set_blocks)
    int32_t* d_vert_table;        // per brick: vertices of its last build (written by the mesh kernel), -1 = never built
    uint8_t* d_edit_dirty;        // per brick: flagged for rebuild
    uint32_t* d_edit_state; uint8_t* d_edit_save; uint32_t* d_edit_sun_list; uint32_t* d_edit_light_list;
    GcBlockEdit* d_edits; int32_t* d_edit_status;   // kEditsPerPass entries each (fixed, so the graph's arguments stay valid)
    cudaGraphExec_t edit_graph; bool edit_graph_tried;
    GcChunkCoord* d_edit_coords; GcChunkCoord* h_edit_coords; uint32_t* h_edit_count;   // rebuild list (n_slots entries) + its length, host copies pinned
    // region-record (RLE) transfer scratch, grown on demand
    uint16_t* d_rle_data; size_t rle_data_cap;         // the uint16 stream
    uint32_t* d_rle_ends; size_t rle_ends_cap;         // one run end per (count, block) pair
    RleJob* d_rle_jobs; RleJob* h_rle_jobs; size_t rle_jobs_cap;
    uint32_t* d_rle_bad; uint32_t* h_rle_bad;          // records refused by the last decode
    uint16_t* d_rle_out; uint16_t* d_rle_dense; uint32_t* d_rle_words; uint32_t* h_rle_words; uint32_t* d_rle_slots; uint16_t* d_rle_offsets;   // encode scratch (kRleGroup bricks)   // last gcmesh_world_light: sunlight candidates, blockLight candidates, active bricks, seeds
};

struct GcBatch
{
    GcContext* ctx;
    int max_chunks;
    uint64_t max_quads;
    GcChunkCoord* d_coords;
    GcChunkOut* d_out;
    uint8_t* d_vertices;
    uint16_t* d_indices;
    unsigned long long* d_cursor;   // control block, 32 bytes: [0] quad cursor | work counter, error flag | CTAs done, pad | [3] quads of the last launch
    int32_t* d_counters;            // [0] work counter, [1] error flag (1: mbarrier watchdog, 2: a halo plane never arrived)
    unsigned long long* d_prof;     // 16 cycle counters when GCMESH_PROFILE=1
    unsigned long long h_prof[512];
    GcChunkOut* h_out;              // pinned
    GcChunkCoord* h_coords;         // pinned staging for coords
    unsigned long long* h_cursor;   // pinned
    int32_t* h_counters;            // pinned
    int n;
    int state;                      // 0 idle, 1 submitted, 2 waited
    int launches;
    cudaEvent_t ev_start, ev_stop, ev_done, ev_coords, ev_kernel;
    bool timing;                    // record ev_start / ev_stop around the kernel (gcmesh_batch_set_timing)
    bool timed;                     // ... and the last submission did
    int order_rows[2];              // gcmesh_submit_exchange: the boundary rows the order was built for (-1: plain submission)
    bool uncapped;                  // GC_BATCH_UNCAPPED: no per-chunk quad cap, 32-bit indices
    size_t index_quad_bytes;        // 12 (six uint16) or 24 (six uint32) bytes of indices per quad
    // work order: chunks are claimed heaviest first so that the kernel's tail is made of cheap ones
    int32_t* d_order;
    int32_t* h_order;               // pinned
    GcChunkCoord* order_coords;     // host copy of the coordinates the order was built for
    int order_n;
    bool order_from_history;        // built from the quad counts of a finished submission of the same list
    FillJob* d_fill_jobs; size_t fill_jobs_cap;   // gcmesh_batch_fill_device
};

static int encode_map(GcContext* ctx, CUtensorMap* map, CUtensorMapDataType type, int elem_bytes, void* base, size_t n_slots)
{
    // One brick arena viewed as a 4-D tensor.  The reference's brick index is x + 32*(y + 64*z) (World.cpp:278-281), so
    // a z-slab (32 x 64 voxels) is contiguous; described as (x*y8 = 256, y/8 = 8, z = 32, slot) with box (256, 8, 1, 1)
    // one TMA op moves a whole slab as 8 rows of 512 B / 256 B (64 rows of 64 B / 32 B made the TMA unit issue 8x
    // more, smaller requests).
    cuuint64_t dims[4] = { 256, 8, 32, (cuuint64_t)n_slots };
    cuuint64_t strides[3] = { (cuuint64_t)(256 * elem_bytes), (cuuint64_t)(32 * 64 * elem_bytes), (cuuint64_t)(GC_CHUNK_SIZE_3 * elem_bytes) };
    cuuint32_t box[4] = { 256, 8, 1, 1 }, estr[4] = { 1, 1, 1, 1 };
    CUresult r = ctx->encode(map, type, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        char buf[64];
        snprintf(buf, sizeof(buf), "CUresult %d", (int)r);
        return fail(GC_ERR_CUDA, "cuTensorMapEncodeTiled", buf);
    }
    return GC_OK;
}

// lightStep / lightEmitted of the reference's block table (Block.cpp:139-276); 255 = opaque (lightStep == INT_MAX)
static void reference_light_rules(uint8_t* step, uint8_t* emitted)
{
    for (int i = 0; i < 256; i++) { step[i] = 255; emitted[i] = 0; }
    step[0] = 1;                      // AIR        Block.cpp:146
    step[8] = 2;                      // WATER      :167
    step[13] = 2;                     // ICE        :211
    step[14] = 1; emitted[14] = 15;   // LANTERN    :217-218
    step[17] = 1; emitted[17] = 10;   // MAGMA      :235-236
    step[20] = 1;                     // KILL_ZONE  :252
    step[21] = 1; emitted[21] = 10;   // LAVA       :261-262
    step[22] = 1;                     // GLASS      :272
}

static int upload_light_rules(GcContext* ctx, const uint8_t* step, const uint8_t* emitted)
{
    LightRules h;
    memcpy(h.step, step, 256); memcpy(h.emitted, emitted, 256);
    cudaError_t e = cudaMemcpyAsync(ctx->d_light_rules, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(GC_ERR_CUDA, "upload light rules", cudaGetErrorString(e));
    return GC_OK;
}

static int upload_tables(GcContext* ctx, const GcBlockInfo* table, int n, int kill_zone_id)
{
    DeviceTables* h = new (std::nothrow) DeviceTables;
    if (!h) return fail(GC_ERR_ARG, "out of host memory");
    for (int i = 0; i < 256; i++)
    {
        // ids beyond the table read as AIR (the reference asserts block < BLOCK_COUNT, World.cpp:294)
        const GcBlockInfo& b = table[i < n ? i : 0];
        const int emits = (i != 0) && (i < n) && !b.invisible;   // BuildChunkAsync skips AIR and invisible, WorldRender.cpp:19-22
        const uint32_t c = class_byte(b.cull, b.opaque, b.mesh_type, emits);
        h->cls8[i] = (uint8_t)c;
        h->lut[i] = spread_class(c);
        h->mat[i].lo = (uint32_t)b.textures[0] | ((uint32_t)b.textures[1] << 8) | ((uint32_t)b.textures[2] << 16) | ((uint32_t)b.textures[3] << 24);
        h->mat[i].hi = (uint32_t)b.textures[4] | ((uint32_t)b.textures[5] << 8) | ((uint32_t)b.alpha << 16);
    }
    h->kill_zone = kill_zone_id;
    cudaError_t e = cudaMemcpyAsync(ctx->d_tables, h, sizeof(DeviceTables), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    delete h;
    if (e != cudaSuccess) return fail(GC_ERR_CUDA, "upload block table", cudaGetErrorString(e));
    ctx->n_block_types = n;
    return GC_OK;
}

template <class T>
static int grow_device(T** p, size_t* cap, size_t need)
{
    if (need <= *cap) return GC_OK;
    cudaFree(*p); *p = nullptr; *cap = 0;
    const size_t n = need + need / 4 + 1024;
    if (cudaMalloc(p, n * sizeof(T)) != cudaSuccess) return fail(GC_ERR_CUDA, "cudaMalloc", "region record scratch");
    *cap = n;
    return GC_OK;
}

extern "C" {

const char* gcmesh_version(void) { return "gcmesh 0.1 (sm_100a)"; }
const char* gcmesh_last_error(void) { return g_error.c_str(); }

int gcmesh_default_block_table(GcBlockInfo* table, int* n, int* kill_zone_id)
{
    if (!table) return fail(GC_ERR_ARG, "table is null");
    reference_block_table(table);
    if (n) *n = 24;                    // BLOCK_COUNT, Block.h:36
    if (kill_zone_id) *kill_zone_id = 20;  // BLOCK_KILL_ZONE, Block.h:32
    return GC_OK;
}

int gcmesh_create(int device, void* stream, GcContext** out)
{
    if (!out) return fail(GC_ERR_ARG, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(GC_ERR_CUDA, "no usable CUDA device (gcmesh has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(GC_ERR_ARG, "device index out of range");
    GC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(GC_ERR_CUDA, "gcmesh kernels are built for sm_100a only", prop.name);

    GcContext* ctx = new (std::nothrow) GcContext();
    if (!ctx) return fail(GC_ERR_ARG, "out of host memory");
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    {
        const int per_sm = mesh_ctas_per_sm();
        if (per_sm < 1) { delete ctx; return fail(GC_ERR_CUDA, "the mesh kernel does not fit this device", cudaGetErrorString(cudaGetLastError())); }
        static const int cap_env = getenv("GCMESH_CTAS_PER_SM") ? atoi(getenv("GCMESH_CTAS_PER_SM")) : 0;
        ctx->mesh_ctas = ctx->num_sms * (cap_env > 0 && cap_env < per_sm ? cap_env : per_sm);
    }
    ctx->own_stream = stream == nullptr;
    ctx->stream = (cudaStream_t)stream;
    const char* env = getenv("GCMESH_NO_TMA");
    ctx->use_tma = (env && env[0] == '1') ? 0 : 1;
    if (ctx->own_stream) { e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking); if (e != cudaSuccess) { delete ctx; return fail(GC_ERR_CUDA, "cudaStreamCreate", cudaGetErrorString(e)); } }

    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { delete ctx; return fail(GC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found"); }
    ctx->encode = (EncodeTiledFn)fn;

    e = cudaMalloc(&ctx->d_tables, sizeof(DeviceTables));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->upload_stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_up2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->download_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    for (int i = 0; i < 64 && e == cudaSuccess; i++)
    {
        e = cudaEventCreate(&ctx->ev_uploaded[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_meshed[i]);
    }
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_strip_cursor, sizeof(unsigned long long) * 64);
    if (e != cudaSuccess) { delete ctx; return fail(GC_ERR_CUDA, "context allocation", cudaGetErrorString(e)); }
    GcBlockInfo table[24];
    reference_block_table(table);
    int st = upload_tables(ctx, table, 24, 20);
    if (st == GC_OK)
    {
        if (cudaMalloc(&ctx->d_light_rules, sizeof(LightRules)) != cudaSuccess) st = fail(GC_ERR_CUDA, "cudaMalloc", "light rules");
        else { uint8_t step[256], emitted[256]; reference_light_rules(step, emitted); st = upload_light_rules(ctx, step, emitted); }
    }
    if (st != GC_OK) { cudaFree(ctx->d_tables); cudaFree(ctx->d_light_rules); delete ctx; return st; }
    *out = ctx;
    return GC_OK;
}

int gcmesh_destroy(GcContext* ctx)
{
    if (!ctx) return GC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_tables); cudaFree(ctx->d_light_rules);
    cudaStreamDestroy(ctx->upload_stream); cudaStreamDestroy(ctx->upload_stream2); cudaStreamDestroy(ctx->download_stream);
    cudaEventDestroy(ctx->ev_up2);
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join);
    for (int i = 0; i < 64; i++) { cudaEventDestroy(ctx->ev_uploaded[i]); cudaEventDestroy(ctx->ev_meshed[i]); }
    cudaFreeHost(ctx->h_strip_cursor);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GC_OK;
}

int gcmesh_set_stream(GcContext* ctx, void* stream)
{
    if (!ctx) return fail(GC_ERR_ARG, "ctx is null");
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)stream;
    return GC_OK;
}

int gcmesh_set_block_table(GcContext* ctx, const GcBlockInfo* table, int n, int kill_zone_id)
{
    if (!ctx || !table) return fail(GC_ERR_ARG, "null argument");
    if (n < 1 || n > GC_MAX_BLOCK_TYPES) return fail(GC_ERR_ARG, "table size must be 1..256");
    if (kill_zone_id < 0 || kill_zone_id >= n) return fail(GC_ERR_ARG, "kill_zone_id outside the table");
    for (int i = 0; i < n; i++)
    {
        if (table[i].cull > GC_CULL_INVISIBLE) return fail(GC_ERR_ARG, "cull must be 0..4 (CullValue, Block.h:40-47)");
        if (table[i].mesh_type >= GC_MESH_TYPE_COUNT) return fail(GC_ERR_ARG, "mesh_type must be 0..4 (BlockMeshType, Mesh.h:8-16)");
    }
    GC_CUDA(cudaSetDevice(ctx->device));
    return upload_tables(ctx, table, n, kill_zone_id);
}

int gcmesh_synchronize(GcContext* ctx)
{
    if (!ctx) return fail(GC_ERR_ARG, "ctx is null");
    GC_CUDA(cudaSetDevice(ctx->device));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->upload_stream));
    GC_CUDA(cudaStreamSynchronize(ctx->upload_stream2));
    GC_CUDA(cudaStreamSynchronize(ctx->download_stream));
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
int gcmesh_world_create(GcContext* ctx, int size_x, int size_z, GcWorld** out)
{
    if (!ctx || !out) return fail(GC_ERR_ARG, "null argument");
    *out = nullptr;
    if (size_x < 3 || size_z < 3) return fail(GC_ERR_ARG, "a window needs at least 3x3 columns (edge columns are halo only)");
    GC_CUDA(cudaSetDevice(ctx->device));
    GcWorld* w = new (std::nothrow) GcWorld();
    if (!w) return fail(GC_ERR_ARG, "out of host memory");
    w->ctx = ctx; w->size_x = size_x; w->size_z = size_z;
    w->n_cols = (size_t)size_x * size_z;
    w->n_slots = w->n_cols * GC_WORLD_CHUNK_HEIGHT;
    cudaError_t e = cudaMalloc(&w->d_blocks, w->n_slots * GC_CHUNK_SIZE_3 * 2);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_sun, w->n_slots * GC_CHUNK_SIZE_3);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_light, w->n_slots * GC_CHUNK_SIZE_3);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_surface, w->n_cols * GC_CHUNK_SIZE_2);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_blocks, 0, w->n_slots * GC_CHUNK_SIZE_3 * 2, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_sun, 0, w->n_slots * GC_CHUNK_SIZE_3, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_light, 0, w->n_slots * GC_CHUNK_SIZE_3, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_surface, 0, w->n_cols * GC_CHUNK_SIZE_2, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocHost(&w->h_items, sizeof(uint32_t) * w->n_slots * 3);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_items, sizeof(uint32_t) * w->n_slots * 3);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_vert_table, sizeof(int32_t) * w->n_slots);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->d_vert_table, 0xFF, sizeof(int32_t) * w->n_slots, ctx->stream);   // -1: never built
    if (e != cudaSuccess) { gcmesh_world_destroy(w); return fail(GC_ERR_CUDA, "world allocation", cudaGetErrorString(e)); }
    w->slot_dirty = (uint8_t*)calloc(w->n_slots, 1);
    w->scan_flags = (uint8_t*)calloc(w->n_slots, 1);
    w->block_flags = (uint8_t*)calloc(w->n_slots, 1);
    w->col_smax = (int16_t*)malloc(sizeof(int16_t) * w->n_cols);
    if (!w->slot_dirty || !w->scan_flags || !w->block_flags || !w->col_smax) { gcmesh_world_destroy(w); return fail(GC_ERR_ARG, "out of host memory"); }
    w->light_zero = true;
    int st = GC_OK;
    if (st == GC_OK) st = encode_map(ctx, &w->maps.blocks_slab, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, w->d_blocks, w->n_slots);
    if (st != GC_OK) { gcmesh_world_destroy(w); return st; }
    *out = w;
    return GC_OK;
}

int gcmesh_world_destroy(GcWorld* w)
{
    if (!w) return GC_OK;
    cudaSetDevice(w->ctx->device);
    cudaStreamSynchronize(w->ctx->stream);
    cudaFree(w->d_blocks); cudaFree(w->d_sun); cudaFree(w->d_light); cudaFree(w->d_surface); cudaFree(w->d_items);
    if (w->h_items) cudaFreeHost(w->h_items);
    cudaFree(w->d_rle_data); cudaFree(w->d_rle_ends); cudaFree(w->d_rle_jobs); cudaFree(w->d_rle_bad);
    cudaFree(w->d_rle_out); cudaFree(w->d_rle_dense); cudaFree(w->d_rle_words); cudaFree(w->d_rle_slots); cudaFree(w->d_rle_offsets);
    if (w->h_rle_jobs) cudaFreeHost(w->h_rle_jobs);
    if (w->h_rle_bad) cudaFreeHost(w->h_rle_bad);
    if (w->h_rle_words) cudaFreeHost(w->h_rle_words);
    cudaFree(w->d_brick_top);
    cudaFree(w->d_light_counters); cudaFree(w->d_brick_flags); cudaFree(w->d_light_bits); cudaFree(w->d_group_sun); cudaFree(w->d_group_light);
    if (w->h_light_counters) cudaFreeHost(w->h_light_counters);
    for (int d = 0; d < 2; d++)
        for (int k = 0; k < 5; k++)
            if (w->halo_ipc[d][k]) cudaIpcCloseMemHandle(w->halo_ipc[d][k]);
    cudaFree(w->d_halo_flags);
    cudaFree(w->d_gen_height); cudaFree(w->d_gen_max_y);
    cudaFree(w->d_vert_table); cudaFree(w->d_edit_dirty); cudaFree(w->d_edit_state); cudaFree(w->d_edit_save); cudaFree(w->d_edit_queues);
    cudaFree(w->d_edit_sun_list); cudaFree(w->d_edit_light_list); cudaFree(w->d_edits); cudaFree(w->d_edit_status); cudaFree(w->d_edit_coords);
    if (w->edit_graph) cudaGraphExecDestroy(w->edit_graph);
    if (w->h_edit_coords) cudaFreeHost(w->h_edit_coords);
    if (w->h_edit_count) cudaFreeHost(w->h_edit_count);
    free(w->slot_dirty); free(w->scan_flags); free(w->block_flags); free(w->col_smax);
    delete w;
    return GC_OK;
}

static bool column_ok(const GcWorld* w, int cx, int cz) { return cx >= 0 && cx < w->size_x && cz >= 0 && cz < w->size_z; }

int gcmesh_world_upload_chunk(GcWorld* w, int cx, int cy, int cz, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (!column_ok(w, cx, cz) || cy < 0 || cy >= GC_WORLD_CHUNK_HEIGHT) return fail(GC_ERR_ARG, "chunk position outside the window");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    const size_t slot = ((size_t)cz * w->size_x + cx) * GC_WORLD_CHUNK_HEIGHT + cy;
    w->slot_dirty[slot] |= (blocks ? 1 : 0) | (sunlight ? 2 : 0) | (block_light ? 4 : 0);
    if (sunlight || block_light) w->light_zero = false;
    if (blocks) GC_CUDA(cudaMemcpyAsync(w->d_blocks + slot * GC_CHUNK_SIZE_3, blocks, GC_CHUNK_SIZE_3 * 2, cudaMemcpyHostToDevice, w->ctx->stream));
    if (sunlight) GC_CUDA(cudaMemcpyAsync(w->d_sun + slot * GC_CHUNK_SIZE_3, sunlight, GC_CHUNK_SIZE_3, cudaMemcpyHostToDevice, w->ctx->stream));
    if (block_light) GC_CUDA(cudaMemcpyAsync(w->d_light + slot * GC_CHUNK_SIZE_3, block_light, GC_CHUNK_SIZE_3, cudaMemcpyHostToDevice, w->ctx->stream));
    return GC_OK;
}

int gcmesh_world_upload_surface(GcWorld* w, int cx, int cz, const uint8_t* surface)
{
    if (!w || !surface) return fail(GC_ERR_ARG, "null argument");
    if (!column_ok(w, cx, cz)) return fail(GC_ERR_ARG, "column outside the window");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    GC_CUDA(cudaMemcpyAsync(w->d_surface + ((size_t)cz * w->size_x + cx) * GC_CHUNK_SIZE_2, surface, GC_CHUNK_SIZE_2, cudaMemcpyHostToDevice, w->ctx->stream));
    return GC_OK;
}

int gcmesh_world_upload_rows(GcWorld* w, int cz0, int cz1, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light, const uint8_t* surface)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (cz0 < 0 || cz1 > w->size_z || cz0 >= cz1) return fail(GC_ERR_ARG, "row range outside the window");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    const size_t s0 = (size_t)cz0 * w->size_x * GC_WORLD_CHUNK_HEIGHT * GC_CHUNK_SIZE_3;
    const size_t ns = (size_t)(cz1 - cz0) * w->size_x * GC_WORLD_CHUNK_HEIGHT * GC_CHUNK_SIZE_3;
    memset(w->slot_dirty + (size_t)cz0 * w->size_x * GC_WORLD_CHUNK_HEIGHT, 7, (size_t)(cz1 - cz0) * w->size_x * GC_WORLD_CHUNK_HEIGHT);
    if (blocks) GC_CUDA(cudaMemcpyAsync(w->d_blocks + s0, blocks + s0, ns * 2, cudaMemcpyHostToDevice, w->ctx->stream));
    if (sunlight || block_light) w->light_zero = false;
    if (sunlight && block_light && cz0 == 0 && cz1 == w->size_z) w->light_partial = false;   // every light array of the window is replaced
    if (sunlight) GC_CUDA(cudaMemcpyAsync(w->d_sun + s0, sunlight + s0, ns, cudaMemcpyHostToDevice, w->ctx->stream));
    if (block_light) GC_CUDA(cudaMemcpyAsync(w->d_light + s0, block_light + s0, ns, cudaMemcpyHostToDevice, w->ctx->stream));
    if (surface)
    {
        const size_t c0 = (size_t)cz0 * w->size_x * GC_CHUNK_SIZE_2, nc = (size_t)(cz1 - cz0) * w->size_x * GC_CHUNK_SIZE_2;
        GC_CUDA(cudaMemcpyAsync(w->d_surface + c0, surface + c0, nc, cudaMemcpyHostToDevice, w->ctx->stream));
    }
    return GC_OK;
}

int gcmesh_world_upload_all(GcWorld* w, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light, const uint8_t* surface)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    return gcmesh_world_upload_rows(w, 0, w->size_z, blocks, sunlight, block_light, surface);
}

// Is a buffer (a multiple of 4 KB) all zero?  Exits at the first non-zero 4 KB block.  Memory-bound: ~20 GB/s per thread
// with AVX2, ~12 GB/s with the scalar form (measured on the B200 box's Xeon); a software prefetch ahead of the loads
// changed nothing.
static bool all_zero_scalar(const void* p, size_t bytes)
{
    const uint64_t* q = (const uint64_t*)p;
    const size_t n = bytes / 8;
    for (size_t i = 0; i < n; i += 512)
    {
        uint64_t acc = 0;
        for (size_t j = 0; j < 512; j += 8) acc |= q[i + j] | q[i + j + 1] | q[i + j + 2] | q[i + j + 3] | q[i + j + 4] | q[i + j + 5] | q[i + j + 6] | q[i + j + 7];
        if (acc) return false;
    }
    return true;
}

__attribute__((target("avx2"))) static bool all_zero_avx2(const void* p, size_t bytes)
{
    const char* q = (const char*)p;
    for (size_t i = 0; i < bytes; i += 4096)
    {
        __m256i acc = _mm256_setzero_si256();
        for (size_t j = 0; j < 4096; j += 256)
        {
            const char* r = q + i + j;
            const __m256i a = _mm256_or_si256(_mm256_loadu_si256((const __m256i*)r), _mm256_loadu_si256((const __m256i*)(r + 32)));
            const __m256i b = _mm256_or_si256(_mm256_loadu_si256((const __m256i*)(r + 64)), _mm256_loadu_si256((const __m256i*)(r + 96)));
            const __m256i c = _mm256_or_si256(_mm256_loadu_si256((const __m256i*)(r + 128)), _mm256_loadu_si256((const __m256i*)(r + 160)));
            const __m256i d = _mm256_or_si256(_mm256_loadu_si256((const __m256i*)(r + 192)), _mm256_loadu_si256((const __m256i*)(r + 224)));
            acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
        }
        if (!_mm256_testz_si256(acc, acc)) return false;
    }
    return true;
}

static inline bool all_zero_small(const void* p, size_t bytes)   // a multiple of 8
{
    const uint64_t* q = (const uint64_t*)p;
    uint64_t acc = 0;
    for (size_t i = 0; i < bytes / 8; i++) acc |= q[i];
    return acc == 0;
}

static bool all_zero(const void* p, size_t bytes)
{
    static bool (*const impl)(const void*, size_t) = __builtin_cpu_supports("avx2") ? all_zero_avx2 : all_zero_scalar;
    return impl(p, bytes);
}

// Zero / non-zero classification of the host brick arrays of one slot into w->scan_flags.
// A sunlight brick that lies wholly at or above its column's surface is never read: every cell there reads 15 whatever
// is stored (GetSunlight, Lighting.cpp:69-79; light_pair_of in mesh_core.cuh).  With the surface map at hand such bricks
// are classified "nothing to transfer" without being scanned — in a terrain world that is most of the sky.
// highest entry of a column's surface map, computed once per scan by whoever asks first (col_smax: -1 = not yet)
static inline int column_smax(GcWorld* w, const uint8_t* surface, size_t col)
{
    int v = w->col_smax[col];
    if (v >= 0) return v;
    const uint8_t* s = surface + col * GC_CHUNK_SIZE_2;
    v = 0;
    for (int i = 0; i < GC_CHUNK_SIZE_2; i++) v = s[i] > v ? s[i] : v;
    w->col_smax[col] = (int16_t)v;
    return v;
}

static inline bool sun_brick_never_read(GcWorld* w, const uint8_t* surface, size_t slot)
{
    if (!surface) return false;
    return (int)(slot & 3) * GC_CHUNK_SIZE_V >= column_smax(w, surface, slot >> 2);
}

// With GC_WORLD_SURFACE_BOUNDS_BLOCKS (every block of a column lies below its surface entry, layer y = 255 aside): a quad
// reads light within one voxel of its block, so at or below the highest surface entry of the block's column; an all-AIR
// brick that starts above the highest entry of the 3 x 3 columns around it is out of every quad's reach — neither of its
// light arrays is scanned or transferred.  (Bricks of the top row are left to the neighbourhood rule: a block on layer 255,
// which the surface map does not see, lights cells of that row.)
static inline bool light_out_of_reach(GcWorld* w, const uint8_t* surface, size_t slot)
{
    if (!w->surface_bounds_blocks || !surface || w->block_flags[slot]) return false;
    const int cy = (int)(slot & 3);
    if (cy >= GC_WORLD_CHUNK_HEIGHT - 1) return false;
    const int col = (int)(slot >> 2), cx = col % w->size_x, cz = col / w->size_x;
    int top = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dx = -1; dx <= 1; dx++)
        {
            const int nx = cx + dx, nz = cz + dz;
            if (nx < 0 || nx >= w->size_x || nz < 0 || nz >= w->size_z) continue;
            const int v = column_smax(w, surface, (size_t)nz * w->size_x + nx);
            top = v > top ? v : top;
        }
    return cy * GC_CHUNK_SIZE_V > top;
}

// The scan of a window runs in two stages per brick.  Stage A classifies the block ids.  Stage B classifies the two light
// arrays, and skips them — nothing to scan, nothing to transfer — for a brick whose whole 3 x 3 x 3 brick neighbourhood
// holds no block at all: light is only ever read for a quad, within one voxel of the block the quad belongs to
// (VertexLight, WorldRender.cpp:329-383), so no chunk that gets meshed can reach such a brick's light.  In a terrain
// world that is most of the sky.  (Like the sunlight rule above this leaves the device copy of such an array as it was.)
// With GC_WORLD_SURFACE_BOUNDS_BLOCKS the caller vouches that `surface` is ComputeSurface of the blocks (Lighting.cpp:5-26:
// highest non-AIR y in 0..254, plus one — what ChunkGroup::surface holds for every pre-processed column; ComputeSurfaceAt
// keeps it so on edits, a stale entry is only ever too high).  A brick lying wholly at or above its column's highest entry
// is then AIR by definition and is not read — except layer y = 255, which ComputeSurface never looks at: the top brick's
// last layer (32 rows of 64 bytes) is still checked.  In a terrain world that is three quarters of the block ids.
static inline void scan_blocks(GcWorld* w, size_t slot, const uint16_t* blocks, const uint8_t* surface)
{
    const uint16_t* b = blocks + slot * GC_CHUNK_SIZE_3;
    if (w->surface_bounds_blocks && sun_brick_never_read(w, surface, slot))
    {
        uint8_t any = 0;
        if ((slot & 3) == GC_WORLD_CHUNK_HEIGHT - 1)
            for (int z = 0; z < GC_CHUNK_SIZE_H && !any; z++)
                any = all_zero_small(b + 32 * (63 + 64 * z), 64) ? 0 : 1;
        w->block_flags[slot] = any;
        return;
    }
    w->block_flags[slot] = all_zero(b, (size_t)GC_CHUNK_SIZE_3 * 2) ? 0 : 1;
}

static inline bool brick_neighbourhood_is_air(const GcWorld* w, size_t slot)
{
    const int cy = (int)(slot & 3), col = (int)(slot >> 2), cx = col % w->size_x, cz = col / w->size_x;
    for (int dz = -1; dz <= 1; dz++)
        for (int dx = -1; dx <= 1; dx++)
        {
            const int nx = cx + dx, nz = cz + dz;
            if (nx < 0 || nx >= w->size_x || nz < 0 || nz >= w->size_z) continue;
            const uint8_t* f = w->block_flags + ((size_t)nz * w->size_x + nx) * GC_WORLD_CHUNK_HEIGHT;
            for (int ny = cy - 1; ny <= cy + 1; ny++)
                if (ny >= 0 && ny < GC_WORLD_CHUNK_HEIGHT && f[ny]) return false;
        }
    return true;
}

static inline void scan_lights(GcWorld* w, size_t slot, const uint8_t* sunlight, const uint8_t* block_light, const uint8_t* surface)
{
    uint8_t f = w->block_flags[slot];
    if (!sunlight) { w->scan_flags[slot] = f; return; }          // (no light arrays handed in: the light is filled on the device)
    static const bool skip_rule = []() { const char* e = getenv("GCMESH_SCAN_LIGHT_SKIP"); return !(e && e[0] == '0'); }();
    if (!skip_rule || !(brick_neighbourhood_is_air(w, slot) || light_out_of_reach(w, surface, slot)))
    {
        if (!sun_brick_never_read(w, surface, slot) && !all_zero(sunlight + slot * GC_CHUNK_SIZE_3, GC_CHUNK_SIZE_3)) f |= 2;
        if (!all_zero(block_light + slot * GC_CHUNK_SIZE_3, GC_CHUNK_SIZE_3)) f |= 4;
    }
    w->scan_flags[slot] = f;
}

// Work order of the two stages over the rows of a window: A0, A1, B0, A2, B1, ... so that stage B of a row starts as soon
// as stage A of the rows around it can be done, and rows still finish in ascending order (the strip pipeline of
// gcmesh_mesh_host consumes them that way).  Unit u of 2 * rows -> (stage, row).
struct ScanPlan
{
    GcWorld* w;
    const uint16_t* blocks; const uint8_t* sunlight; const uint8_t* block_light; const uint8_t* surface;
    size_t slots_per_row;
    int rows;
    std::vector<std::atomic<int>> a_remaining;   // per row: slots whose stage A is still to do
    std::atomic<size_t> next;

    ScanPlan(GcWorld* w_, const uint16_t* b, const uint8_t* s, const uint8_t* l, const uint8_t* sf)
        : w(w_), blocks(b), sunlight(s), block_light(l), surface(sf), slots_per_row((size_t)w_->size_x * GC_WORLD_CHUNK_HEIGHT), rows(w_->size_z),
          a_remaining((size_t)w_->size_z), next(0)
    {
        memset(w->col_smax, 0xFF, sizeof(int16_t) * w->n_cols);
        for (auto& r : a_remaining) r.store((int)slots_per_row);
    }

    // runs on every scan thread; done(slot) is called when both stages of a slot are finished
    void work(const std::function<void(size_t)>& done)
    {
        const size_t total = 2 * slots_per_row * (size_t)rows;
        for (;;)
        {
            const size_t t0 = next.fetch_add(4);
            if (t0 >= total) break;
            const size_t t1 = t0 + 4 < total ? t0 + 4 : total;
            for (size_t t = t0; t < t1; t++)
            {
                const int u = (int)(t / slots_per_row);
                const size_t in_row = t % slots_per_row;
                bool stage_b; int row;
                if (u == 0) { stage_b = false; row = 0; }
                else if (u == 2 * rows - 1) { stage_b = true; row = rows - 1; }
                else if (u & 1) { stage_b = false; row = (u + 1) / 2; }
                else { stage_b = true; row = u / 2 - 1; }
                const size_t slot = (size_t)row * slots_per_row + in_row;
                if (!stage_b)
                {
                    scan_blocks(w, slot, blocks, surface);
                    a_remaining[(size_t)row].fetch_sub(1, std::memory_order_release);
                }
                else
                {
                    // stage A of the rows around was claimed before this task; wait for whoever still runs it
                    for (int r = row - 1; r <= row + 1; r++)
                        if (r >= 0 && r < rows)
                            while (a_remaining[(size_t)r].load(std::memory_order_acquire) > 0) std::this_thread::yield();
                    scan_lights(w, slot, sunlight, block_light, surface);
                    done(slot);
                }
            }
        }
    }
};

static int scan_threads(int threads)
{
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    return nt > 64 ? 64 : nt;
}

// Enqueue on `s` the upload of rows [cz0, cz1) according to w->scan_flags: non-zero arrays are pulled by the GPU (or copied
// one by one when the host arrays are not pinned), arrays that turned zero are cleared.  Returns kernels launched (< 0: error).
static int enqueue_rows(GcWorld* w, int cz0, int cz1, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light,
                        bool mapped, cudaStream_t s, cudaStream_t s_light = nullptr)
{
    // s_light: a second stream for the strided copies of the two light arrays — a strided copy pays per row, two of them
    // side by side (two copy engines) keep the link busy.  The caller orders s_light before and after this call.
    if (!s_light) s_light = s;
    const size_t s_begin = (size_t)cz0 * w->size_x * GC_WORLD_CHUNK_HEIGHT, s_end = (size_t)cz1 * w->size_x * GC_WORLD_CHUNK_HEIGHT;
    const size_t c_begin = s_begin / GC_WORLD_CHUNK_HEIGHT, c_end = s_end / GC_WORLD_CHUNK_HEIGHT;
    uint32_t* items = w->h_items + s_begin * 3;   // this row range's own part of the pinned work list
    w->light_zero = false;
    w->light_partial = true;
    int n_items = 0, n_copies = 0;
    // Non-zero arrays.  A terrain world's non-zero bricks lie at the same height in column after column, and a column's four
    // bricks are consecutive slots: the same array of brick cy of k consecutive columns is a strided block — k rows of one
    // brick array, four brick arrays apart — that ONE 2-D copy moves with the copy engine at full link speed in large
    // requests.  (Measured: the GPU pull kernel reads pinned host memory at 51 GB/s alone but 31 GB/s while the meshes of the
    // previous strip travel the other way — its many small read requests share the upstream direction with that data — where
    // copy-engine traffic keeps 47 GB/s each way; one copy per array gets 20 GB/s.)  Short runs go to the pull kernel.
    static const int env_run = []() { const char* e = getenv("GCMESH_COPY_RUN"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 4; }();
    const int min_run = w->copy_run > 0 ? w->copy_run : env_run;
    for (int a = 0; a < 3; a++)
    {
        const size_t bytes = a == 0 ? (size_t)GC_CHUNK_SIZE_3 * 2 : (size_t)GC_CHUNK_SIZE_3;
        const uint8_t* src0 = a == 0 ? (const uint8_t*)blocks : (a == 1 ? sunlight : block_light);
        if (!src0) continue;
        uint8_t* dst0 = a == 0 ? (uint8_t*)w->d_blocks : (a == 1 ? w->d_sun : w->d_light);
        for (int cy = 0; cy < GC_WORLD_CHUNK_HEIGHT; cy++)
        {
            size_t col = c_begin;
            while (col < c_end)
            {
                if (!(w->scan_flags[col * GC_WORLD_CHUNK_HEIGHT + cy] & (1 << a))) { col++; continue; }
                size_t run = 1;
                while (col + run < c_end && (w->scan_flags[(col + run) * GC_WORLD_CHUNK_HEIGHT + cy] & (1 << a))) run++;
                w->upload_bytes += bytes * run;
                const size_t first = col * GC_WORLD_CHUNK_HEIGHT + cy;
                if ((int)run >= min_run || !mapped)
                {
                    // (no light arrays handed in: the block-id runs take turns on the two streams — two copy engines side by side)
                    cudaStream_t sc = a == 0 ? ((!sunlight && (n_copies++ & 1)) ? s_light : s) : s_light;
                    const cudaError_t e = run == 1 ? cudaMemcpyAsync(dst0 + first * bytes, src0 + first * bytes, bytes, cudaMemcpyHostToDevice, sc)
                                                   : cudaMemcpy2DAsync(dst0 + first * bytes, bytes * GC_WORLD_CHUNK_HEIGHT, src0 + first * bytes, bytes * GC_WORLD_CHUNK_HEIGHT,
                                                                       bytes, run, cudaMemcpyHostToDevice, sc);
                    if (e != cudaSuccess) return -1;
                }
                else
                    for (size_t k = 0; k < run; k++) items[n_items++] = (uint32_t)((first + k * GC_WORLD_CHUNK_HEIGHT) * 4 + a);
                col += run;
            }
        }
    }
    // arrays that held data on the device and are zero now
    for (size_t slot = s_begin; slot < s_end; slot++)
    {
        const uint8_t keep = sunlight ? 0 : 6;                    // (no light arrays handed in: the device's light arrays are left alone)
        const uint8_t f = w->scan_flags[slot] | (w->slot_dirty[slot] & keep), stale = w->slot_dirty[slot] & ~f;
        for (int a = 0; a < 3 && stale; a++)
            if (stale & (1 << a))
            {
                const size_t bytes = a == 0 ? (size_t)GC_CHUNK_SIZE_3 * 2 : (size_t)GC_CHUNK_SIZE_3;
                uint8_t* dst = a == 0 ? (uint8_t*)w->d_blocks + slot * bytes : (a == 1 ? w->d_sun : w->d_light) + slot * bytes;
                if (cudaMemsetAsync(dst, 0, bytes, s) != cudaSuccess) return -1;
            }
        w->slot_dirty[slot] = f;
    }
    if (n_items)
    {
        w->upload_bytes += sizeof(uint32_t) * (size_t)n_items;
        if (cudaMemcpyAsync(w->d_items + s_begin * 3, items, sizeof(uint32_t) * (size_t)n_items, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
        if (launch_upload_items(w->d_items + s_begin * 3, n_items, blocks, sunlight, block_light, w->d_blocks, w->d_sun, w->d_light, s) != cudaSuccess) return -1;
        return 1;
    }
    return 0;
}

// Classify rows [cz0, cz1) (scan with `threads` host threads, or the caller's hint) and enqueue their upload.
static int upload_rows_sparse(GcWorld* w, int cz0, int cz1, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light,
                              const uint8_t* surface, const uint8_t* nonzero_hint, int threads, bool mapped, cudaStream_t s)
{
    const size_t s_begin = (size_t)cz0 * w->size_x * GC_WORLD_CHUNK_HEIGHT, s_end = (size_t)cz1 * w->size_x * GC_WORLD_CHUNK_HEIGHT;
    if (nonzero_hint) memcpy(w->scan_flags + s_begin, nonzero_hint + s_begin, s_end - s_begin);
    else
    {
        // (the two-stage scan looks at neighbouring rows: the window is always classified as a whole)
        const int nt = scan_threads(threads);
        ScanPlan plan(w, blocks, sunlight, block_light, surface);
        auto work = [&]() { plan.work([](size_t) {}); };
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    }
    return enqueue_rows(w, cz0, cz1, blocks, sunlight, block_light, mapped, s);
}

static bool host_arrays_mapped(const void* blocks, const void* sunlight, const void* block_light)
{
    cudaPointerAttributes attr;
    const bool mapped = cudaPointerGetAttributes(&attr, blocks) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr
                        && (!sunlight || (cudaPointerGetAttributes(&attr, sunlight) == cudaSuccess && attr.type == cudaMemoryTypeHost))
                        && (!block_light || (cudaPointerGetAttributes(&attr, block_light) == cudaSuccess && attr.type == cudaMemoryTypeHost));
    cudaGetLastError();
    return mapped;
}

int gcmesh_world_upload_sparse(GcWorld* w, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light, const uint8_t* surface,
                               const uint8_t* nonzero_hint, int threads)
{
    if (!w || !blocks || !sunlight || !block_light) return fail(GC_ERR_ARG, "null argument");
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    w->upload_bytes = surface ? w->n_cols * GC_CHUNK_SIZE_2 : 0;
    const int k = upload_rows_sparse(w, 0, w->size_z, blocks, sunlight, block_light, surface, nonzero_hint, threads,
                                     host_arrays_mapped(blocks, sunlight, block_light), ctx->stream);
    if (k < 0) return fail(GC_ERR_CUDA, "sparse upload", cudaGetErrorString(cudaGetLastError()));
    w->launches = k;
    if (surface) GC_CUDA(cudaMemcpyAsync(w->d_surface, surface, w->n_cols * GC_CHUNK_SIZE_2, cudaMemcpyHostToDevice, ctx->stream));
    return GC_OK;
}

int gcmesh_world_set_option(GcWorld* w, int option, int value)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (option == GC_WORLD_SURFACE_BOUNDS_BLOCKS) { w->surface_bounds_blocks = value != 0; return GC_OK; }
    if (option == GC_WORLD_SURFACE_BOUNDS_MESH) { w->surface_bounds_mesh = value != 0; return GC_OK; }
    if (option == GC_WORLD_EDIT_REPLAY) { w->edit_replay = value != 0; return GC_OK; }
    if (option == GC_WORLD_COPY_RUN) { if (value < 0) return fail(GC_ERR_ARG, "run length must not be negative"); w->copy_run = value; return GC_OK; }
    return fail(GC_ERR_ARG, "unknown world option");
}

int gcmesh_world_set_origin(GcWorld* w, int origin_cx, int origin_cz)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    w->origin_cx = origin_cx; w->origin_cz = origin_cz;
    return GC_OK;
}

int gcmesh_world_validate_blocks(GcWorld* w, unsigned long long* n_invalid, GcChunkCoord* first_chunk, int* first_voxel)
{
    if (!w || !n_invalid) return fail(GC_ERR_ARG, "null argument");
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    unsigned long long* d_out = nullptr;
    GC_CUDA(cudaMalloc(&d_out, 2 * sizeof(unsigned long long)));
    unsigned long long h[2] = { 0ull, ~0ull };
    cudaError_t e = cudaMemcpyAsync(d_out, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_validate_blocks(w->d_blocks, w->n_slots * GC_CHUNK_SIZE_3, ctx->n_block_types, d_out, ctx->num_sms, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(GC_ERR_CUDA, "validate blocks", cudaGetErrorString(e));
    *n_invalid = h[0];
    if (h[0])
    {
        const size_t slot = (size_t)(h[1] / GC_CHUNK_SIZE_3);
        if (first_chunk)
        {
            first_chunk->cy = (int)(slot % GC_WORLD_CHUNK_HEIGHT);
            first_chunk->cx = (int)((slot / GC_WORLD_CHUNK_HEIGHT) % w->size_x);
            first_chunk->cz = (int)((slot / GC_WORLD_CHUNK_HEIGHT) / w->size_x);
        }
        if (first_voxel) *first_voxel = (int)(h[1] % GC_CHUNK_SIZE_3);
    }
    return GC_OK;
}

int gcmesh_world_upload_launches(const GcWorld* w) { return w ? w->launches : 0; }
unsigned long long gcmesh_world_upload_bytes(const GcWorld* w) { return w ? w->upload_bytes : 0; }

int gcmesh_world_device_ptrs(GcWorld* w, void** blocks, void** sunlight, void** block_light, void** surface)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (blocks) *blocks = w->d_blocks;
    if (sunlight) *sunlight = w->d_sun;
    if (block_light) *block_light = w->d_light;
    if (sunlight || block_light) w->light_untracked = true;   // a producer outside this library may write them from now on
    if (surface) *surface = w->d_surface;
    return GC_OK;
}

size_t gcmesh_world_halo_bytes(const GcWorld* w) { return w ? halo_bytes(w->size_x) : 0; }

int gcmesh_world_pack_halo(GcWorld* w, int cz, int z, void* device_buf)
{
    if (!w || !device_buf) return fail(GC_ERR_ARG, "null argument");
    if (cz < 0 || cz >= w->size_z || z < 0 || z >= GC_CHUNK_SIZE_H) return fail(GC_ERR_ARG, "row / plane outside the window");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    GC_CUDA(launch_halo_copy(w->d_blocks, w->d_sun, w->d_light, w->d_surface, w->size_x, cz, z, (uint8_t*)device_buf, false, w->ctx->stream));
    return GC_OK;
}

int gcmesh_world_unpack_halo(GcWorld* w, int cz, int z, const void* device_buf)
{
    if (!w || !device_buf) return fail(GC_ERR_ARG, "null argument");
    if (cz < 0 || cz >= w->size_z || z < 0 || z >= GC_CHUNK_SIZE_H) return fail(GC_ERR_ARG, "row / plane outside the window");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    w->light_zero = false;
    GC_CUDA(launch_halo_copy(w->d_blocks, w->d_sun, w->d_light, w->d_surface, w->size_x, cz, z, (uint8_t*)device_buf, true, w->ctx->stream));
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Column pre-processing on the device: the producers of the mesher's light inputs (ComputeSurface, SetLightNodes).
int gcmesh_default_light_rules(uint8_t* step, uint8_t* emitted)
{
    if (!step || !emitted) return fail(GC_ERR_ARG, "null argument");
    reference_light_rules(step, emitted);
    return GC_OK;
}

int gcmesh_set_light_rules(GcContext* ctx, const uint8_t* step, const uint8_t* emitted)
{
    if (!ctx || !step || !emitted) return fail(GC_ERR_ARG, "null argument");
    for (int i = 0; i < 256; i++)
        if (emitted[i] > 15 || (step[i] != 255 && (step[i] < 1 || step[i] > 15))) return fail(GC_ERR_ARG, "lightStep must be 1..15 or 255 (opaque), lightEmitted 0..15");
    GC_CUDA(cudaSetDevice(ctx->device));
    return upload_light_rules(ctx, step, emitted);
}

static int surface_scratch(GcWorld* w)
{
    if (w->d_brick_top) return GC_OK;
    GC_CUDA(cudaSetDevice(w->ctx->device));
    GC_CUDA(cudaMalloc(&w->d_brick_top, w->n_slots * GC_CHUNK_SIZE_2));
    return GC_OK;
}

static LightWorld light_world(GcWorld* w)
{
    LightWorld lw;
    lw.blocks = w->d_blocks; lw.sun = w->d_sun; lw.light = w->d_light; lw.surface = w->d_surface;
    lw.size_x = w->size_x; lw.size_z = w->size_z; lw.rules = w->ctx->d_light_rules;
    return lw;
}

int gcmesh_world_compute_surface(GcWorld* w)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    if (surface_scratch(w) != GC_OK) return GC_ERR_CUDA;
    GC_CUDA(launch_surface(light_world(w), w->d_brick_top, w->ctx->stream));
    return GC_OK;
}

// Scratch of the light fill (sized for the whole window once; a fill over a range of rows uses the front of it).
static int light_scratch(GcWorld* w)
{
    if (w->d_light_counters) return GC_OK;
    const size_t bits_bytes = w->n_slots * 2048 * sizeof(uint32_t);
    const size_t flag_stride = (w->n_slots + 255) & ~(size_t)255;
    const size_t unit_stride = (w->n_slots * 8 + 255) & ~(size_t)255;
    GC_CUDA(cudaMalloc(&w->d_light_counters, kLightCounterWords * sizeof(uint32_t)));
    GC_CUDA(cudaMallocHost(&w->h_light_counters, kLightCounterWords * sizeof(uint32_t)));
    // two bit maps + the two visit lists (2 fills x 8 units per brick each) + three "row changed" maps (2 fills x 64 words per brick each)
    GC_CUDA(cudaMalloc(&w->d_light_bits, 2 * bits_bytes + (32 + 3 * 128) * w->n_slots * sizeof(uint32_t)));
    GC_CUDA(cudaMalloc(&w->d_brick_flags, 2 * (size_t)(kLightSweeps + 1) * flag_stride + unit_stride));
    return GC_OK;
}

// Enqueue the fill of `lw` — the window, or a range of its rows seen as a window of its own (rows are contiguous ranges of
// slots) — on `s`.  lw.sun / lw.light must hold zeros.  Nothing below depends on a count the device produces: candidates
// are bit maps of fixed size (a word per x-row of a brick), the active set is a byte per brick and sweep — a fixed sequence.
static int enqueue_light_fill(GcWorld* w, const LightWorld& lw, cudaStream_t s, int ctas_per_sm = 0)
{
    if (light_scratch(w) != GC_OK) return GC_ERR_CUDA;
    const size_t n_slots = (size_t)lw.size_x * lw.size_z * GC_WORLD_CHUNK_HEIGHT;
    const size_t bits_bytes = n_slots * 2048 * sizeof(uint32_t);
    const size_t flag_stride = (n_slots + 255) & ~(size_t)255;
    const size_t unit_stride = (n_slots * 8 + 255) & ~(size_t)255;               // per unit (256 rows; 8 per brick): holds a sunlight candidate
    const size_t flags_bytes = 2 * (size_t)(kLightSweeps + 1) * flag_stride + unit_stride;   // sunlight, then blockLight: sweep k reads [k], writes [k + 1]; then the unit bytes
    uint32_t* c = w->d_light_counters;
    uint32_t* sun_bits = w->d_light_bits;
    uint32_t* opq_bits = w->d_light_bits + n_slots * 2048;
    uint8_t* sun_flags = w->d_brick_flags;
    uint8_t* light_flags = w->d_brick_flags + (size_t)(kLightSweeps + 1) * flag_stride;
    uint8_t* sun_unit = w->d_brick_flags + 2 * (size_t)(kLightSweeps + 1) * flag_stride;
    const int sms = w->ctx->num_sms;
    GC_CUDA(launch_zero(c, kLightCounterWords * sizeof(uint32_t), sms, s));
    GC_CUDA(launch_zero(w->d_light_bits, 2 * bits_bytes, sms, s));
    GC_CUDA(launch_zero(w->d_brick_flags, flags_bytes, sms, s));
    // 1. one pass over the column cells up to maxY: candidate bit maps, emitters seeded, the first sweep's active bricks
    GC_CUDA(launch_light_scan(lw, sun_bits, opq_bits, sun_flags, sun_unit, light_flags, c, s));
    uint32_t* visit_lists = w->d_light_bits + 2 * n_slots * 2048;
    uint32_t* row_maps = visit_lists + 32 * n_slots;
    GC_CUDA(launch_zero(row_maps + 128 * n_slots, 128 * n_slots * sizeof(uint32_t), sms, s));   // the map the first sweep writes (the kernel clears the others)
    // 2. relaxation: a value falls by at least one per hop and nothing at or below MIN_LIGHT is stored, so 14 in-place sweeps
    // reach the fixpoint; one persistent kernel runs them with grid barriers in between, visits per sweep only the bricks
    // that changed in the sweep before (or whose face neighbours did) — in units of 256 rows whose candidates a CTA spreads over
    // its threads one cell each — and ends with the first sweep that finds none
    GC_CUDA(launch_relax_all(lw, sun_bits, opq_bits, sun_unit, sun_flags, light_flags, flag_stride, visit_lists, row_maps, c, w->ctx->num_sms, ctas_per_sm, s));
    return GC_OK;
}

int gcmesh_world_light(GcWorld* w)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (w->n_slots > 65535) return fail(GC_ERR_CAPACITY, "light fill addresses voxels with 32 bits: at most 65535 bricks per window");
    if (w->size_x < 3 || w->size_z < 3) return fail(GC_ERR_ARG, "a window needs an edge ring of columns around the pre-processed ones");
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // freshly generated chunks hold no light (World.cpp:632): the fill starts from zero (arenas known to be zero are not cleared again)
    if (!(w->light_zero && !w->light_untracked))
    {
        GC_CUDA(cudaMemsetAsync(w->d_sun, 0, w->n_slots * GC_CHUNK_SIZE_3, s));
        GC_CUDA(cudaMemsetAsync(w->d_light, 0, w->n_slots * GC_CHUNK_SIZE_3, s));
    }
    w->light_zero = false;
    w->light_partial = false;   // from here on the arenas hold the whole fill
    if (enqueue_light_fill(w, light_world(w), s) != GC_OK) return GC_ERR_CUDA;
    w->light_launches = 2;
    // the sparse-upload bookkeeping no longer knows what the device arrays hold
    memset(w->slot_dirty, 7, w->n_slots);
    return GC_OK;
}

int gcmesh_world_light_stats(GcWorld* w, uint32_t* stats4, int* launches)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (stats4)
    {
        memset(stats4, 0, 4 * sizeof(uint32_t));
        if (w->d_light_counters)
        {
            GC_CUDA(cudaSetDevice(w->ctx->device));
            GC_CUDA(cudaMemcpyAsync(w->h_light_counters, w->d_light_counters, kLightCounterWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->ctx->stream));
            GC_CUDA(cudaStreamSynchronize(w->ctx->stream));
            uint32_t visits = 0;
            for (int k = 0; k < kLightSweeps; k++) visits += w->h_light_counters[LC_VISITS + k];
            stats4[0] = w->h_light_counters[LC_SUN_CAND]; stats4[1] = visits;
            stats4[2] = w->h_light_counters[LC_SWEEPS]; stats4[3] = w->h_light_counters[LC_SEEDS];
        }
    }
    if (launches) *launches = w->light_launches;
    return GC_OK;
}

int gcmesh_world_download(GcWorld* w, uint16_t* blocks, uint8_t* sunlight, uint8_t* block_light, uint8_t* surface)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    cudaStream_t s = w->ctx->stream;
    if (blocks) GC_CUDA(cudaMemcpyAsync(blocks, w->d_blocks, w->n_slots * GC_CHUNK_SIZE_3 * 2, cudaMemcpyDeviceToHost, s));
    if (sunlight) GC_CUDA(cudaMemcpyAsync(sunlight, w->d_sun, w->n_slots * GC_CHUNK_SIZE_3, cudaMemcpyDeviceToHost, s));
    if (block_light) GC_CUDA(cudaMemcpyAsync(block_light, w->d_light, w->n_slots * GC_CHUNK_SIZE_3, cudaMemcpyDeviceToHost, s));
    if (surface) GC_CUDA(cudaMemcpyAsync(surface, w->d_surface, w->n_cols * GC_CHUNK_SIZE_2, cudaMemcpyDeviceToHost, s));
    GC_CUDA(cudaStreamSynchronize(s));
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Terrain of freshly created columns, generated in place (Generation.cpp:83-369, :772-836); kernels in gcgen_kernel.cu.
int gcmesh_world_generate(GcWorld* w, int kind, uint32_t seed, int origin_cx, int origin_cz, int radius, int falloff)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (kind < GC_GEN_FLAT || kind > GC_GEN_DUNGEON) return fail(GC_ERR_ARG, "unknown generator");
    if (kind != GC_GEN_VOID && kind != GC_GEN_DUNGEON && (radius <= 0 || falloff < 0 || falloff >= radius)) return fail(GC_ERR_ARG, "need 0 <= falloff < radius");
    // Square() in IsWithinIsland is integer arithmetic (Generation.cpp:68): keep world coordinates where it cannot overflow
    {
        const long long lim = 32000;
        const long long x0 = (long long)origin_cx * 32, x1 = ((long long)origin_cx + w->size_x) * 32;
        const long long z0 = (long long)origin_cz * 32, z1 = ((long long)origin_cz + w->size_z) * 32;
        if (x0 < -lim || x1 > lim || z0 < -lim || z1 > lim) return fail(GC_ERR_ARG, "window further than 32 000 blocks from the world origin");
    }
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    w->origin_cx = origin_cx; w->origin_cz = origin_cz;
    if (!w->d_gen_height)
    {
        GC_CUDA(cudaMalloc(&w->d_gen_height, sizeof(int16_t) * w->n_cols * GC_CHUNK_SIZE_2));
        GC_CUDA(cudaMalloc(&w->d_gen_max_y, sizeof(int32_t) * w->n_cols));
    }
    // CreateChunkGroup hands out zeroed chunks (World.cpp:624-643): all AIR, no light, surface 0
    GC_CUDA(cudaMemsetAsync(w->d_blocks, 0, w->n_slots * GC_CHUNK_SIZE_3 * 2, s));
    GC_CUDA(cudaMemsetAsync(w->d_sun, 0, w->n_slots * GC_CHUNK_SIZE_3, s));
    GC_CUDA(cudaMemsetAsync(w->d_light, 0, w->n_slots * GC_CHUNK_SIZE_3, s));
    GC_CUDA(cudaMemsetAsync(w->d_surface, 0, w->n_cols * GC_CHUNK_SIZE_2, s));
    GC_CUDA(cudaMemsetAsync(w->d_vert_table, 0xFF, sizeof(int32_t) * w->n_slots, s));   // new chunks have never been built
    GenArgs a;
    a.blocks = w->d_blocks; a.surface = w->d_surface; a.height = w->d_gen_height; a.max_y = w->d_gen_max_y;
    a.size_x = w->size_x; a.size_z = w->size_z; a.origin_cx = origin_cx; a.origin_cz = origin_cz;
    a.seed = seed; a.radius = radius; a.falloff = falloff;
    a.biome = kind == GC_GEN_ISLANDS ? 1 : (kind == GC_GEN_SNOW ? 2 : (kind == GC_GEN_DESERT ? 3 : (kind == GC_GEN_VOLCANIC ? 4 : 0)));
    if (kind == GC_GEN_FLAT) GC_CUDA(launch_generate_flat(a, s));
    else if (kind == GC_GEN_DUNGEON) GC_CUDA(launch_generate_dungeon(a, s));
    else if (kind == GC_GEN_GRID || kind == GC_GEN_VOID)
    {
        if (surface_scratch(w) != GC_OK) return GC_ERR_CUDA;
        if (kind == GC_GEN_GRID) GC_CUDA(launch_generate_grid(a, s)); else GC_CUDA(launch_generate_void(a, s));
        GC_CUDA(launch_surface(light_world(w), w->d_brick_top, s));
    }
    else GC_CUDA(launch_generate_terrain(a, s));   // (the terrain and flat kernels know their columns' tops: they write the surface map themselves)
    // the sparse-upload bookkeeping no longer knows what the device arrays hold
    memset(w->slot_dirty, 7, w->n_slots);
    w->light_zero = true;
    w->light_partial = false;
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Slab halo exchange over peer memory (multi-GPU, one process per GPU): kernels in gchalo_kernel.cu.
static int halo_flags(GcWorld* w)
{
    if (w->d_halo_flags) return GC_OK;
    GC_CUDA(cudaSetDevice(w->ctx->device));
    GC_CUDA(preload_halo_kernels());
    GC_CUDA(cudaMalloc(&w->d_halo_flags, HF_WORDS * sizeof(uint32_t)));
    // cleared on the context stream and waited for: a neighbour may write the block as soon as it knows the pointer
    GC_CUDA(cudaMemsetAsync(w->d_halo_flags, 0, HF_WORDS * sizeof(uint32_t), w->ctx->stream));
    GC_CUDA(cudaStreamSynchronize(w->ctx->stream));
    return GC_OK;
}

int gcmesh_world_halo_export(GcWorld* w, void* handles)
{
    if (!w || !handles) return fail(GC_ERR_ARG, "null argument");
    if (halo_flags(w) != GC_OK) return GC_ERR_CUDA;
    w->light_untracked = true;   // neighbours write this window's halo rows from now on
    void* ptrs[5] = { w->d_blocks, w->d_sun, w->d_light, w->d_surface, w->d_halo_flags };
    cudaIpcMemHandle_t* out = (cudaIpcMemHandle_t*)handles;
    static_assert(sizeof(cudaIpcMemHandle_t) * 5 == GC_HALO_HANDLE_BYTES, "GC_HALO_HANDLE_BYTES");
    for (int k = 0; k < 5; k++) GC_CUDA(cudaIpcGetMemHandle(out + k, ptrs[k]));
    return GC_OK;
}

static int halo_side_ok(GcWorld* w, int side, int peer_row)
{
    if (side < 0 || side > 1) return fail(GC_ERR_ARG, "side is 0 (the rank below, z - 1) or 1 (the rank above)");
    if (peer_row < 0 || peer_row >= w->size_z) return fail(GC_ERR_ARG, "peer row outside the window (neighbours have windows of one size)");
    if (w->halo_peer[side].connected) return fail(GC_ERR_ARG, "side already connected");
    return GC_OK;
}

int gcmesh_world_halo_connect_ipc(GcWorld* w, int side, const void* handles, int peer_row)
{
    if (!w || !handles) return fail(GC_ERR_ARG, "null argument");
    if (halo_side_ok(w, side, peer_row) != GC_OK) return GC_ERR_ARG;
    if (halo_flags(w) != GC_OK) return GC_ERR_CUDA;
    GC_CUDA(cudaSetDevice(w->ctx->device));
    const cudaIpcMemHandle_t* in = (const cudaIpcMemHandle_t*)handles;
    for (int k = 0; k < 5; k++)
    {
        cudaError_t e = cudaIpcOpenMemHandle(&w->halo_ipc[side][k], in[k], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
        {
            for (int j = 0; j < k; j++) { cudaIpcCloseMemHandle(w->halo_ipc[side][j]); w->halo_ipc[side][j] = nullptr; }
            w->halo_ipc[side][k] = nullptr;
            return fail(GC_ERR_CUDA, "cudaIpcOpenMemHandle", cudaGetErrorString(e));
        }
    }
    HaloPeer& p = w->halo_peer[side];
    p.blocks = (uint16_t*)w->halo_ipc[side][0]; p.sun = (uint8_t*)w->halo_ipc[side][1]; p.light = (uint8_t*)w->halo_ipc[side][2];
    p.surface = (uint8_t*)w->halo_ipc[side][3]; p.flags = (uint32_t*)w->halo_ipc[side][4];
    p.row = peer_row; p.connected = 1;
    return GC_OK;
}

int gcmesh_world_halo_connect_local(GcWorld* w, int side, GcWorld* peer, int peer_row)
{
    if (!w || !peer) return fail(GC_ERR_ARG, "null argument");
    if (peer->size_x != w->size_x || peer->size_z != w->size_z) return fail(GC_ERR_ARG, "neighbour windows must have one size");
    // each side's exchange waits for the other's: on one stream the second could never start (it would only time out)
    if (peer == w || peer->ctx == w->ctx || peer->ctx->stream == w->ctx->stream)
        return fail(GC_ERR_ARG, "neighbour worlds need contexts (streams) of their own");
    if (halo_side_ok(w, side, peer_row) != GC_OK) return GC_ERR_ARG;
    if (halo_flags(w) != GC_OK || halo_flags(peer) != GC_OK) return GC_ERR_CUDA;
    if (peer->ctx->device != w->ctx->device)
    {
        GC_CUDA(cudaSetDevice(w->ctx->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->ctx->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fail(GC_ERR_CUDA, "cudaDeviceEnablePeerAccess", cudaGetErrorString(e));
    }
    peer->light_untracked = true;   // this world writes the neighbour's halo rows from now on
    HaloPeer& p = w->halo_peer[side];
    p.blocks = peer->d_blocks; p.sun = peer->d_sun; p.light = peer->d_light; p.surface = peer->d_surface; p.flags = peer->d_halo_flags;
    p.row = peer_row; p.connected = 1;
    return GC_OK;
}

static unsigned long long halo_wait_ns()
{
    static const unsigned long long wait_ms = []() -> unsigned long long
    {
        const char* env = getenv("GCMESH_HALO_WAIT_MS");
        const long v = env ? atol(env) : 0;
        return v > 0 ? (unsigned long long)v : 10000ull;
    }();
    return wait_ms * 1000000ull;
}

int gcmesh_world_halo_push(GcWorld* w, int send_row_down, int send_row_up)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (!w->halo_peer[0].connected && !w->halo_peer[1].connected) return fail(GC_ERR_ARG, "no neighbour connected");
    if ((w->halo_peer[0].connected && (send_row_down < 0 || send_row_down >= w->size_z)) ||
        (w->halo_peer[1].connected && (send_row_up < 0 || send_row_up >= w->size_z))) return fail(GC_ERR_ARG, "send row outside the window");
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    HaloPushArgs a;
    a.blocks = w->d_blocks; a.sun = w->d_sun; a.light = w->d_light; a.surface = w->d_surface; a.flags = w->d_halo_flags;
    a.size_x = w->size_x; a.send_row[0] = send_row_down; a.send_row[1] = send_row_up;
    a.peer[0] = w->halo_peer[0]; a.peer[1] = w->halo_peer[1];
    a.serial = ++w->halo_serial;
    a.wait_ns = halo_wait_ns();
    GC_CUDA(launch_halo_push(a, ctx->num_sms, ctx->stream));
    // the neighbours' planes land in this window's halo rows: the sparse-upload bookkeeping no longer knows them
    for (int d = 0; d < 2; d++)
        if (w->halo_peer[d].connected)
        {
            const int row = d == 0 ? send_row_down - 1 : send_row_up + 1;
            if (row >= 0 && row < w->size_z) memset(w->slot_dirty + (size_t)row * w->size_x * GC_WORLD_CHUNK_HEIGHT, 7, (size_t)w->size_x * GC_WORLD_CHUNK_HEIGHT);
        }
    return GC_OK;
}

int gcmesh_world_halo_status(GcWorld* w, int* timed_out, uint32_t* serial)
{
    if (!w) return fail(GC_ERR_ARG, "world is null");
    if (timed_out) *timed_out = 0;
    if (serial) *serial = w->halo_serial;
    if (!w->d_halo_flags) return GC_OK;
    GC_CUDA(cudaSetDevice(w->ctx->device));
    uint32_t flags[HF_WORDS];
    GC_CUDA(cudaMemcpyAsync(flags, w->d_halo_flags, sizeof(flags), cudaMemcpyDeviceToHost, w->ctx->stream));
    GC_CUDA(cudaStreamSynchronize(w->ctx->stream));
    if (timed_out) *timed_out = (int)flags[HF_ERROR];
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Block edits on the device window: SetBlock -> ChunkOverflowed -> RecomputeLight -> FlagChunkForUpdate
// (World.cpp:507-582, WorldRender.cpp:412-439, Lighting.cpp:283-449); kernels in gcedit_kernel.cu.
int gcmesh_world_set_blocks(GcWorld* w, const GcBlockEdit* edits, int n, int32_t* status, GcChunkCoord* rebuild, int rebuild_capacity, int* n_rebuild)
{
    if (n_rebuild) *n_rebuild = 0;
    if (!w || (!edits && n > 0)) return fail(GC_ERR_ARG, "null argument");
    if (n < 0 || rebuild_capacity < 0 || (!rebuild && rebuild_capacity > 0)) return fail(GC_ERR_ARG, "bad edit or rebuild list size");
    if (w->n_slots > 65535) return fail(GC_ERR_CAPACITY, "light fill addresses voxels with 32 bits: at most 65535 bricks per window");
    for (int i = 0; i < n; i++)
    {
        // SetBlock asserts !IsEdgeChunk (World.cpp:560); a position outside the window has no chunk at all
        const int cx = edits[i].wx >> 5, cz = edits[i].wz >> 5;
        if (edits[i].wx < 0 || edits[i].wz < 0 || cx < 1 || cx >= w->size_x - 1 || cz < 1 || cz >= w->size_z - 1)
            return fail(GC_ERR_ARG, "edit outside the pre-processed (non-edge) columns of the window");
    }
    if (n == 0) return GC_OK;
    // the relight reads the cells around its box as the fixpoint they are supposed to hold: a sparse upload left light arrays
    // out (those no mesh depends on), so the window's light is not that fixpoint
    if (w->light_partial)
        return fail(GC_ERR_STATE, "the window's light arrays come from a sparse upload (gcmesh_world_upload_sparse / gcmesh_mesh_host): run gcmesh_world_light or a full upload before editing");
    w->light_zero = false;
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    constexpr int kEditsPerPass = 4096;
    if (!w->d_edit_state)
    {
        GC_CUDA(cudaMalloc(&w->d_edit_dirty, w->n_slots));
        GC_CUDA(cudaMemsetAsync(w->d_edit_dirty, 0, w->n_slots, s));
        GC_CUDA(cudaMalloc(&w->d_edit_save, 2 * (size_t)kEditBoxCells));
        GC_CUDA(cudaMalloc(&w->d_edit_sun_list, sizeof(uint32_t) * kEditBoxCells));
        GC_CUDA(cudaMalloc(&w->d_edit_light_list, sizeof(uint32_t) * kEditBoxCells));
        GC_CUDA(cudaMalloc(&w->d_edit_coords, sizeof(GcChunkCoord) * w->n_slots));
        GC_CUDA(cudaMallocHost(&w->h_edit_coords, sizeof(GcChunkCoord) * w->n_slots));
        GC_CUDA(cudaMallocHost(&w->h_edit_count, sizeof(uint32_t)));
        GC_CUDA(cudaMalloc(&w->d_edits, sizeof(GcBlockEdit) * kEditsPerPass));
        GC_CUDA(cudaMalloc(&w->d_edit_status, sizeof(int32_t) * kEditsPerPass));
        GC_CUDA(cudaMalloc(&w->d_edit_state, sizeof(uint32_t) * kEditWords));
    }
    EditArgs a;
    a.w = light_world(w); a.blocks_rw = w->d_blocks; a.tables = ctx->d_tables; a.vert_table = w->d_vert_table;
    a.edits = w->d_edits; a.status = w->d_edit_status; a.dirty = w->d_edit_dirty; a.state = w->d_edit_state;
    a.save = w->d_edit_save; a.sun_list = w->d_edit_sun_list; a.light_list = w->d_edit_light_list;
    a.replay = w->edit_replay ? 1 : 0; a.queue_a = nullptr; a.queue_b = nullptr;
    if (w->edit_replay)
    {
        if (!w->d_edit_queues) GC_CUDA(cudaMalloc(&w->d_edit_queues, 2 * sizeof(uint64_t) * (size_t)kRelightQueueNodes));
        a.queue_a = w->d_edit_queues; a.queue_b = w->d_edit_queues + kRelightQueueNodes;
    }
    // One edit is 19 tiny launches whose arguments never change (the kernels find the edit through a device-side cursor):
    // captured once per world into a CUDA graph and replayed per edit.  GCMESH_EDIT_GRAPH=0, a stream that cannot be
    // captured (the legacy default stream) or a failed capture fall back to plain launches of the same sequence.
    if (!w->edit_graph_tried && !w->edit_replay)
    {
        w->edit_graph_tried = true;
        const char* env = getenv("GCMESH_EDIT_GRAPH");
        if (!(env && env[0] == '0') && s != nullptr && s != cudaStreamLegacy)
        {
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
            {
                const cudaError_t le = launch_edit(a, s);
                const cudaError_t ce = cudaStreamEndCapture(s, &graph);
                if (le == cudaSuccess && ce == cudaSuccess && graph && cudaGraphInstantiate(&w->edit_graph, graph, 0) != cudaSuccess) w->edit_graph = nullptr;
                if (graph) cudaGraphDestroy(graph);
            }
            cudaGetLastError();   // a failed capture leaves a sticky-free error behind; the fallback path starts clean
        }
    }
    for (int first = 0; first < n; first += kEditsPerPass)
    {
        const int count = n - first < kEditsPerPass ? n - first : kEditsPerPass;
        GC_CUDA(cudaMemcpyAsync(w->d_edits, edits + first, sizeof(GcBlockEdit) * count, cudaMemcpyHostToDevice, s));
        GC_CUDA(cudaMemsetAsync(w->d_edit_state + EP_CURSOR, 0, sizeof(uint32_t), s));
        for (int i = 0; i < count; i++)
        {
            if (w->edit_replay) GC_CUDA(launch_edit_replay(a, s));
            else if (w->edit_graph) GC_CUDA(cudaGraphLaunch(w->edit_graph, s));
            else GC_CUDA(launch_edit(a, s));
        }
        if (status) GC_CUDA(cudaMemcpyAsync(status + first, w->d_edit_status, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, s));
        if (first + count < n) GC_CUDA(cudaStreamSynchronize(s));   // `edits` of the next pass overwrite the device copy
    }
    uint32_t* d_count = w->d_edit_state + EP_REBUILD_COUNT;
    GC_CUDA(launch_edit_list(a.w, w->d_edit_dirty, w->d_edit_coords, d_count, s));
    GC_CUDA(cudaMemcpyAsync(w->h_edit_count, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    // the first coordinates travel with the count (an edit usually flags a handful of chunks): one synchronisation
    const uint32_t eager = w->n_slots < 64 ? (uint32_t)w->n_slots : 64u;
    GC_CUDA(cudaMemcpyAsync(w->h_edit_coords, w->d_edit_coords, sizeof(GcChunkCoord) * eager, cudaMemcpyDeviceToHost, s));
    GC_CUDA(cudaStreamSynchronize(s));
    const uint32_t flagged = *w->h_edit_count;
    if (flagged)
    {
        if (flagged > eager)
        {
            GC_CUDA(cudaMemcpyAsync(w->h_edit_coords + eager, w->d_edit_coords + eager, sizeof(GcChunkCoord) * (flagged - eager), cudaMemcpyDeviceToHost, s));
            GC_CUDA(cudaStreamSynchronize(s));
        }
        std::sort(w->h_edit_coords, w->h_edit_coords + flagged, [](const GcChunkCoord& p, const GcChunkCoord& q)
        {
            if (p.cz != q.cz) return p.cz < q.cz;
            if (p.cx != q.cx) return p.cx < q.cx;
            return p.cy < q.cy;
        });
    }
    // the sparse-upload bookkeeping no longer knows what the device arrays hold
    memset(w->slot_dirty, 7, w->n_slots);
    if (n_rebuild) *n_rebuild = (int)flagged;
    const uint32_t stored = flagged < (uint32_t)rebuild_capacity ? flagged : (uint32_t)rebuild_capacity;
    if (stored) memcpy(rebuild, w->h_edit_coords, sizeof(GcChunkCoord) * stored);
    if (flagged > (uint32_t)rebuild_capacity) return fail(GC_ERR_CAPACITY, "more chunks flagged for rebuild than the list holds");
    return GC_OK;
}

int gcmesh_world_vertex_counts(GcWorld* w, int32_t* counts)
{
    if (!w || !counts) return fail(GC_ERR_ARG, "null argument");
    GC_CUDA(cudaSetDevice(w->ctx->device));
    GC_CUDA(cudaMemcpyAsync(counts, w->d_vert_table, sizeof(int32_t) * w->n_slots, cudaMemcpyDeviceToHost, w->ctx->stream));
    GC_CUDA(cudaStreamSynchronize(w->ctx->stream));
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Region records (the reference's saved-chunk format, WorldIO.cpp:167-307) in and out of the device world.
int gcmesh_world_upload_rle(GcWorld* w, const GcRleChunk* chunks, int n, const uint16_t* data, size_t data_words)
{
    if (!w || (!chunks && n > 0) || (!data && data_words > 0)) return fail(GC_ERR_ARG, "null argument");
    if (n < 0) return fail(GC_ERR_ARG, "negative record count");
    if (n == 0) return GC_OK;
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if ((size_t)n > w->rle_jobs_cap)
    {
        cudaFree(w->d_rle_jobs); if (w->h_rle_jobs) cudaFreeHost(w->h_rle_jobs);
        w->d_rle_jobs = nullptr; w->h_rle_jobs = nullptr; w->rle_jobs_cap = 0;
        const size_t cap = (size_t)n + 1024;
        GC_CUDA(cudaMalloc(&w->d_rle_jobs, cap * sizeof(RleJob)));
        GC_CUDA(cudaMallocHost(&w->h_rle_jobs, cap * sizeof(RleJob)));
        w->rle_jobs_cap = cap;
    }
    if (!w->d_rle_bad)
    {
        GC_CUDA(cudaMalloc(&w->d_rle_bad, sizeof(uint32_t)));
        GC_CUDA(cudaMallocHost(&w->h_rle_bad, sizeof(uint32_t)));
    }
    // the staging list of the previous call may still be in flight
    GC_CUDA(cudaStreamSynchronize(s));
    size_t pairs = 0;
    for (int i = 0; i < n; i++)
    {
        const GcRleChunk& c = chunks[i];
        if (!column_ok(w, c.cx, c.cz) || c.cy < 0 || c.cy >= GC_WORLD_CHUNK_HEIGHT) return fail(GC_ERR_ARG, "record for a chunk outside the window");
        if ((c.first_word & 1u) || c.words < 4 || (c.words & 1u) || (size_t)c.first_word + c.words > data_words)
            return fail(GC_ERR_ARG, "record outside the stream, odd-aligned or shorter than one run");
        RleJob& j = w->h_rle_jobs[i];
        j.slot = (uint32_t)(((size_t)c.cz * w->size_x + c.cx) * GC_WORLD_CHUNK_HEIGHT + c.cy);
        j.first_word = c.first_word; j.words = c.words; j.pair_base = (uint32_t)pairs;
        pairs += (c.words - 2) / 2;
        w->slot_dirty[j.slot] |= 1;
    }
    if (grow_device(&w->d_rle_data, &w->rle_data_cap, data_words) != GC_OK) return GC_ERR_CUDA;
    if (grow_device(&w->d_rle_ends, &w->rle_ends_cap, pairs) != GC_OK) return GC_ERR_CUDA;
    GC_CUDA(cudaMemcpyAsync(w->d_rle_data, data, data_words * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    GC_CUDA(cudaMemcpyAsync(w->d_rle_jobs, w->h_rle_jobs, (size_t)n * sizeof(RleJob), cudaMemcpyHostToDevice, s));
    GC_CUDA(cudaMemsetAsync(w->d_rle_bad, 0, sizeof(uint32_t), s));
    GC_CUDA(launch_rle_decode(w->d_rle_jobs, n, w->d_rle_data, w->d_rle_ends, w->d_blocks, w->d_rle_bad, s));
    GC_CUDA(cudaMemcpyAsync(w->h_rle_bad, w->d_rle_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return GC_OK;
}

int gcmesh_world_rle_refused(GcWorld* w, int* refused)
{
    if (!w || !refused) return fail(GC_ERR_ARG, "null argument");
    *refused = 0;
    if (!w->h_rle_bad) return GC_OK;
    GC_CUDA(cudaSetDevice(w->ctx->device));
    GC_CUDA(cudaStreamSynchronize(w->ctx->stream));
    *refused = (int)*w->h_rle_bad;
    return GC_OK;
}

int gcmesh_world_encode_rle(GcWorld* w, const GcChunkCoord* coords, int n, uint16_t* records, size_t capacity_words, GcRleChunk* out)
{
    if (!w || (!coords && n > 0) || !records || !out) return fail(GC_ERR_ARG, "null argument");
    if (n < 0) return fail(GC_ERR_ARG, "negative chunk count");
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    constexpr int kRleGroup = 256;                            // bricks encoded per launch (worst-case record: 256 KB each)
    if (!w->d_rle_out)
    {
        GC_CUDA(cudaMalloc(&w->d_rle_out, (size_t)kRleGroup * kRleMaxWords * sizeof(uint16_t)));
        GC_CUDA(cudaMalloc(&w->d_rle_dense, (size_t)kRleGroup * kRleMaxWords * sizeof(uint16_t)));
        GC_CUDA(cudaMalloc(&w->d_rle_words, (kRleGroup + 1) * sizeof(uint32_t)));      // [kRleGroup] = the group's total
        GC_CUDA(cudaMalloc(&w->d_rle_slots, kRleGroup * sizeof(uint32_t)));
        GC_CUDA(cudaMalloc(&w->d_rle_offsets, kRleGroup * sizeof(uint16_t)));
        GC_CUDA(cudaMallocHost(&w->h_rle_words, (kRleGroup + 1) * sizeof(uint32_t) + kRleGroup * (sizeof(uint32_t) + sizeof(uint16_t))));
    }
    uint32_t* h_words = w->h_rle_words;
    uint32_t* h_slots = h_words + kRleGroup + 1;
    uint16_t* h_offsets = reinterpret_cast<uint16_t*>(h_slots + kRleGroup);
    size_t used = 0;
    for (int g0 = 0; g0 < n; g0 += kRleGroup)
    {
        const int m = n - g0 < kRleGroup ? n - g0 : kRleGroup;
        for (int i = 0; i < m; i++)
        {
            const GcChunkCoord c = coords[g0 + i];
            if (!column_ok(w, c.cx, c.cz) || c.cy < 0 || c.cy >= GC_WORLD_CHUNK_HEIGHT) return fail(GC_ERR_ARG, "chunk outside the window");
            h_slots[i] = (uint32_t)(((size_t)c.cz * w->size_x + c.cx) * GC_WORLD_CHUNK_HEIGHT + c.cy);
            // RegionIndex of (ChunkP & REGION_MASK) (WorldIO.cpp:5-8,259-260): the chunk's WORLD position — LoadRegionFile files the
            // record under this word (WorldIO.cpp:48-62)
            const int wx = w->origin_cx + c.cx, wz = w->origin_cz + c.cz;
            h_offsets[i] = (uint16_t)((wx & 7) + 8 * (c.cy + GC_WORLD_CHUNK_HEIGHT * (wz & 7)));
        }
        GC_CUDA(cudaMemcpyAsync(w->d_rle_slots, h_slots, m * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        GC_CUDA(cudaMemcpyAsync(w->d_rle_offsets, h_offsets, m * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        GC_CUDA(launch_rle_encode(w->d_rle_slots, w->d_rle_offsets, m, w->d_blocks, w->d_rle_out, w->d_rle_words, s));
        // the group's records packed back to back on the device, then one copy
        GC_CUDA(launch_rle_pack(w->d_rle_out, w->d_rle_words, m, w->d_rle_dense, w->d_rle_words + kRleGroup, s));
        GC_CUDA(cudaMemcpyAsync(h_words, w->d_rle_words, (kRleGroup + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        GC_CUDA(cudaStreamSynchronize(s));
        const uint32_t total = h_words[kRleGroup];
        if (used + total > capacity_words) return fail(GC_ERR_CAPACITY, "record buffer too small");
        GC_CUDA(cudaMemcpyAsync(records + used, w->d_rle_dense, (size_t)total * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
        for (int i = 0; i < m; i++)
        {
            const GcChunkCoord c = coords[g0 + i];
            out[g0 + i].cx = c.cx; out[g0 + i].cy = c.cy; out[g0 + i].cz = c.cz;
            out[g0 + i].first_word = (uint32_t)used; out[g0 + i].words = h_words[i];
            used += h_words[i];
        }
        GC_CUDA(cudaStreamSynchronize(s));
    }
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
int gcmesh_batch_create(GcContext* ctx, int max_chunks, uint64_t max_quads, GcBatch** out)
{
    return gcmesh_batch_create_ex(ctx, max_chunks, max_quads, 0, out);
}

int gcmesh_batch_create_ex(GcContext* ctx, int max_chunks, uint64_t max_quads, uint32_t flags, GcBatch** out)
{
    if (!ctx || !out) return fail(GC_ERR_ARG, "null argument");
    *out = nullptr;
    if (flags & ~(uint32_t)GC_BATCH_UNCAPPED) return fail(GC_ERR_ARG, "unknown batch flag");
    if (max_chunks < 1) return fail(GC_ERR_ARG, "max_chunks must be positive");
    if (max_quads < 1) max_quads = 1;
    if (max_quads > 0xFFFFFFFFull) return fail(GC_ERR_ARG, "max_quads must fit 32 bits (GcChunkOut::quad_base)");
    GC_CUDA(cudaSetDevice(ctx->device));
    GcBatch* b = new (std::nothrow) GcBatch();
    if (!b) return fail(GC_ERR_ARG, "out of host memory");
    b->ctx = ctx; b->max_chunks = max_chunks; b->max_quads = max_quads;
    b->uncapped = (flags & GC_BATCH_UNCAPPED) != 0;
    b->index_quad_bytes = b->uncapped ? 24 : 12;
    cudaError_t e = cudaMalloc(&b->d_coords, sizeof(GcChunkCoord) * (size_t)max_chunks);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_out, sizeof(GcChunkOut) * (size_t)max_chunks);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_vertices, (size_t)max_quads * 48);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_indices, (size_t)max_quads * b->index_quad_bytes);
    // one 32-byte control block {quad cursor | work counter, error flag | CTAs done, pad | quads of the last launch}: cleared
    // here once — the last CTA of every launch leaves it ready for the next one — and read back once per submission
    if (e == cudaSuccess) e = cudaMalloc(&b->d_cursor, 32);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_cursor, 0, 32, ctx->stream);
    if (e == cudaSuccess) b->d_counters = reinterpret_cast<int32_t*>(b->d_cursor + 1);
    if (e == cudaSuccess && getenv("GCMESH_PROFILE")) e = cudaMalloc(&b->d_prof, sizeof(unsigned long long) * 512);
    if (e == cudaSuccess) e = cudaMallocHost(&b->h_out, sizeof(GcChunkOut) * (size_t)max_chunks);
    if (e == cudaSuccess) e = cudaMallocHost(&b->h_coords, sizeof(GcChunkCoord) * (size_t)max_chunks);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_order, sizeof(int32_t) * (size_t)max_chunks);
    if (e == cudaSuccess) e = cudaMallocHost(&b->h_order, sizeof(int32_t) * (size_t)max_chunks);
    if (e == cudaSuccess) { b->order_coords = (GcChunkCoord*)malloc(sizeof(GcChunkCoord) * (size_t)max_chunks); if (!b->order_coords) e = cudaErrorMemoryAllocation; }
    if (e == cudaSuccess) e = cudaMallocHost(&b->h_cursor, 32);
    if (e == cudaSuccess) b->h_counters = reinterpret_cast<int32_t*>(b->h_cursor + 1);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_stop);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_coords, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_kernel, cudaEventDisableTiming);
    if (e != cudaSuccess) { gcmesh_batch_destroy(b); return fail(GC_ERR_CUDA, "batch allocation", cudaGetErrorString(e)); }
    b->order_rows[0] = b->order_rows[1] = -1;
    b->timing = true;
    *out = b;
    return GC_OK;
}

int gcmesh_batch_destroy(GcBatch* b)
{
    if (!b) return GC_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaStreamSynchronize(b->ctx->download_stream);
    cudaFree(b->d_order); if (b->h_order) cudaFreeHost(b->h_order); free(b->order_coords); cudaFree(b->d_fill_jobs);
    cudaFree(b->d_coords); cudaFree(b->d_out); cudaFree(b->d_vertices); cudaFree(b->d_indices); cudaFree(b->d_cursor); cudaFree(b->d_prof);
    if (b->h_out) cudaFreeHost(b->h_out);
    if (b->h_coords) cudaFreeHost(b->h_coords);
    if (b->h_cursor) cudaFreeHost(b->h_cursor);
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    if (b->ev_stop) cudaEventDestroy(b->ev_stop);
    if (b->ev_done) cudaEventDestroy(b->ev_done);
    if (b->ev_coords) cudaEventDestroy(b->ev_coords);
    if (b->ev_kernel) cudaEventDestroy(b->ev_kernel);
    delete b;
    return GC_OK;
}

static int validate_coords(const GcWorld* w, const GcChunkCoord* coords, int n)
{
    for (int i = 0; i < n; i++)
    {
        const GcChunkCoord c = coords[i];
        // edge columns are inputs only (WorldRender.cpp:235); chunk rows are 0..3 (World.h:21)
        if (c.cx < 1 || c.cx > w->size_x - 2 || c.cz < 1 || c.cz > w->size_z - 2 || c.cy < 0 || c.cy >= GC_WORLD_CHUNK_HEIGHT)
            return fail(GC_ERR_ARG, "chunk on the window edge or outside the window");
    }
    return GC_OK;
}

// Launch the mesh kernel over d_coords[first, first + count) on stream s; descriptors go to d_out[first ...].  The last CTA of
// a launch resets the work counter and (unless keep_cursor) the quad cursor, and leaves the cursor's value in word 3 of the
// control block: no memset between launches.
static cudaError_t launch_range(GcWorld* w, GcBatch* b, int first, int count, cudaStream_t s, bool keep_cursor = false, const int32_t* order = nullptr,
                                const HaloFused* halo = nullptr, const uint8_t* sun = nullptr, const uint8_t* light = nullptr, unsigned long long* total_host = nullptr)
{
    GcContext* ctx = w->ctx;
    if (count <= 0) return cudaSuccess;
    MeshParams p;
    p.blocks = w->d_blocks; p.sun = sun ? sun : w->d_sun; p.light = light ? light : w->d_light; p.surface = w->d_surface;
    p.size_x = w->size_x; p.size_z = w->size_z;
    p.coords = b->d_coords + first; p.order = order; p.n_chunks = count; p.out = b->d_out + first;
    p.vertex_arena = b->d_vertices; p.index_arena = b->d_indices; p.max_quads = b->max_quads;
    p.quad_cursor = b->d_cursor; p.work_counter = b->d_counters; p.tables = ctx->d_tables;
    p.error_flag = b->d_counters + 1; p.use_tma = ctx->use_tma; p.prof = b->d_prof;
    p.done_counter = reinterpret_cast<uint32_t*>(b->d_cursor + 2); p.total_out = b->d_cursor + 3; p.keep_cursor = keep_cursor ? 1 : 0;
    p.total_out_host = total_host;
    p.vert_table = w->d_vert_table;
    p.wide = b->uncapped ? 1 : 0;
    p.trust_surface = w->surface_bounds_mesh ? 1 : 0;
    p.n_push = 0;
    memset(&p.halo, 0, sizeof(p.halo));
    if (halo && halo->enabled)
    {
        p.halo = *halo;
        p.n_push = halo->tasks_per_dir * ((halo->peer[0].connected ? 1 : 0) + (halo->peer[1].connected ? 1 : 0));
    }
    const int items = count + p.n_push;
    const int grid = items < ctx->mesh_ctas ? items : ctx->mesh_ctas;
    cudaError_t e = launch_mesh(p, w->maps, grid, s);
    if (e == cudaSuccess) b->launches++;
    return e;
}

// gcmesh_submit / gcmesh_submit_exchange.  rows: null, or {send_row_down, send_row_up} for the fused halo exchange.
static int submit_impl(GcWorld* w, GcBatch* b, const GcChunkCoord* coords, int n, const int* rows)
{
    if (!w || !b || (!coords && n > 0)) return fail(GC_ERR_ARG, "null argument");
    if (w->ctx != b->ctx) return fail(GC_ERR_ARG, "world and batch belong to different contexts");
    if (n < 0) return fail(GC_ERR_ARG, "negative chunk count");
    if (n > b->max_chunks) return fail(GC_ERR_CAPACITY, "batch created for fewer chunks");
    if (validate_coords(w, coords, n) != GC_OK) return GC_ERR_ARG;
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    HaloFused halo;
    memset(&halo, 0, sizeof(halo));
    if (rows)
    {
        if (b->uncapped) return fail(GC_ERR_ARG, "an uncapped batch cannot carry the halo exchange");
        if (!w->halo_peer[0].connected && !w->halo_peer[1].connected) return fail(GC_ERR_ARG, "no neighbour connected");
        if ((w->halo_peer[0].connected && (rows[0] < 1 || rows[0] >= w->size_z - 1)) || (w->halo_peer[1].connected && (rows[1] < 1 || rows[1] >= w->size_z - 1)))
            return fail(GC_ERR_ARG, "send row outside the meshable rows of the window");
        if (n == 0) return fail(GC_ERR_ARG, "an exchange needs at least one chunk to mesh (every rank of a chain launches one kernel per exchange)");
        static const int tasks_env = getenv("GCMESH_HALO_TASKS") ? atoi(getenv("GCMESH_HALO_TASKS")) : 0;
        halo.enabled = 1;
        halo.tasks_per_dir = tasks_env > 0 ? tasks_env : 16;
        if (halo.tasks_per_dir > w->size_x * 4) halo.tasks_per_dir = w->size_x * 4;
        halo.blocks = w->d_blocks; halo.sun = w->d_sun; halo.light = w->d_light; halo.surface = w->d_surface; halo.size_x = w->size_x;
        halo.send_row[0] = rows[0]; halo.send_row[1] = rows[1];
        halo.peer[0] = w->halo_peer[0]; halo.peer[1] = w->halo_peer[1];
        halo.flags = w->d_halo_flags;
        halo.wait_ns = halo_wait_ns();
        static const int debug_env = getenv("GCMESH_HALO_DEBUG") ? atoi(getenv("GCMESH_HALO_DEBUG")) : 0;
        halo.debug = debug_env;
    }
    // the pinned coordinate staging buffer is reused: wait until the previous submission has consumed it (its kernel
    // may still be queued or running, so back-to-back submissions keep the GPU busy)
    if (b->state == 1) GC_CUDA(cudaEventSynchronize(b->ev_coords));
    // Work order (longest first): a chunk with thousands of quads costs several times an air chunk, and the chunks claimed
    // last decide when the kernel ends.  With the quad counts of a finished submission of the same list the order is by
    // those; otherwise lower chunks first (terrain sits at the bottom of the world, the sky above it is air).  With a fused
    // exchange the heavy chunks of the boundary rows come after the heavy chunks of all other rows — by then the neighbours'
    // planes have long arrived — and the light (air) chunks of all rows still make up the tail.
    const int r0 = rows && w->halo_peer[0].connected ? rows[0] : -1, r1 = rows && w->halo_peer[1].connected ? rows[1] : -1;
    const bool same_list = n == b->order_n && n > 0 && r0 == b->order_rows[0] && r1 == b->order_rows[1] &&
                           memcmp(coords, b->order_coords, sizeof(GcChunkCoord) * (size_t)n) == 0;
    const bool order_reused = same_list && (b->order_from_history || b->state != 2);
    if (!order_reused)
    {
        constexpr int kBuckets = 24;
        const bool history = same_list && b->state == 2;
        int start[3 * kBuckets + 1] = { 0 };
        auto bucket = [&](int i) -> int
        {
            const bool boundary = coords[i].cz == r0 || coords[i].cz == r1;
            int weight; bool heavy;
            if (!history) { const int cy = coords[i].cy < 4 ? coords[i].cy : 3; weight = kBuckets - 1 - cy; heavy = cy < 2; }   // cy 0 first
            else
            {
                const uint32_t q = b->h_out[i].quads;
                weight = q == 0 ? 0 : 1 + (q >= 11264 ? 22 : (int)(q >> 9));                                             // 512-quad classes
                heavy = q != 0;
            }
            return (heavy ? (boundary ? 1 : 2) : 0) * kBuckets + weight;
        };
        for (int i = 0; i < n; i++) start[bucket(i) + 1]++;
        for (int k = 0; k < 3 * kBuckets; k++) start[k + 1] += start[k];
        for (int i = 0; i < n; i++) b->h_order[n - 1 - start[bucket(i)]++] = i;             // descending bucket
        memcpy(b->order_coords, coords, sizeof(GcChunkCoord) * (size_t)n);
        b->order_n = n; b->order_from_history = history; b->order_rows[0] = r0; b->order_rows[1] = r1;
    }
    b->n = n;
    b->launches = 0;
    memcpy(b->h_coords, coords, sizeof(GcChunkCoord) * (size_t)n);
    cudaStream_t s = ctx->stream;
    // (a list identical to the previous submission's is already on the device)
    if (!same_list) GC_CUDA(cudaMemcpyAsync(b->d_coords, b->h_coords, sizeof(GcChunkCoord) * (size_t)n, cudaMemcpyHostToDevice, s));
    if (!order_reused) GC_CUDA(cudaMemcpyAsync(b->d_order, b->h_order, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    if (!same_list || !order_reused) GC_CUDA(cudaEventRecord(b->ev_coords, s));
    if (n == 0) GC_CUDA(cudaMemsetAsync(b->d_cursor + 3, 0, sizeof(unsigned long long), s));
    if (b->d_prof) GC_CUDA(cudaMemsetAsync(b->d_prof, 0, sizeof(unsigned long long) * 512, s));
    b->timed = b->timing;
    if (b->timing) GC_CUDA(cudaEventRecord(b->ev_start, s));
    if (halo.enabled)
    {
        halo.serial = ++w->halo_serial;
        // the neighbours' planes land in this window's halo rows: the sparse-upload bookkeeping no longer knows them
        for (int d = 0; d < 2; d++)
            if (w->halo_peer[d].connected)
            {
                const int row = d == 0 ? rows[0] - 1 : rows[1] + 1;
                memset(w->slot_dirty + (size_t)row * w->size_x * GC_WORLD_CHUNK_HEIGHT, 7, (size_t)w->size_x * GC_WORLD_CHUNK_HEIGHT);
            }
    }
    GC_CUDA(launch_range(w, b, 0, n, s, false, b->d_order, &halo));
    if (b->timing) GC_CUDA(cudaEventRecord(b->ev_stop, s));
    GC_CUDA(cudaEventRecord(b->ev_kernel, s));
    // the descriptors travel back on the download stream: the context stream is free for the next submission's kernel
    cudaStream_t sd = ctx->download_stream;
    GC_CUDA(cudaStreamWaitEvent(sd, b->ev_kernel, 0));
    GC_CUDA(cudaMemcpyAsync(b->h_out, b->d_out, sizeof(GcChunkOut) * (size_t)n, cudaMemcpyDeviceToHost, sd));
    GC_CUDA(cudaMemcpyAsync(b->h_cursor, b->d_cursor, 32, cudaMemcpyDeviceToHost, sd));
    if (b->d_prof) GC_CUDA(cudaMemcpyAsync(b->h_prof, b->d_prof, sizeof(unsigned long long) * 512, cudaMemcpyDeviceToHost, sd));
    GC_CUDA(cudaEventRecord(b->ev_done, sd));
    b->state = 1;
    return GC_OK;
}

int gcmesh_submit(GcWorld* w, GcBatch* b, const GcChunkCoord* coords, int n) { return submit_impl(w, b, coords, n, nullptr); }

int gcmesh_submit_exchange(GcWorld* w, GcBatch* b, const GcChunkCoord* coords, int n, int send_row_down, int send_row_up)
{
    const int rows[2] = { send_row_down, send_row_up };
    return submit_impl(w, b, coords, n, rows);
}

int gcmesh_batch_set_timing(GcBatch* b, int on)
{
    if (!b) return fail(GC_ERR_ARG, "batch is null");
    b->timing = on != 0;
    return GC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Host arrays in, host meshes out — the call a host-side integration makes per frame / per rebuild list.
// The window is processed in strips of rows so that the zero-scan of strip k+1 (host threads), the PCIe pull of strip
// k (upload stream), the meshing of strip k-1 (context stream) and the read-back of strip k-2 (download stream) overlap.
int gcmesh_mesh_host(GcWorld* w, GcBatch* b, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light,
                     const uint8_t* surface, const uint8_t* nonzero_hint, const GcChunkCoord* coords, int n,
                     void* host_vertices, void* host_indices, uint64_t host_capacity_quads, int threads)
{
    if (!w || !b || !blocks || !surface || (!coords && n > 0)) return fail(GC_ERR_ARG, "null argument");
    if ((sunlight == nullptr) != (block_light == nullptr)) return fail(GC_ERR_ARG, "hand in both light arrays or neither");
    // neither light array: the light is filled on the device from the block ids and the surface map (what PreprocessGroup
    // computes, WorldRender.cpp:240-262), group of rows by group of rows, and only block ids cross the link
    const bool device_light = sunlight == nullptr;
    if (device_light && (size_t)w->size_x * GC_WORLD_CHUNK_HEIGHT * 16 > 65535) return fail(GC_ERR_CAPACITY, "window too wide for the light fill (32-bit voxel addresses)");
    if (w->ctx != b->ctx) return fail(GC_ERR_ARG, "world and batch belong to different contexts");
    if (n < 0 || n > b->max_chunks) return fail(GC_ERR_CAPACITY, "batch created for fewer chunks");
    if (validate_coords(w, coords, n) != GC_OK) return GC_ERR_ARG;
    GcContext* ctx = w->ctx;
    GC_CUDA(cudaSetDevice(ctx->device));
    if (b->state == 1) GC_CUDA(cudaEventSynchronize(b->ev_done));

    // short strips keep the tail (the last strip's upload -> mesh -> read-back, which nothing overlaps) small
    static const int strip_rows_env = getenv("GCMESH_STRIP_ROWS") ? atoi(getenv("GCMESH_STRIP_ROWS")) : 0;
    int kStripRows = strip_rows_env > 0 ? strip_rows_env : 2;
    while ((w->size_z + kStripRows - 1) / kStripRows > 64) kStripRows++;
    const int n_strips = (w->size_z + kStripRows - 1) / kStripRows;

    // chunks bucketed by strip and, inside a strip, lower chunks first (the heavy ones: terrain sits at the bottom of the
    // world); `perm` maps the sorted position back to the caller's order
    std::vector<int> first(n_strips + 1, 0), perm((size_t)n);
    {
        std::vector<int> sub((size_t)n_strips * 4 + 1, 0);
        auto key = [&](int i) { return (coords[i].cz / kStripRows) * 4 + coords[i].cy; };
        for (int i = 0; i < n; i++) sub[key(i) + 1]++;
        for (size_t k = 0; k + 1 < sub.size(); k++) sub[k + 1] += sub[k];
        for (int k = 0; k <= n_strips; k++) first[k] = sub[(size_t)k * 4];
        for (int i = 0; i < n; i++) { const int pos = sub[key(i)]++; perm[pos] = i; b->h_coords[pos] = coords[i]; }
    }

    cudaStream_t sm = ctx->stream, su = ctx->upload_stream, su2 = ctx->upload_stream2, sd = ctx->download_stream;
    const bool mapped = host_arrays_mapped(blocks, sunlight, block_light);
    b->n = n; b->launches = 0; w->launches = 0;
    w->upload_bytes = w->n_cols * GC_CHUNK_SIZE_2 + sizeof(GcChunkCoord) * (size_t)n;
    b->order_n = 0;                                           // d_coords is about to hold the strip-sorted list
    GC_CUDA(cudaEventRecord(ctx->ev_fork, sm));
    GC_CUDA(cudaStreamWaitEvent(su, ctx->ev_fork, 0));        // the upload overwrites what earlier work on the context stream read
    GC_CUDA(cudaStreamWaitEvent(su2, ctx->ev_fork, 0));
    GC_CUDA(cudaStreamWaitEvent(sd, ctx->ev_fork, 0));
    GC_CUDA(cudaMemcpyAsync(b->d_coords, b->h_coords, sizeof(GcChunkCoord) * (size_t)n, cudaMemcpyHostToDevice, su));
    GC_CUDA(cudaMemcpyAsync(w->d_surface, surface, w->n_cols * GC_CHUNK_SIZE_2, cudaMemcpyHostToDevice, su));
    // (the quad cursor and the work counter are zero: the last CTA of every launch leaves them so; the strips of this call keep
    // the cursor running and the last one resets it)
    GC_CUDA(cudaMemsetAsync(b->d_cursor + 3, 0, sizeof(unsigned long long), su));
    GC_CUDA(cudaEventRecord(b->ev_start, sm));
    b->timed = true;
    int last_nonempty = -1;
    for (int j = 0; j < n_strips; j++)
        if (first[j + 1] > first[j]) last_nonempty = j;

    unsigned long long* cursors = ctx->h_strip_cursor;        // pinned: arena cursor after each strip's kernel, stored by the kernel itself
    unsigned long long* cursors_dev = nullptr;
    GC_CUDA(cudaHostGetDevicePointer((void**)&cursors_dev, cursors, 0));
    uint64_t copied = 0;
    std::vector<char> no_cursor((size_t)n_strips, 0);
    for (int j = 0; j < n_strips; j++) no_cursor[j] = first[j + 1] == first[j];
    auto read_back = [&](int j) -> int                        // strip j has been meshed: copy its arena range to the host buffers
    {
        GC_CUDA(cudaEventSynchronize(ctx->ev_meshed[j]));
        if (no_cursor[j]) return GC_OK;                       // (a strip without chunks — the window's edge rows — launched nothing; a strip meshed by its group's one launch has no cursor of its own)
        uint64_t end = cursors[j] < b->max_quads ? cursors[j] : b->max_quads;
        if (end > host_capacity_quads) end = host_capacity_quads;
        if (end > copied)
        {
            if (host_vertices) GC_CUDA(cudaMemcpyAsync((uint8_t*)host_vertices + copied * 48, b->d_vertices + copied * 48, (size_t)(end - copied) * 48, cudaMemcpyDeviceToHost, sd));
            if (host_indices) GC_CUDA(cudaMemcpyAsync((uint8_t*)host_indices + copied * b->index_quad_bytes, (uint8_t*)b->d_indices + copied * b->index_quad_bytes,
                                                      (size_t)(end - copied) * b->index_quad_bytes, cudaMemcpyDeviceToHost, sd));
            copied = end;
        }
        return GC_OK;
    };
    auto mesh_strip = [&](int j) -> int                       // strips <= j + 1 are on the device (or queued on the upload stream)
    {
        const int up = j + 1 < n_strips ? j + 1 : n_strips - 1;
        GC_CUDA(cudaStreamWaitEvent(sm, ctx->ev_uploaded[up], 0));
        if (first[j + 1] > first[j])
            GC_CUDA(launch_range(w, b, first[j], first[j + 1] - first[j], sm, j != last_nonempty, nullptr, nullptr, nullptr, nullptr, cursors_dev + j));
        GC_CUDA(cudaEventRecord(ctx->ev_meshed[j], sm));
        return GC_OK;
    };

    // ---- device light: groups of strips ----
    // A group's light is filled into a scratch pair of arenas over the rows [z0 - 2, z1 + 2) around its rows [z0, z1) seen as a
    // window of their own: its edge rows only receive, the rows next to them are seeded like every inner column — light reaches
    // less than a column (15 voxels), so the cells the group's chunks read (one voxel into rows z0 - 1 and z1) hold exactly
    // what the fill of the whole window gives them.  The group's chunks are then meshed against the scratch arenas (base pointers
    // shifted so that window slots address them).  Fill and meshing follow each other on the context stream, so one scratch
    // pair serves every group; uploads of the next group and read-backs of the previous one run beside them.
    static const int group_rows_env = getenv("GCMESH_LIGHT_GROUP_ROWS") ? atoi(getenv("GCMESH_LIGHT_GROUP_ROWS")) : 0;
    const int group_strips = device_light ? std::max(1, (group_rows_env > 0 ? group_rows_env : 6) / kStripRows) : 0;
    // the first groups are short — one strip, then two — so that the first meshes start their way back early; the read-back
    // is the longest leg of the call
    std::vector<int> gfirst;                                     // group g = strips [gfirst[g], gfirst[g + 1])
    if (device_light)
        for (int s0 = 0, len = 1; s0 < n_strips;)
        {
            gfirst.push_back(s0);
            s0 += len;
            // ... and so are the last ones (two strips, then one): what follows the last upload — fill, meshing, read-back of the
            // last group — is not overlapped by anything
            const int left = n_strips - s0;
            len = std::min(len + 1, group_strips);
            if (left <= 3) len = left == 3 ? 2 : 1;
            else if (left - len < 3 && left - len > 0) len = left - 3;
            if (len < 1) len = 1;
        }
    const int n_groups = (int)gfirst.size();
    gfirst.push_back(n_strips);
    if (device_light)
    {
        const size_t need = (size_t)(group_strips * kStripRows + 4) * w->size_x * GC_WORLD_CHUNK_HEIGHT;
        if (need > 65535) return fail(GC_ERR_CAPACITY, "light group too large for the light fill (32-bit voxel addresses): lower GCMESH_LIGHT_GROUP_ROWS");
        if (w->group_light_slots < need)
        {
            cudaFree(w->d_group_sun); cudaFree(w->d_group_light); w->d_group_sun = w->d_group_light = nullptr; w->group_light_slots = 0;
            GC_CUDA(cudaMalloc(&w->d_group_sun, need * GC_CHUNK_SIZE_3));
            GC_CUDA(cudaMalloc(&w->d_group_light, need * GC_CHUNK_SIZE_3));
            w->group_light_slots = need;
        }
        if (light_scratch(w) != GC_OK) return GC_ERR_CUDA;
    }
    auto light_and_mesh_group = [&](int g) -> int                // the strips of group g and the one after it are uploaded (or queued)
    {
        const int s0 = gfirst[g], s1 = gfirst[g + 1];
        const int z0 = s0 * kStripRows, z1 = std::min(s1 * kStripRows, w->size_z);
        const int r0 = std::max(z0 - 2, 0), r1 = std::min(z1 + 2, w->size_z);
        const int up = std::min((r1 - 1) / kStripRows, n_strips - 1);
        GC_CUDA(cudaStreamWaitEvent(sm, ctx->ev_uploaded[up], 0));
        bool any = false;
        for (int j = s0; j < s1; j++) any = any || first[j + 1] > first[j];
        const size_t slot0 = (size_t)r0 * w->size_x * GC_WORLD_CHUNK_HEIGHT, n_sub = (size_t)(r1 - r0) * w->size_x * GC_WORLD_CHUNK_HEIGHT;
        if (any)
        {
            GC_CUDA(launch_zero(w->d_group_sun, n_sub * GC_CHUNK_SIZE_3, ctx->num_sms, sm));
            GC_CUDA(launch_zero(w->d_group_light, n_sub * GC_CHUNK_SIZE_3, ctx->num_sms, sm));
            LightWorld lw = light_world(w);
            lw.blocks = w->d_blocks + slot0 * GC_CHUNK_SIZE_3; lw.surface = w->d_surface + (slot0 / GC_WORLD_CHUNK_HEIGHT) * GC_CHUNK_SIZE_2;
            lw.sun = w->d_group_sun; lw.light = w->d_group_light; lw.size_z = r1 - r0;
            if (enqueue_light_fill(w, lw, sm, 1) != GC_OK) return GC_ERR_CUDA;   // (one CTA per SM: uploads' pull kernels run beside it)
            b->launches += 2;
        }
        // one launch for the group's chunks (a strip of 288 chunks leaves each of the 296 persistent CTAs one chunk: as long as a
        // launch of three strips); its cursor is filed under the group's last strip
        if (any)
            GC_CUDA(launch_range(w, b, first[s0], first[s1] - first[s0], sm, !(last_nonempty >= s0 && last_nonempty < s1), nullptr, nullptr,
                                 w->d_group_sun - slot0 * GC_CHUNK_SIZE_3, w->d_group_light - slot0 * GC_CHUNK_SIZE_3, cursors_dev + (s1 - 1)));
        for (int j = s0; j < s1; j++)
        {
            no_cursor[j] = j != s1 - 1 || !any;
            GC_CUDA(cudaEventRecord(ctx->ev_meshed[j], sm));
        }
        return GC_OK;
    };

    const auto t_begin = std::chrono::steady_clock::now();
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    double t_scan_done[64] = { 0 };                           // (trace) when the scan threads finished each strip
    // the zero-scan runs on background threads over the whole window, slots in ascending order; the pipeline below
    // consumes it strip by strip (remaining[k] = slots of strip k still to classify)
    const size_t slots_per_row = (size_t)w->size_x * GC_WORLD_CHUNK_HEIGHT;
    std::vector<std::atomic<int>> remaining((size_t)n_strips);
    std::vector<std::thread> pool;
    ScanPlan plan(w, blocks, sunlight, block_light, surface);
    if (nonzero_hint) memcpy(w->scan_flags, nonzero_hint, w->n_slots);
    else
    {
        for (int k = 0; k < n_strips; k++)
        {
            const int cz0 = k * kStripRows, cz1 = cz0 + kStripRows < w->size_z ? cz0 + kStripRows : w->size_z;
            remaining[k].store((int)((cz1 - cz0) * slots_per_row));
        }
        // one core stays with this thread, which feeds the three streams as strips become ready
        int nt = scan_threads(threads);
        if (threads <= 0 && nt > 1) nt--;
        // no light arrays and block bricks bounded by the surface maps: the scan reads a few kilobytes per column, and starting
        // fifteen threads (this thread does, one after the other) would take longer than the scan itself
        if (device_light && w->surface_bounds_blocks && nt > 4) nt = 4;
        auto work = [&, slots_per_row, kStripRows]()
        {
            plan.work([&](size_t slot)
            {
                const size_t k = (slot / slots_per_row) / kStripRows;
                if (remaining[k].fetch_sub(1, std::memory_order_release) == 1) t_scan_done[k] = now_ms();
            });
        };
        for (int t = 0; t < nt; t++) pool.emplace_back(work);
    }
    struct Joiner { std::vector<std::thread>& p; ~Joiner() { for (auto& t : p) if (t.joinable()) t.join(); } } joiner{ pool };

    const bool trace = getenv("GCMESH_TRACE") != nullptr;
    double t_scanned[64], t_enqueued[64];
    int next_group = 0, next_read = 0;
    for (int k = 0; k < n_strips; k++)
    {
        const int cz0 = k * kStripRows, cz1 = cz0 + kStripRows < w->size_z ? cz0 + kStripRows : w->size_z;
        if (!nonzero_hint) while (remaining[k].load(std::memory_order_acquire) > 0) std::this_thread::yield();
        t_scanned[k] = now_ms();
        const int launched = enqueue_rows(w, cz0, cz1, blocks, sunlight, block_light, mapped, su, su2);
        if (launched < 0) return fail(GC_ERR_CUDA, "sparse upload", cudaGetErrorString(cudaGetLastError()));
        w->launches += launched;
        GC_CUDA(cudaEventRecord(ctx->ev_up2, su2));
        GC_CUDA(cudaStreamWaitEvent(su, ctx->ev_up2, 0));
        GC_CUDA(cudaEventRecord(ctx->ev_uploaded[k], su));
        t_enqueued[k] = now_ms();
        if (device_light)
        {
            // a group goes to the device once the strip behind its last row is queued; the group before it is read back then
            while (next_group < n_groups)
            {
                const int s1 = gfirst[next_group + 1];
                const int need = std::min((std::min(s1 * kStripRows, w->size_z) + 1) / kStripRows, n_strips - 1);
                if (need > k) break;
                { const int st = light_and_mesh_group(next_group); if (st != GC_OK) return st; }
                next_group++;
            }
            // strips that are meshed by now start their way back; this thread does not wait for any (it has uploads to queue)
            while (next_read < gfirst[next_group] && cudaEventQuery(ctx->ev_meshed[next_read]) == cudaSuccess) { const int st = read_back(next_read++); if (st != GC_OK) return st; }
            continue;
        }
        if (k >= 1) { const int st = mesh_strip(k - 1); if (st != GC_OK) return st; }
        // strips that are meshed by now start their way back; this thread does not wait for any of them here — it has uploads to
        // queue (waiting for strip k - 2 at this point ran upload, meshing and read-back in lock step)
        while (next_read < k && cudaEventQuery(ctx->ev_meshed[next_read]) == cudaSuccess) { const int st = read_back(next_read++); if (st != GC_OK) return st; }
    }
    if (device_light)
        while (next_group < n_groups) { const int st = light_and_mesh_group(next_group++); if (st != GC_OK) return st; }
    else { const int st = mesh_strip(n_strips - 1); if (st != GC_OK) return st; }
    GC_CUDA(cudaEventRecord(b->ev_stop, sm));
    GC_CUDA(cudaMemcpyAsync(b->h_out, b->d_out, sizeof(GcChunkOut) * (size_t)n, cudaMemcpyDeviceToHost, sm));
    GC_CUDA(cudaMemcpyAsync(b->h_cursor, b->d_cursor, 32, cudaMemcpyDeviceToHost, sm));
    while (next_read < n_strips) { const int st = read_back(next_read++); if (st != GC_OK) return st; }
    GC_CUDA(cudaEventRecord(ctx->ev_join, sd));
    GC_CUDA(cudaStreamWaitEvent(sm, ctx->ev_join, 0));
    GC_CUDA(cudaEventRecord(b->ev_done, sm));
    GC_CUDA(cudaEventRecord(b->ev_coords, sm));
    GC_CUDA(cudaEventSynchronize(b->ev_done));
    b->state = 2;
    if (trace)
    {
        fprintf(stderr, "[gcmesh_mesh_host] total %.2f ms host\n", now_ms());
        for (int k = 0; k < n_strips; k++)
        {
            float up = 0, me = 0;
            cudaEventElapsedTime(&up, b->ev_start, ctx->ev_uploaded[k]); cudaEventElapsedTime(&me, b->ev_start, ctx->ev_meshed[k]);
            fprintf(stderr, "  strip %2d: scan done %.2f  seen %.2f  enqueued %.2f | uploaded %.2f  meshed %.2f (device, since start)\n", k, t_scan_done[k], t_scanned[k], t_enqueued[k], up, me);
        }
        float dn = 0; cudaEventElapsedTime(&dn, b->ev_start, b->ev_stop);
        fprintf(stderr, "  last kernel done %.2f\n", dn);
    }
    if (b->h_counters[1] != 0) return fail(GC_ERR_CUDA, "mesh kernel watchdog fired (mbarrier wait did not complete)");

    // descriptors back into the caller's order
    std::vector<GcChunkOut> tmp(b->h_out, b->h_out + n);
    for (int pos = 0; pos < n; pos++) b->h_out[perm[pos]] = tmp[pos];
    return GC_OK;
}

int gcmesh_wait(GcBatch* b)
{
    if (!b) return fail(GC_ERR_ARG, "batch is null");
    if (b->state == 0) return fail(GC_ERR_STATE, "nothing submitted");
    GC_CUDA(cudaSetDevice(b->ctx->device));
    GC_CUDA(cudaEventSynchronize(b->ev_done));
    b->state = 2;
    if (b->h_counters[1] == 2) return fail(GC_ERR_CUDA, "halo exchange: a neighbour's boundary plane never arrived (chunks flagged GC_CHUNK_HALO_MISSING)");
    if (b->h_counters[1] != 0) return fail(GC_ERR_CUDA, "mesh kernel watchdog fired (mbarrier wait did not complete)");
    return GC_OK;
}

int gcmesh_query(GcBatch* b, int* done)
{
    if (!b || !done) return fail(GC_ERR_ARG, "null argument");
    *done = 0;
    if (b->state == 0) return fail(GC_ERR_STATE, "nothing submitted");
    if (b->state == 2) { *done = 1; return GC_OK; }
    GC_CUDA(cudaSetDevice(b->ctx->device));
    const cudaError_t e = cudaEventQuery(b->ev_done);
    if (e == cudaErrorNotReady) { cudaGetLastError(); return GC_OK; }
    if (e != cudaSuccess) return fail(GC_ERR_CUDA, "cudaEventQuery", cudaGetErrorString(e));
    *done = 1;
    return gcmesh_wait(b);          // (finished: the same bookkeeping and error checks as a wait that does not block)
}

int gcmesh_batch_results(GcBatch* b, const GcChunkOut** out, int* n, uint64_t* total_quads)
{
    if (!b) return fail(GC_ERR_ARG, "batch is null");
    if (b->state != 2) return fail(GC_ERR_STATE, "call gcmesh_wait first");
    if (out) *out = b->h_out;
    if (n) *n = b->n;
    if (total_quads) *total_quads = b->h_cursor[3] < b->max_quads ? b->h_cursor[3] : b->max_quads;
    return GC_OK;
}

int gcmesh_batch_download(GcBatch* b, void* vertices, void* indices)
{
    if (!b) return fail(GC_ERR_ARG, "batch is null");
    if (b->state != 2) return fail(GC_ERR_STATE, "call gcmesh_wait first");
    GC_CUDA(cudaSetDevice(b->ctx->device));
    const uint64_t q = b->h_cursor[3] < b->max_quads ? b->h_cursor[3] : b->max_quads;
    if (q == 0) return GC_OK;
    if (vertices) GC_CUDA(cudaMemcpyAsync(vertices, b->d_vertices, (size_t)q * 48, cudaMemcpyDeviceToHost, b->ctx->stream));
    if (indices) GC_CUDA(cudaMemcpyAsync(indices, b->d_indices, (size_t)q * b->index_quad_bytes, cudaMemcpyDeviceToHost, b->ctx->stream));
    GC_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return GC_OK;
}

int gcmesh_batch_fill_meshdata(GcBatch* b, int i, GcMeshData* out)
{
    if (!b || !out) return fail(GC_ERR_ARG, "null argument");
    if (b->state != 2) return fail(GC_ERR_STATE, "call gcmesh_wait first");
    if (i < 0 || i >= b->n) return fail(GC_ERR_ARG, "chunk index outside the batch");
    if (b->uncapped) return fail(GC_ERR_STATE, "GcMeshData is the reference's capped, 16-bit MeshData: not available from an uncapped batch");
    GC_CUDA(cudaSetDevice(b->ctx->device));
    const GcChunkOut& c = b->h_out[i];
    out->vert_count = (int32_t)c.vert_count;
    for (int t = 0; t < GC_MESH_TYPE_COUNT; t++) out->index_count[t] = (int32_t)c.index_count[t];
    if (c.vert_count == 0) return GC_OK;
    cudaStream_t s = b->ctx->stream;
    GC_CUDA(cudaMemcpyAsync(out->vertices, b->d_vertices + (size_t)c.quad_base * 48, (size_t)c.vert_count * 12, cudaMemcpyDeviceToHost, s));
    for (int t = 0; t < GC_MESH_TYPE_COUNT; t++)
        if (c.index_count[t])
            GC_CUDA(cudaMemcpyAsync(out->indices[t], b->d_indices + (size_t)c.quad_base * 6 + c.index_offset[t], (size_t)c.index_count[t] * 2, cudaMemcpyDeviceToHost, s));
    GC_CUDA(cudaStreamSynchronize(s));
    return GC_OK;
}

int gcmesh_batch_fill_device(GcBatch* b, const GcDeviceMesh* targets, int n)
{
    if (!b || (!targets && n > 0)) return fail(GC_ERR_ARG, "null argument");
    if (b->state != 2) return fail(GC_ERR_STATE, "call gcmesh_wait first");
    if (n < 0) return fail(GC_ERR_ARG, "negative target count");
    for (int i = 0; i < n; i++)
        if (targets[i].chunk < 0 || targets[i].chunk >= b->n) return fail(GC_ERR_ARG, "chunk index outside the batch");
    if (n == 0) return GC_OK;
    GC_CUDA(cudaSetDevice(b->ctx->device));
    if (grow_device(&b->d_fill_jobs, &b->fill_jobs_cap, (size_t)n) != GC_OK) return GC_ERR_CUDA;
    // resolved from the host copy of the descriptors (the caller's order, also after gcmesh_mesh_host)
    std::vector<FillJob> jobs((size_t)n);
    for (int i = 0; i < n; i++)
    {
        const GcChunkOut& c = b->h_out[targets[i].chunk];
        FillJob& j = jobs[(size_t)i];
        memset(&j, 0, sizeof(j));
        if (c.vert_count == 0) continue;
        if (targets[i].vertices)
        {
            j.src[0] = b->d_vertices + (size_t)c.quad_base * 48; j.dst[0] = targets[i].vertices; j.bytes[0] = c.vert_count * GC_VERTEX_BYTES;
        }
        for (int t = 0; t < GC_MESH_TYPE_COUNT; t++)
            if (targets[i].indices[t] && c.index_count[t])
            {
                const size_t ib = b->index_quad_bytes / 6;   // bytes per index
                j.src[1 + t] = (uint8_t*)b->d_indices + ((size_t)c.quad_base * 6 + c.index_offset[t]) * ib; j.dst[1 + t] = targets[i].indices[t];
                j.bytes[1 + t] = (uint32_t)(c.index_count[t] * ib);
            }
    }
    cudaStream_t s = b->ctx->stream;
    GC_CUDA(cudaMemcpyAsync(b->d_fill_jobs, jobs.data(), sizeof(FillJob) * (size_t)n, cudaMemcpyHostToDevice, s));   // pageable: staged before the call returns
    GC_CUDA(launch_fill_device(b->d_fill_jobs, n, s));
    return GC_OK;
}

int gcmesh_batch_device_ptrs(GcBatch* b, void** vertices, void** indices, void** chunk_out)
{
    if (!b) return fail(GC_ERR_ARG, "batch is null");
    if (vertices) *vertices = b->d_vertices;
    if (indices) *indices = b->d_indices;
    if (chunk_out) *chunk_out = b->d_out;
    return GC_OK;
}

int gcmesh_batch_elapsed_ms(GcBatch* b, float* ms)
{
    if (!b || !ms) return fail(GC_ERR_ARG, "null argument");
    if (b->state != 2) return fail(GC_ERR_STATE, "call gcmesh_wait first");
    if (!b->timed) return fail(GC_ERR_STATE, "the last submission was not timed (gcmesh_batch_set_timing)");
    GC_CUDA(cudaEventElapsedTime(ms, b->ev_start, b->ev_stop));
    return GC_OK;
}

int gcmesh_batch_launches(GcBatch* b) { return b ? b->launches : 0; }

// Development aid (GCMESH_PROFILE=1): the 512 in-kernel cycle counters of the last submission.
int gcmesh_batch_profile(GcBatch* b, unsigned long long* out512)
{
    if (!b || !out512) return fail(GC_ERR_ARG, "null argument");
    if (!b->d_prof) return fail(GC_ERR_STATE, "profiling is off (set GCMESH_PROFILE=1 before creating the batch)");
    memcpy(out512, b->h_prof, sizeof(b->h_prof));
    return GC_OK;
}

int gcmesh_kernel_info(int* regs, int* smem_bytes, int* threads)
{
    int r = 0, ss = 0, md = 0;
    cudaError_t e = mesh_kernel_attributes(&r, &ss, &md);
    if (e != cudaSuccess) return fail(GC_ERR_CUDA, "cudaFuncGetAttributes", cudaGetErrorString(e));
    if (regs) *regs = r;
    if (smem_bytes) *smem_bytes = (int)mesh_smem_bytes();
    if (threads) *threads = kMeshThreads;
    return GC_OK;
}

}  // extern "C"

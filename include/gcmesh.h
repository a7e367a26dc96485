/*
 * gcmesh — B200-native chunk mesher for Gamecraft: C ABI of libgcmesh.so.
 *
 * Drop-in boundary for ONE path of the reference: the per-chunk mesh build that the reference runs as
 * work-queue jobs,
 *     QueueAsync(state, BuildChunkAsync,    world, chunk,  OnChunkBuilt)      Code/WorldRender.cpp:69-78
 *     QueueAsync(state, RebuildChunksAsync, world, &list,  OnChunksRebuilt)   Code/WorldRender.cpp:80-91
 * (job type AsyncFunc, Code/Async.h:5-6).  Here the same work is ONE batched GPU submission:
 * chunk block / light arrays in, per chunk one 12-byte vertex stream + five uint16 index streams out,
 * byte-identical to BuildChunkAsync (Code/WorldRender.cpp:6-27) -> BuildBlock (:445-560).
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types in any signature (streams travel as void*).
 * There is no CPU fallback: every entry point that needs the GPU fails with GC_ERR_CUDA when none is usable.
 * All functions return 0 (GC_OK) or a negative GcStatus; gcmesh_last_error() explains the last failure.
 */
#ifndef GCMESH_H
#define GCMESH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- geometry and caps: mirrored from the reference ---- */
#define GC_CHUNK_SIZE_H 32          /* Code/World.h:6  CHUNK_SIZE_H */
#define GC_CHUNK_SIZE_V 64          /* Code/World.h:7  CHUNK_SIZE_V */
#define GC_CHUNK_SIZE_3 65536       /* Code/World.h:16 CHUNK_SIZE_3 */
#define GC_CHUNK_SIZE_2 1024        /* Code/World.h:17 CHUNK_SIZE_2 */
#define GC_WORLD_CHUNK_HEIGHT 4     /* Code/World.h:21 */
#define GC_WORLD_BLOCK_HEIGHT 256   /* Code/World.h:22 */
#define GC_MAX_VERTICES 40000       /* Code/Mesh.h:5 */
#define GC_MAX_INDICES 60000        /* Code/Mesh.h:6 */
#define GC_MAX_QUADS 10000          /* MAX_VERTICES / 4 == MAX_INDICES / 6 */
#define GC_VERTEX_BYTES 12          /* sizeof(VertexInfo), Code/Mesh.h:37-44 */
#define GC_MAX_BLOCK_TYPES 256

/* BlockMeshType, Code/Mesh.h:8-16 */
enum { GC_MESH_OPAQUE = 0, GC_MESH_CUTOUT = 1, GC_MESH_TRANSPARENT = 2, GC_MESH_FLUID = 3, GC_MESH_MAGMA = 4, GC_MESH_TYPE_COUNT = 5 };
/* CullValue, Code/Block.h:40-47 */
enum { GC_CULL_OPAQUE = 0, GC_CULL_CUTOUT = 1, GC_CULL_TRANSPARENT = 2, GC_CULL_FLUID = 3, GC_CULL_INVISIBLE = 4 };

typedef enum GcStatus
{
    GC_OK = 0,
    GC_ERR_ARG = -1,        /* null pointer, out-of-range coordinate, edge chunk, bad table ... */
    GC_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed (including: no usable GPU) */
    GC_ERR_CAPACITY = -3,   /* batch holds more chunks than it was created for */
    GC_ERR_STATE = -4       /* results requested before gcmesh_wait / submit */
} GcStatus;

/* The fields of the reference's BlockData that the mesher reads (Code/Block.h:85-100; accessors Block.cpp:5-68). */
typedef struct GcBlockInfo
{
    uint8_t textures[6]; /* texture-array layer per face: top, bottom, front, back, right, left (Block.h:49-57) */
    uint8_t mesh_type;   /* GC_MESH_*: which index stream the block's quads go to (Block.cpp:25-28) */
    uint8_t cull;        /* GC_CULL_* 0..4: a face is drawn iff cull(block) < cull(neighbour) (WorldRender.cpp:387-391) */
    uint8_t alpha;       /* vertex alpha (Block.cpp:50-53) */
    uint8_t invisible;   /* never emits quads (Block.cpp:20-23) */
    uint8_t opaque;      /* lightStep == INT_MAX (Block.cpp:65-68): darkens the 3-sample corner average */
    uint8_t reserved;
} GcBlockInfo;

/* Chunk position in the window: lcPos of the reference's Chunk (Code/World.h:71). */
typedef struct GcChunkCoord { int32_t cx, cy, cz; } GcChunkCoord;

/* Per-chunk result flags. */
#define GC_CHUNK_OVERFLOW 1u   /* more than GC_MAX_QUADS quads: the reference asserts (WorldRender.cpp:558); nothing is emitted */
#define GC_CHUNK_NO_SPACE 2u   /* the batch arena was too small for this chunk; nothing is emitted */
#define GC_CHUNK_HALO_MISSING 4u /* gcmesh_submit_exchange: the neighbour's boundary plane never arrived; nothing is emitted */

/*
 * Per-chunk result: the counters of the reference's MeshData / MeshIndexData (Code/Mesh.h:31-52) plus where the
 * chunk's streams live in the batch arenas.  Vertices of chunk i: vertex arena bytes
 * [quad_base*48, quad_base*48 + vert_count*12).  Index stream t: index arena uint16 elements
 * [quad_base*6 + index_offset[t], ... + index_count[t]).
 */
typedef struct GcChunkOut
{
    uint32_t quad_base;
    uint32_t vert_count;                         /* == chunk->totalVertices (WorldRender.cpp:559) unless flags != 0 */
    uint32_t index_count[GC_MESH_TYPE_COUNT];    /* MeshIndexData::count per mesh type */
    uint32_t index_offset[GC_MESH_TYPE_COUNT];   /* uint16 elements from the chunk's index segment start */
    uint32_t quads;                              /* quads the chunk has (also when flags != 0) */
    uint32_t flags;
    uint32_t reserved[2];
} GcChunkOut;

/* The reference's MeshData (Code/Mesh.h:46-52) flattened: what BuildChunkAsync leaves in chunk->meshData. */
typedef struct GcMeshData
{
    uint8_t vertices[GC_MAX_VERTICES * GC_VERTEX_BYTES];
    uint16_t indices[GC_MESH_TYPE_COUNT][GC_MAX_INDICES];
    int32_t index_count[GC_MESH_TYPE_COUNT];
    int32_t vert_count;
} GcMeshData;

typedef struct GcContext GcContext;
typedef struct GcWorld GcWorld;
typedef struct GcBatch GcBatch;

/* ---- context ---- */
const char* gcmesh_version(void);
const char* gcmesh_last_error(void);
/* Binds to CUDA device `device`; work runs on `stream` (a cudaStream_t passed as void*; NULL = a private stream). */
int gcmesh_create(int device, void* stream, GcContext** out);
int gcmesh_destroy(GcContext* ctx);
int gcmesh_set_stream(GcContext* ctx, void* stream);
/* World::blockData (Code/World.h:166) as the mesher needs it.  n <= GC_MAX_BLOCK_TYPES; id 0 must be AIR
 * (BuildChunkAsync skips it by id, WorldRender.cpp:19); kill_zone_id is what a cell below the world reads as
 * (GetBlockSafe, World.cpp:422-428).  A context starts with the reference's 24-entry table (Block.cpp:139-276). */
int gcmesh_set_block_table(GcContext* ctx, const GcBlockInfo* table, int n, int kill_zone_id);
int gcmesh_default_block_table(GcBlockInfo* table, int* n, int* kill_zone_id);
int gcmesh_synchronize(GcContext* ctx);

/* ---- device world: the window of columns, World::groups (Code/World.h:131, World.cpp:153-160) ---- */
/* size_x * size_z columns of 4 chunk bricks + a 32x32 surface map each, zero-initialised (all AIR, no light).
 * slot(cx,cy,cz) = (cz*size_x + cx)*4 + cy; brick layout as the reference: x + 32*(y + 64*z) (World.cpp:278-281). */
int gcmesh_world_create(GcContext* ctx, int size_x, int size_z, GcWorld** out);
int gcmesh_world_destroy(GcWorld* world);
/* Host -> device copy of one chunk's arrays (Chunk::blocks / sunlight / blockLight, World.h:74-78); NULL skips an array.
 * Asynchronous on the context stream; the host arrays must stay valid until gcmesh_synchronize / gcmesh_wait. */
int gcmesh_world_upload_chunk(GcWorld* world, int cx, int cy, int cz, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light);
/* ChunkGroup::surface (World.h:94). */
int gcmesh_world_upload_surface(GcWorld* world, int cx, int cz, const uint8_t* surface);
/* Whole window at once from brick-major host arrays (nslots*65536 elements each, surface ncols*1024). */
int gcmesh_world_upload_all(GcWorld* world, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light, const uint8_t* surface);
/* Rows [cz0, cz1) of columns only (a z-slab of the window), same host layout as upload_all (pointers to the full arrays). */
int gcmesh_world_upload_rows(GcWorld* world, int cz0, int cz1, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light, const uint8_t* surface);
/* Same meshes as after gcmesh_world_upload_all, moving only what is not zero (and the same device contents, with two
 * exceptions that no mesh can see: a sunlight brick lying wholly at or above its column's `surface` — nothing ever reads
 * stored sunlight there, Lighting.cpp:69-79 — and the sunlight / blockLight of a brick whose 3 x 3 x 3 brick
 * neighbourhood holds no block at all — light is only read within one voxel of a block that emits a quad,
 * WorldRender.cpp:329-383 — are left cleared: neither scanned nor transferred).  Each brick array of each chunk
 * is classified zero / non-zero — by a multi-threaded host scan (nonzero_hint == NULL; `threads` <= 0 = all cores), or by
 * the caller's per-slot hint byte (bit 0 blocks, bit 1 sunlight, bit 2 blockLight non-zero; the caller vouches for it).
 * Non-zero arrays are pulled by a GPU copy kernel straight from the host arrays when those are pinned (cudaHostAlloc /
 * cudaHostRegister), else by one cudaMemcpyAsync each; arrays that were non-zero on the device and are zero now are cleared. */
int gcmesh_world_upload_sparse(GcWorld* world, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light,
                               const uint8_t* surface, const uint8_t* nonzero_hint, int threads);
/* Options of a window.  GC_WORLD_SURFACE_BOUNDS_BLOCKS (default 0): the caller vouches that the `surface` maps handed to
 * gcmesh_world_upload_sparse / gcmesh_mesh_host are ComputeSurface of the block ids handed with them (Lighting.cpp:5-26 —
 * what ChunkGroup::surface holds for every pre-processed column; a stale entry is only ever too high).  The host scan then
 * takes a brick lying wholly at or above its column's highest entry for AIR without reading it (layer y = 255, which
 * ComputeSurface does not look at, is still read).
 * GC_WORLD_SURFACE_BOUNDS_MESH (default 0): the same promise for the window on the device — its surface map is ComputeSurface of
 * its block ids, as it is after gcmesh_world_compute_surface, gcmesh_world_generate, gcmesh_world_set_blocks, or uploads of a
 * pre-processed window (in the reference a chunk is only ever meshed after PreprocessGroup, WorldRender.cpp:240-262).  The mesh
 * kernel then takes a brick lying wholly at or above its column's highest surface entry for AIR — an empty mesh — from the
 * 1 KB surface map, without reading the brick's 128 KB of block ids (the top brick's layer y = 255 is still looked at).
 * With the option off (default) the mesher is a function of the block ids alone, whatever the surface map says.
 * GC_WORLD_EDIT_REPLAY (default 0): gcmesh_world_set_blocks relights by replaying the reference's own RecomputeLight
 * (Lighting.cpp:283-449) node for node — the same FIFO queues, neighbour order and comparisons, on one device thread — instead
 * of refilling the box around the edit: block ids, the surface map (including the entry the reference leaves stale when a column
 * empties), both light arrays and the set of chunks flagged for rebuild come out byte for byte as the reference's SetBlock leaves
 * them (also where its removal queues stop short of a from-scratch fill).  Slower per edit than the default; see below.
 * GC_WORLD_COPY_RUN (default 0 = 4): sparse uploads move the same brick array of at least this many consecutive columns
 * with one strided copy-engine copy; shorter runs are pulled by a GPU kernel out of pinned host memory (a huge value: always). */
enum { GC_WORLD_SURFACE_BOUNDS_BLOCKS = 1, GC_WORLD_COPY_RUN = 2, GC_WORLD_SURFACE_BOUNDS_MESH = 3, GC_WORLD_EDIT_REPLAY = 4 };
int gcmesh_world_set_option(GcWorld* world, int option, int value);
/* World-space column (ChunkGroup::pos, World.h) of window column (0, 0).  Only region records depend on it (gcmesh_world_encode_rle). */
int gcmesh_world_set_origin(GcWorld* world, int origin_cx, int origin_cz);
/* The reference asserts block < BLOCK_COUNT wherever a block is stored (World.cpp:294).  The mesh kernel indexes its class
 * table with the low byte of an id; ids from the table size up to 255 read as AIR, larger ones alias.  This scans the
 * device window's block ids once (HBM-bound) and reports how many are >= the size of the block table in use and, if any,
 * the chunk and voxel index (BlockIndex order) of the first.  Synchronous on the context stream. */
int gcmesh_world_validate_blocks(GcWorld* world, unsigned long long* n_invalid, GcChunkCoord* first_chunk, int* first_voxel);
/* Kernels launched by the last gcmesh_world_upload_sparse call. */
int gcmesh_world_upload_launches(const GcWorld* world);
/* Host -> device bytes the last gcmesh_world_upload_sparse / gcmesh_mesh_host call moved (brick arrays, work lists, surface maps, coordinates). */
unsigned long long gcmesh_world_upload_bytes(const GcWorld* world);
/* Raw device pointers of the four arenas, for device-resident producers (and for NCCL halo exchange). */
int gcmesh_world_device_ptrs(GcWorld* world, void** blocks, void** sunlight, void** block_light, void** surface);
/* Slab halo (multi-GPU): the z-boundary plane of one row of columns — for every column cx and chunk cy the
 * 32x64 voxels at z = `z` of blocks (4096 B) + sunlight (2048 B) + blockLight (2048 B), then the surface row
 * (32 B) per column.  pack gathers row `cz`, plane `z` into `device_buf`; unpack scatters it into row `cz`. */
size_t gcmesh_world_halo_bytes(const GcWorld* world);
int gcmesh_world_pack_halo(GcWorld* world, int cz, int z, void* device_buf);
int gcmesh_world_unpack_halo(GcWorld* world, int cz, int z, const void* device_buf);
/* The same exchange without a buffer or a collective: each rank stores its boundary planes straight into its
 * neighbours' window arrays over NVLink peer memory (one process per GPU: CUDA IPC handles; one process, several
 * worlds: direct).  Neighbouring windows have one size.  side 0 = the rank below (z - 1), 1 = the rank above.
 *   gcmesh_world_halo_export        GC_HALO_HANDLE_BYTES of handles for this window (send them to both neighbours)
 *   gcmesh_world_halo_connect_ipc   open a neighbour's handles; peer_row = the neighbour's halo row that receives this
 *                                   rank's plane (its first owned row - 1 for the rank above, its last owned row + 1 below)
 *   gcmesh_world_halo_connect_local the neighbour is a world of this process (any device with peer access)
 *   gcmesh_world_halo_push          one exchange, one kernel enqueued on the context stream: tell the neighbours this rank's earlier
 *                                   work is done and wait for the same from them (their halo rows may then be
 *                                   overwritten), copy the z = 0 plane of row send_row_down to the rank below and the
 *                                   z = 31 plane of row send_row_up to the rank above, wait for the neighbours' planes.
 *                                   Work enqueued afterwards (gcmesh_submit) sees the new halo.  Every rank of a chain
 *                                   calls it the same number of times.
 *                                   A neighbour that does not enter the exchange within the wait bound is left alone (nothing
 *                                   is stored into its window, no arrival is published to it) and the failure is recorded in
 *                                   both ranks' flag blocks; gcmesh_submit / gcmesh_wait do not look there — poll
 *                                   gcmesh_world_halo_status, or use gcmesh_submit_exchange, whose gcmesh_wait reports it.
 *   gcmesh_world_halo_status        synchronises; timed_out != 0: a neighbour did not show up within 10 s
 *                                   (GCMESH_HALO_WAIT_MS, read at the first exchange of the process) — an error, not a hang;
 *                                   the value is (exchange serial << 8) | (1 + flag word waited for) of the first such wait,
 *                                   here or as reported by a neighbour that waited for this rank */
#define GC_HALO_HANDLE_BYTES 320
int gcmesh_world_halo_export(GcWorld* world, void* handles);
int gcmesh_world_halo_connect_ipc(GcWorld* world, int side, const void* handles, int peer_row);
int gcmesh_world_halo_connect_local(GcWorld* world, int side, GcWorld* peer, int peer_row);
int gcmesh_world_halo_push(GcWorld* world, int send_row_down, int send_row_up);
int gcmesh_world_halo_status(GcWorld* world, int* timed_out, uint32_t* serial);

/* ---- column pre-processing on the device: the producers of the mesher's light inputs ----
 * Replaces, for a window whose block ids are on the device, ComputeSurface (Code/Lighting.cpp:5-26) and SetLightNodes /
 * ScatterSunlight / ScatterBlockLight (Lighting.cpp:28-281) as PreprocessGroup runs them over every non-edge column
 * (WorldRender.cpp:240-262).  The flood fill is computed as the least fixpoint of the reference's push rule: sunlight is
 * identical to the reference's; block light is identical except where the reference's result depends on the order in
 * which columns are processed (an emitter re-assigned below light it had already received, Lighting.cpp:240) — there
 * the brighter, order-independent value is kept. */
/* lightStep per block id (1..15, 255 = opaque i.e. INT_MAX, Block.cpp:65-68) and lightEmitted (0..15), Block.h:94-95. */
int gcmesh_default_light_rules(uint8_t* step256, uint8_t* emitted256);
int gcmesh_set_light_rules(GcContext* ctx, const uint8_t* step256, const uint8_t* emitted256);
/* surface[column][z*32 + x] = highest non-AIR y in 0..254, plus one (0 for an empty column cell).  Asynchronous. */
int gcmesh_world_compute_surface(GcWorld* world);
/* Clears the sunlight / blockLight arenas and fills them from the block ids and the surface map.  Asynchronous: a fixed
 * sequence on the context stream — clears, one scan kernel, one persistent relaxation kernel (cooperative launch; candidates are
 * bit maps, the active set a byte per brick and sweep: nothing is sized from a device count) — and no synchronisation. */
int gcmesh_world_light(GcWorld* world);
/* Of the last gcmesh_world_light: {sunlight candidate cells, brick visits of the relaxation (both fills, all sweeps), sweeps
 * that found something to visit, seeded emitters}, kernels launched.  Synchronises the context stream to read them. */
int gcmesh_world_light_stats(GcWorld* world, uint32_t* stats4, int* launches);
/* Device -> host copy of the window's arrays (NULL skips one); synchronous. */
int gcmesh_world_download(GcWorld* world, uint16_t* blocks, uint8_t* sunlight, uint8_t* block_light, uint8_t* surface);

/* ---- region records: the reference's saved-chunk format (Code/WorldIO.cpp:167-307) ----
 * A chunk's block ids as SaveGroup stores them in region->chunks[]: [offset][items][(count, block) ...], all uint16, runs in
 * voxel index order; a chunk of one block is the single pair (0, block) (65 536 wrapped).  Saved chunks cross PCIe as
 * these records (a few KB) and are expanded / produced on the device. */
typedef struct GcRleChunk
{
    int32_t cx, cy, cz;     /* chunk position in the window */
    uint32_t first_word;    /* where its record starts in the uint16 stream (even) */
    uint32_t words;         /* record length in uint16, the two header words included */
} GcRleChunk;
/* LoadGroupFromDisk's decode (WorldIO.cpp:196-216) for `n` records of `data` into their bricks.  Asynchronous on the
 * context stream (`data` must stay valid until gcmesh_synchronize).  A record whose runs do not add up to exactly one chunk
 * is refused: its brick is cleared and counted (gcmesh_world_rle_refused); the reference would write past the array. */
int gcmesh_world_upload_rle(GcWorld* world, const GcRleChunk* chunks, int n, const uint16_t* data, size_t data_words);
int gcmesh_world_rle_refused(GcWorld* world, int* refused);
/* SaveGroup's encode (WorldIO.cpp:272-303) of `n` bricks of the device world: records are appended to `records` (host,
 * `capacity_words` uint16), out[i] says where chunk i's record lies.  Word 0 of a record is RegionIndex(ChunkP & REGION_MASK)
 * of the chunk's WORLD position (WorldIO.cpp:259-260; LoadRegionFile files the record under it, :48-62): window position +
 * the window's origin as set by gcmesh_world_set_origin or the last gcmesh_world_generate (default (0, 0)).  Synchronous. */
int gcmesh_world_encode_rle(GcWorld* world, const GcChunkCoord* coords, int n, uint16_t* records, size_t capacity_words, GcRleChunk* out);

/* ---- terrain generation on the device: columns are born in HBM ----
 * The reference's nine generators (Environment.cpp:83-143) — GenerateFlatTerrain Code/Generation.cpp:772-836,
 * GenerateGridTerrain :371-410, GenerateVoidTerrain :838-852, GenerateDungeon Code/Dungeon.cpp:17-31: no noise involved,
 * identical to the reference's own output; GenerateForestTerrain :83-226, GenerateIslandsTerrain :228-369,
 * GenerateSnowTerrain :412-556, GenerateDesertTerrain :558-645, GenerateVolcanicTerrain :647-770 with CreateTree :50-64,
 * IsWithinIsland :66-81 — fill the block ids of every column
 * of the window and its surface map (as ComputeSurface would, Lighting.cpp:5-26); sunlight and blockLight are cleared
 * (follow with gcmesh_world_light) and every chunk counts as never built (gcmesh_world_set_blocks refuses it until it is meshed).  origin = world-space column of window column (0, 0); radius / falloff = WorldProperties::radius
 * and World::falloffRadius in blocks (an unbounded world: INT_MAX and INT_MAX - 64).  For the five noise terrains everything the reference's
 * code does is reproduced (its own Generation.cpp, run over this project's noise, gives the same blocks — tested); its noise
 * library (FastNoiseSIMD, not vendored) and its thread-timing dependent rand() are replaced by a seeded simplex noise and a
 * per-column PRNG — the same ones as the host input producer gamecraft_b200/worldgen, whose worlds this call reproduces block
 * for block.  Asynchronous on the context stream. */
enum { GC_GEN_FLAT = 0, GC_GEN_FOREST = 1, GC_GEN_ISLANDS = 2, GC_GEN_GRID = 3, GC_GEN_VOID = 4,
       GC_GEN_SNOW = 5, GC_GEN_DESERT = 6, GC_GEN_VOLCANIC = 7, GC_GEN_DUNGEON = 8 };
int gcmesh_world_generate(GcWorld* world, int kind, uint32_t seed, int origin_cx, int origin_cz, int radius, int falloff);

/* ---- block edits on a device-resident window: the interactive caller of the mesher ----
 * SetBlock(world, lwX, lwY, lwZ, block) (Code/World.cpp:548-582) for a window whose block ids, surface map and light
 * arrays are on the device (gcmesh_world_compute_surface + gcmesh_world_light, or FULL uploads of both light arrays: after a
 * sparse upload — gcmesh_world_upload_sparse, gcmesh_mesh_host — the light arrays no mesh depends on were left out, the
 * window does not hold the light fixpoint the relight builds on, and the call fails with GC_ERR_STATE until
 * gcmesh_world_light or gcmesh_world_upload_all with both light arrays has run):
 *   - the reference's decisions: y outside the world is ignored; a chunk that was never meshed refuses the edit
 *     (chunk->state < CHUNK_NEEDS_FILL; here: no gcmesh_submit has covered it); an unchanged block does nothing;
 *     ChunkOverflowed (WorldRender.cpp:412-439) — the new block's own drawable faces on top of the chunk's vertex count
 *     of its last build — leaves AIR in the cell and neither relights nor flags, exactly as the reference does;
 *   - RecomputeLight (Lighting.cpp:283-449): the column's surface entry is recomputed (a column left without any block
 *     reads 0, the value a freshly generated empty column has) and both light arrays are refilled inside the full-height
 *     box of radius 16 around the edited column from the unchanged cells around it.  The result is what
 *     gcmesh_world_light computes from scratch on the edited blocks — the reference's own fill of that world when it is
 *     next loaded; its incremental queues can stop a few cells short of that;
 *   - FlagChunkForUpdate (World.cpp:507-546): the chunks around the edited cell, plus every chunk that has a cell whose
 *     block, blockLight or effective sunlight changed inside its one-voxel halo.  Chunks outside the list mesh to the
 *     bytes they meshed to before.
 * With GC_WORLD_EDIT_REPLAY the last two items are the reference's to the letter instead: its incremental queues, and
 * every built chunk a queue touched flagged (a superset of the chunks whose mesh changes).
 * Edits are applied in order; the caller's player-overlap test (World.cpp:552) stays with the caller. */
typedef struct GcBlockEdit
{
    int32_t wx, wy, wz;     /* window voxel coordinates (LWorldP): 0 <= wx < 32*size_x, 0 <= wz < 32*size_z; edge columns are rejected */
    uint16_t block;
    uint16_t reserved;
} GcBlockEdit;
enum
{
    GC_EDIT_APPLIED = 0,        /* block stored, light refilled, chunks flagged */
    GC_EDIT_UNCHANGED = 1,      /* the cell already held this block */
    GC_EDIT_IGNORED = 2,        /* wy outside 0..255 (World.cpp:550) */
    GC_EDIT_NOT_BUILT = 3,      /* the chunk has never been meshed (World.cpp:562-565) */
    GC_EDIT_OVERFLOW = 4        /* ChunkOverflowed: the cell now holds AIR (World.cpp:572-573) */
};
/* status: n entries (GC_EDIT_*), may be NULL.  rebuild: up to rebuild_capacity coordinates, sorted by (cz, cx, cy);
 * *n_rebuild = how many chunks are flagged (GC_ERR_CAPACITY when more than fit; the first rebuild_capacity are stored).
 * Runs on the context stream behind earlier submissions; synchronous. */
int gcmesh_world_set_blocks(GcWorld* world, const GcBlockEdit* edits, int n, int32_t* status,
                            GcChunkCoord* rebuild, int rebuild_capacity, int* n_rebuild);
/* chunk->totalVertices of every brick ((cz * size_x + cx) * 4 + cy; -1 = never meshed) as the last submissions left it. */
int gcmesh_world_vertex_counts(GcWorld* world, int32_t* counts);

/* ---- batches: one GPU submission = the reference's vector<Chunk*> job (WorldRender.cpp:29-37) ---- */
/* Arenas for up to max_chunks chunks and max_quads quads in total (48 B of vertices + 12 B of indices per quad). */
int gcmesh_batch_create(GcContext* ctx, int max_chunks, uint64_t max_quads, GcBatch** out);
/* The same with flags.  GC_BATCH_UNCAPPED: the "stress" form of SURVEY 8(d) config 4 — NOT a reference behaviour (the reference
 * asserts at 40 000 vertices, WorldRender.cpp:558): chunks of any quad count are meshed (a full 3-D checkerboard has
 * 196 608), never flagged GC_CHUNK_OVERFLOW, and every index is a uint32 (SetIndices' pattern o + 2, o + 1, o, o + 3, o + 2, o
 * without the uint16 wrap): the index arena holds 24 bytes per quad, index_count / index_offset count uint32 elements.
 * Vertex bytes are what the capped form writes.  gcmesh_batch_fill_meshdata and gcmesh_submit_exchange refuse such a batch. */
#define GC_BATCH_UNCAPPED 1u
int gcmesh_batch_create_ex(GcContext* ctx, int max_chunks, uint64_t max_quads, uint32_t flags, GcBatch** out);
int gcmesh_batch_destroy(GcBatch* batch);
/* Mesh `n` chunks of `world`.  Chunks on the window edge are rejected (the reference never meshes them,
 * WorldRender.cpp:235).  Asynchronous; inputs must not change until gcmesh_wait (the reference's gating rule,
 * WorldRender.cpp:265-277). `coords` is copied. */
int gcmesh_submit(GcWorld* world, GcBatch* batch, const GcChunkCoord* coords, int n);
/* Slab halo exchange + mesh build as ONE kernel (multi-GPU; replaces gcmesh_world_halo_push followed by gcmesh_submit):
 * the kernel's first work items store this rank's boundary planes — the z = 0 plane of row send_row_down to the rank below,
 * the z = 31 plane of row send_row_up to the rank above — straight into the neighbours' halo rows over NVLink peer memory
 * while the other CTAs already mesh; the chunks of rows send_row_down / send_row_up, the only ones that read a neighbour's
 * plane, are claimed last and wait for that plane's "arrived" flag.  Nothing separates the exchange from the meshing: no
 * second launch, no flag round trip ahead of the first chunk.  Neighbours are connected as for gcmesh_world_halo_push;
 * every rank of a chain calls it the same number of times (n >= 1).  A plane that does not arrive within the wait bound
 * (GCMESH_HALO_WAIT_MS) flags the chunks that needed it GC_CHUNK_HALO_MISSING and makes gcmesh_wait fail. */
int gcmesh_submit_exchange(GcWorld* world, GcBatch* batch, const GcChunkCoord* coords, int n, int send_row_down, int send_row_up);
int gcmesh_wait(GcBatch* batch);
/* The same without blocking: *done = 1 once the submission has finished (results may then be read as after gcmesh_wait), else 0.
 * What the reference's main thread does when it polls its job callbacks once per frame (Async.cpp:48-60). */
int gcmesh_query(GcBatch* batch, int* done);
/* After gcmesh_wait: per-chunk descriptors (host memory owned by the batch, n entries, submit order). */
int gcmesh_batch_results(GcBatch* batch, const GcChunkOut** out, int* n, uint64_t* total_quads);
/* After gcmesh_wait: copy the used arena prefix to host memory (total_quads*48 and total_quads*12 bytes). */
int gcmesh_batch_download(GcBatch* batch, void* vertices, void* indices);
/* After gcmesh_wait: chunk i as the reference's MeshData. */
int gcmesh_batch_fill_meshdata(GcBatch* batch, int i, GcMeshData* out);
int gcmesh_batch_device_ptrs(GcBatch* batch, void** vertices, void** indices, void** chunk_out);
/* FillMeshData (Code/Mesh.cpp:71-130) for buffers that live on the device: where the reference uploads a chunk's
 * MeshData with glBufferData — one GL_ARRAY_BUFFER of vert_count * 12 bytes (attributes: position 3 x u8 at 0, texture
 * coordinates 3 x u8 at 3, colour 4 x u8 normalised at 6, alpha u8 normalised at 10; Mesh.h:37-44) and one
 * GL_ELEMENT_ARRAY_BUFFER of index_count[t] uint16 per mesh type with indices — an integration maps those buffers into
 * CUDA (cudaGraphicsGLRegisterBuffer) and passes the mapped pointers here: the streams go from the batch arenas to the
 * graphics buffers without crossing PCIe.  After gcmesh_wait (sizes come from gcmesh_batch_results); NULL skips a
 * stream; asynchronous on the context stream, `targets` is copied. */
typedef struct GcDeviceMesh
{
    int32_t chunk;                              /* index into the batch, submit order */
    int32_t reserved;
    void* vertices;                             /* device memory, vert_count * 12 bytes */
    void* indices[GC_MESH_TYPE_COUNT];          /* device memory, index_count[t] * 2 bytes each */
} GcDeviceMesh;
int gcmesh_batch_fill_device(GcBatch* batch, const GcDeviceMesh* targets, int n);
/* Host arrays in, host meshes out, in ONE call: what an integration does where the reference queues
 * RebuildChunksAsync over a vector<Chunk*> (WorldRender.cpp:80-91).  Equivalent to gcmesh_world_upload_sparse +
 * gcmesh_submit + gcmesh_wait + gcmesh_batch_download, but pipelined in strips of rows on three streams (host zero-scan,
 * PCIe pull, meshing and read-back overlap).  On return (synchronous) host_vertices / host_indices hold the arena prefix
 * (min(total quads, host_capacity_quads) quads) and gcmesh_batch_results gives the descriptors in the caller's order.
 * Host arrays: brick-major as for gcmesh_world_upload_all; pinned memory lets the GPU pull them itself. */
int gcmesh_mesh_host(GcWorld* world, GcBatch* batch, const uint16_t* blocks, const uint8_t* sunlight, const uint8_t* block_light,
                     const uint8_t* surface, const uint8_t* nonzero_hint, const GcChunkCoord* coords, int n,
                     void* host_vertices, void* host_indices, uint64_t host_capacity_quads, int threads);
/* Milliseconds the last submission's kernels took on the device (CUDA events on the context stream).  The two timing events
 * around the kernel are recorded by default; gcmesh_batch_set_timing(batch, 0) leaves them out (back-to-back submissions
 * then have nothing but the launch between their kernels) and gcmesh_batch_elapsed_ms fails with GC_ERR_STATE. */
int gcmesh_batch_elapsed_ms(GcBatch* batch, float* ms);
int gcmesh_batch_set_timing(GcBatch* batch, int on);
/* Number of kernel launches the last submission made. */
int gcmesh_batch_launches(GcBatch* batch);

/* Development aid: 512 in-kernel cycle counters of the last submission (needs GCMESH_PROFILE=1 in the environment). */
int gcmesh_batch_profile(GcBatch* batch, unsigned long long* out512);
/* Registers per thread, dynamic shared memory and block size of the mesh kernel (for reports). */
int gcmesh_kernel_info(int* regs, int* smem_bytes, int* threads);

#ifdef __cplusplus
}
#endif
#endif /* GCMESH_H */

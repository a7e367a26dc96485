/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for the chunk mesher: a from-scratch plain-C restatement of the
 * algorithm in the reference's Code/WorldRender.cpp / Code/Mesh.cpp / Code/World.cpp / Code/Lighting.cpp
 * (file:line cited at each function in gc_mesh_cpu.c).  Only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs may load it; the product library (libgcmesh.so) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py byte-compares this oracle with the reference's
 * own mesher compiled verbatim (oracle/_ref/libgcref.so, built by oracle/Makefile from /root/reference), and
 * tests/golden/ holds vectors minted from that gold build (tests/golden/make_golden.py).
 */
#ifndef GC_ORACLE_H
#define GC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCOR_CHUNK_H 32          /* Code/World.h:6  CHUNK_SIZE_H */
#define GCOR_CHUNK_V 64          /* Code/World.h:7  CHUNK_SIZE_V */
#define GCOR_CHUNK_VOXELS 65536  /* Code/World.h:16 CHUNK_SIZE_3 */
#define GCOR_WORLD_CHUNKS_Y 4    /* Code/World.h:21 WORLD_CHUNK_HEIGHT */
#define GCOR_WORLD_HEIGHT 256    /* Code/World.h:22 WORLD_BLOCK_HEIGHT */
#define GCOR_MAX_VERTICES 40000  /* Code/Mesh.h:5 */
#define GCOR_MAX_INDICES 60000   /* Code/Mesh.h:6 */
#define GCOR_MESH_TYPES 5        /* Code/Mesh.h:8-16 */
#define GCOR_MAX_BLOCKS 256

/* The fields of the reference's BlockData the mesher reads (Code/Block.h:85-100, Code/Block.cpp:5-68). */
typedef struct GcOrBlockInfo
{
    uint8_t textures[6]; /* top, bottom, front, back, right, left (Code/Block.h:49-57) */
    uint8_t mesh_type;   /* BlockMeshType */
    uint8_t cull;        /* CullValue 0..4 */
    uint8_t alpha;
    uint8_t invisible;
    uint8_t opaque;      /* lightStep == INT_MAX (Code/Block.cpp:65-68) */
    uint8_t reserved;
} GcOrBlockInfo;

/*
 * A window of columns like World::groups (Code/World.h:131, World.cpp:153-160), stored brick-major:
 *   slot(cx,cy,cz) = (cz*size_x + cx)*4 + cy
 *   blocks[slot*65536 + x + 32*(y + 64*z)]   (Code/World.cpp:278-281), same index for sun / light
 *   surface[(cz*size_x + cx)*1024 + z*32 + x] (Code/World.h:94, Lighting.cpp:7)
 * The arrays are borrowed, never copied.
 */
typedef struct GcOrWorld
{
    int size_x, size_z;
    const uint16_t* blocks;
    const uint8_t* sun;
    const uint8_t* light;
    const uint8_t* surface;
    GcOrBlockInfo table[GCOR_MAX_BLOCKS];
    int table_count;
    int kill_zone_id; /* Code/Block.h:32 BLOCK_KILL_ZONE */
} GcOrWorld;

/* Result of one BuildChunkAsync: the reference's MeshData (Code/Mesh.h:46-52) with inline index arrays. */
typedef struct GcOrMesh
{
    uint8_t vertices[GCOR_MAX_VERTICES * 12];
    uint16_t indices[GCOR_MESH_TYPES][GCOR_MAX_INDICES];
    int32_t vert_count;
    int32_t index_count[GCOR_MESH_TYPES];
    int32_t index_allocated[GCOR_MESH_TYPES]; /* WorldRender.cpp:454-459: array exists once a visible block of that type was visited */
    int32_t overflow;                         /* would have exceeded MAX_VERTICES (WorldRender.cpp:558) — output truncated */
} GcOrMesh;

void gcor_default_block_table(GcOrBlockInfo* table, int* count, int* kill_zone_id);
void gcor_world_init(GcOrWorld* w, int size_x, int size_z, const uint16_t* blocks, const uint8_t* sun,
                     const uint8_t* light, const uint8_t* surface);
/* BuildChunkAsync (WorldRender.cpp:6-27).  Returns totalVertices. */
int gcor_build_chunk(const GcOrWorld* w, int cx, int cy, int cz, GcOrMesh* out);
/* BuildChunkAsync with the fixed-size MeshData lifted (the "stress" variant, no reference behaviour: no cap, 32-bit indices). */
int64_t gcor_build_chunk_uncapped(const GcOrWorld* w, int cx, int cy, int cz, uint8_t* vertices, uint32_t* const* indices, int64_t max_quads,
                                  int64_t* index_count);
/* ChunkOverflowed (WorldRender.cpp:412-439). */
int gcor_chunk_overflowed(const GcOrWorld* w, int cx, int cy, int cz, int x, int y, int z, int total_vertices);
/* One-chunk-per-job pool (Async.cpp:5-35 restated as an atomic counter). Returns seconds. */
double gcor_bench_build(const GcOrWorld* w, const int32_t* coords, int n, int threads, int64_t* quads_out);
/* ---- column pre-processing (gc_light_cpu.c): the producers of the mesher's light inputs ---- */
/* ComputeSurface (Lighting.cpp:5-26) over the whole window. */
void gcor_compute_surface(int size_x, int size_z, const uint16_t* blocks, uint8_t* surface);
/* SetLightNodes / ScatterSunlight / ScatterBlockLight (Lighting.cpp:28-281) over all non-edge columns, as the least
 * fixpoint of the reference's push rule (block light: canonical max form, see gc_light_cpu.c).  step[id] = lightStep,
 * 255 = opaque; emitted[id] = lightEmitted.  sun / light (whole-window arenas) are overwritten. */
void gcor_light_fill(int size_x, int size_z, const uint16_t* blocks, const uint8_t* surface, const uint8_t* step, const uint8_t* emitted,
                     uint8_t* sun, uint8_t* light);
void gcor_default_light_rules(uint8_t* step, uint8_t* emitted);
/* ---- region chunk format (gc_rle_cpu.c): SaveGroup / LoadGroupFromDisk's run-length coding of a chunk's block ids ---- */
int gcor_rle_encode_chunk(const uint16_t* blocks, uint16_t offset, uint16_t* out /* 2 + 131072 words */);
int gcor_rle_decode_chunk(const uint16_t* record, int words, uint16_t* blocks);
/* ---- block edits (gc_edit_cpu.c): SetBlock's decisions, stored block, surface entry and FlagChunkForUpdate set ---- */
int gcor_set_block(int size_x, int size_z, uint16_t* blocks, uint8_t* surface, const int32_t* vert_table, const uint8_t* cull,
                   int kill_zone, int wx, int wy, int wz, int block, uint8_t* flagged);
/* FNV-1a-64 over bytes (used for golden digests). */
uint64_t gcor_fnv1a64(const void* data, uint64_t n, uint64_t h);

#ifdef __cplusplus
}
#endif
#endif

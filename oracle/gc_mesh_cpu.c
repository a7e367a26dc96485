/*
 * TEST INFRASTRUCTURE ONLY — see gc_oracle.h.  Plain-C restatement of the reference chunk mesher.
 * Written from the behaviour of the reference (file:line cited per function), not from its text:
 * the world is a dense window of bricks and every neighbour lookup goes through one probe function.
 */
#define _GNU_SOURCE
#include "gc_oracle.h"

#include <sched.h>

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- block table: Code/Block.cpp:139-276 with image ids from Code/Assets.h:7-34 ---- */
enum { OR_CULL_OPAQUE = 0, OR_CULL_CUTOUT = 1, OR_CULL_TRANSPARENT = 2, OR_CULL_FLUID = 3, OR_CULL_INVISIBLE = 4 };
enum { OR_MESH_OPAQUE = 0, OR_MESH_CUTOUT = 1, OR_MESH_TRANSPARENT = 2, OR_MESH_FLUID = 3, OR_MESH_MAGMA = 4 };

static void or_solid(GcOrBlockInfo* b, int top, int bottom, int side)
{
    /* CreateDefaultBlock (Block.cpp:128-137): opaque, alpha 255, cull 0, mesh type 0 */
    memset(b, 0, sizeof(*b));
    b->textures[0] = (uint8_t)top;
    b->textures[1] = (uint8_t)bottom;
    b->textures[2] = b->textures[3] = b->textures[4] = b->textures[5] = (uint8_t)side;
    b->alpha = 255;
    b->opaque = 1;
}

void gcor_default_block_table(GcOrBlockInfo* t, int* count, int* kill_zone_id)
{
    memset(t, 0, sizeof(GcOrBlockInfo) * 24);
    or_solid(&t[0], 0, 0, 0);   /* AIR  Block.cpp:141-148 */
    t[0].invisible = 1; t[0].cull = OR_CULL_INVISIBLE; t[0].opaque = 0;
    or_solid(&t[1], 8, 6, 9);   /* GRASS :150-152 */
    or_solid(&t[2], 6, 6, 6);   /* DIRT :154-156 */
    or_solid(&t[3], 48, 48, 48); /* STONE :158-160 */
    or_solid(&t[4], 5, 5, 5);   /* CRATE :177-179 */
    or_solid(&t[5], 44, 44, 44); /* SAND :173-175 */
    or_solid(&t[6], 49, 49, 49); /* STONE_BRICK :181-183 */
    or_solid(&t[7], 42, 42, 42); /* METAL_CRATE :185-187 */
    or_solid(&t[8], 51, 51, 51); /* WATER :162-171 */
    t[8].mesh_type = OR_MESH_FLUID; t[8].cull = OR_CULL_FLUID; t[8].alpha = 127; t[8].opaque = 0;
    or_solid(&t[9], 60, 60, 59); /* WOOD :189-191 */
    or_solid(&t[10], 21, 21, 21); /* LEAVES :193-195 */
    or_solid(&t[11], 3, 3, 3);  /* CLAY :197-199 */
    or_solid(&t[12], 46, 6, 47); /* SNOW :201-203 */
    or_solid(&t[13], 10, 10, 10); /* ICE :205-213 */
    t[13].mesh_type = OR_MESH_TRANSPARENT; t[13].cull = OR_CULL_TRANSPARENT; t[13].alpha = 190; t[13].opaque = 0;
    or_solid(&t[14], 12, 12, 12); /* LANTERN :215-219 (lightStep 1 => not opaque; cull stays 0) */
    t[14].opaque = 0;
    or_solid(&t[15], 50, 50, 50); /* TRAMPOLINE :221-224 */
    or_solid(&t[16], 2, 0, 1);  /* CACTUS :226-230 */
    or_solid(&t[17], 22, 22, 22); /* MAGMA :232-239 */
    t[17].mesh_type = OR_MESH_MAGMA; t[17].opaque = 0;
    or_solid(&t[18], 4, 4, 4);  /* COOLED_MAGMA :241-243 */
    or_solid(&t[19], 43, 43, 43); /* OBSIDIAN :245-247 */
    or_solid(&t[20], 0, 0, 0);  /* KILL_ZONE :249-254 */
    t[20].invisible = 1; t[20].cull = OR_CULL_INVISIBLE; t[20].opaque = 0;
    or_solid(&t[21], 13, 13, 13); /* LAVA :256-266 */
    t[21].mesh_type = OR_MESH_FLUID; t[21].cull = OR_CULL_FLUID; t[21].alpha = 127; t[21].opaque = 0;
    or_solid(&t[22], 7, 7, 7);  /* GLASS :268-273 */
    t[22].mesh_type = OR_MESH_CUTOUT; t[22].cull = OR_CULL_CUTOUT; t[22].opaque = 0;
    or_solid(&t[23], 45, 45, 45); /* SHOCKER :275-277 */
    *count = 24;          /* BLOCK_COUNT, Block.h:36 */
    *kill_zone_id = 20;   /* BLOCK_KILL_ZONE, Block.h:32 */
}

void gcor_world_init(GcOrWorld* w, int size_x, int size_z, const uint16_t* blocks, const uint8_t* sun,
                     const uint8_t* light, const uint8_t* surface)
{
    memset(w, 0, sizeof(*w));
    w->size_x = size_x;
    w->size_z = size_z;
    w->blocks = blocks;
    w->sun = sun;
    w->light = light;
    w->surface = surface;
    gcor_default_block_table(w->table, &w->table_count, &w->kill_zone_id);
}

/* ---- addressing ---- */

/* A probed cell: where == 0 inside the window, -1 below the world, +1 above it. */
typedef struct OrCell
{
    int where;
    int64_t voxel;   /* index into blocks/sun/light */
    int64_t column;  /* index of the cell's own column's surface entry */
    int world_y;
} OrCell;

/*
 * Rebase (World.cpp:380-415) for a position relative to chunk (cx,cy,cz): horizontal exits move to the
 * neighbouring column first (x, then z), then a vertical exit either moves to the chunk above/below or,
 * outside 0..3, yields the null-chunk marker with rY = -1 / +1.  Because the horizontal steps never fail
 * (meshed chunks are not on the window edge, WorldRender.cpp:235), resolving the three axes independently
 * gives the same cell.
 */
static OrCell or_probe(const GcOrWorld* w, int cx, int cy, int cz, int x, int y, int z)
{
    OrCell c;
    int wy = cy * GCOR_CHUNK_V + y;
    if (wy < 0) { c.where = -1; c.voxel = 0; c.column = 0; c.world_y = wy; return c; }
    if (wy >= GCOR_WORLD_HEIGHT) { c.where = 1; c.voxel = 0; c.column = 0; c.world_y = wy; return c; }

    int gx = cx * GCOR_CHUNK_H + x, gz = cz * GCOR_CHUNK_H + z;
    int ncx = gx >> 5, ncz = gz >> 5, ncy = wy >> 6;
    int64_t col = (int64_t)ncz * w->size_x + ncx;
    c.where = 0;
    c.world_y = wy;
    /* BlockIndex (World.cpp:278-281): x + 32*(y + 64*z) */
    c.voxel = (col * GCOR_WORLD_CHUNKS_Y + ncy) * GCOR_CHUNK_VOXELS + (gx & 31) + 32 * ((wy & 63) + 64 * (gz & 31));
    c.column = col * 1024 + (gz & 31) * 32 + (gx & 31);
    return c;
}

/* GetBlockSafe (World.cpp:422-431): below the world reads KILL_ZONE, above reads AIR. */
static int or_block(const GcOrWorld* w, OrCell c)
{
    if (c.where < 0) return w->kill_zone_id;
    if (c.where > 0) return 0;
    return w->blocks[c.voxel];
}

typedef struct OrColor { int l, s; } OrColor; /* Colori is (l,l,l,s), WorldRender.cpp:326 */

/* GetFinalLight (WorldRender.cpp:316-327) with GetSunlight (Lighting.cpp:69-79) and lightOutput[i] == 17*i (:314). */
static OrColor or_cell_light(const GcOrWorld* w, OrCell c)
{
    OrColor out;
    if (c.where < 0) { out.l = 0; out.s = 0; return out; }
    if (c.where > 0) { out.l = 255; out.s = 255; return out; }
    int sun;
    if (c.world_y >= w->surface[c.column]) sun = 15;               /* MAX_LIGHT, Lighting.h:6 */
    else { sun = w->sun[c.voxel]; if (sun == 0) sun = 1; }          /* MIN_LIGHT, Lighting.h:5 */
    out.l = 17 * (w->light[c.voxel] & 15);
    out.s = 17 * (sun & 15);
    return out;
}

/*
 * VertexLight (WorldRender.cpp:329-383).  n = outward normal axis (0 x, 1 y, 2 z); (dx,dy,dz) = corner signs.
 * a = face-adjacent cell, b/c = the two edge-adjacent cells, d = the corner cell, all in the layer in front
 * of the face.  Only b and c are tested for opacity; four-sample average shifts, three-sample divides by 3
 * (AverageColor, :304-312).
 */
static OrColor or_vertex_light(const GcOrWorld* w, int cx, int cy, int cz, int axis, int x, int y, int z, int dx, int dy, int dz)
{
    int bx, by, bz, ex, ey, ez; /* offsets of b and c from a */
    int ax = x, ay = y, az = z;
    if (axis == 0) { ax += dx; bx = 0; by = dy; bz = 0; ex = 0; ey = 0; ez = dz; }
    else if (axis == 1) { ay += dy; bx = dx; by = 0; bz = 0; ex = 0; ey = 0; ez = dz; }
    else { az += dz; bx = dx; by = 0; bz = 0; ex = 0; ey = dy; ez = 0; }

    OrCell ca = or_probe(w, cx, cy, cz, ax, ay, az);
    OrCell cb = or_probe(w, cx, cy, cz, ax + bx, ay + by, az + bz);
    OrCell cc = or_probe(w, cx, cy, cz, ax + ex, ay + ey, az + ez);

    int open_b = !w->table[or_block(w, cb)].opaque;
    int open_c = !w->table[or_block(w, cc)].opaque;

    OrColor la = or_cell_light(w, ca), lb = or_cell_light(w, cb), lc = or_cell_light(w, cc), r;
    if (open_b || open_c)
    {
        OrCell cd = or_probe(w, cx, cy, cz, ax + bx + ex, ay + by + ey, az + bz + ez);
        OrColor ld = or_cell_light(w, cd);
        r.l = (la.l + lb.l + lc.l + ld.l) >> 2;
        r.s = (la.s + lb.s + lc.s + ld.s) >> 2;
    }
    else
    {
        r.l = (la.l + lb.l + lc.l) / 3;
        r.s = (la.s + lb.s + lc.s) / 3;
    }
    return r;
}

/* ---- quad emission ---- */

/* Per face (order: up, down, front, back, right, left — WorldRender.cpp:468-556): normal axis, neighbour
 * offset, and the four corners as offsets from the voxel origin.  Corner sign = 2*corner-1 on each axis. */
typedef struct OrFace { int axis; int nx, ny, nz; int c[4][3]; } OrFace;
static const OrFace OR_FACES[6] = {
    { 1, 0, 1, 0, { {1,1,0}, {1,1,1}, {0,1,1}, {0,1,0} } },   /* up    :475-478 */
    { 1, 0, -1, 0, { {0,0,0}, {0,0,1}, {1,0,1}, {1,0,0} } },  /* down  :490-493 */
    { 2, 0, 0, 1, { {0,0,1}, {0,1,1}, {1,1,1}, {1,0,1} } },   /* front :505-508 */
    { 2, 0, 0, -1, { {1,0,0}, {1,1,0}, {0,1,0}, {0,0,0} } },  /* back  :520-523 */
    { 0, 1, 0, 0, { {1,0,1}, {1,1,1}, {1,1,0}, {1,0,0} } },   /* right :535-538 */
    { 0, -1, 0, 0, { {0,0,0}, {0,1,0}, {0,1,1}, {0,0,1} } },  /* left  :550-553 */
};
static const uint8_t OR_UV[4][2] = { {0,1}, {0,0}, {1,0}, {1,1} };

/* SetIndices (Mesh.cpp:32-50): o+2,o+1,o, o+3,o+2,o with o = (uint16)vertCount. */
static void or_push_indices(GcOrMesh* m, int type)
{
    uint16_t o = (uint16_t)m->vert_count;
    int n = m->index_count[type];
    if (n + 6 > GCOR_MAX_INDICES) { m->overflow = 1; return; }
    uint16_t* d = m->indices[type] + n;
    d[0] = (uint16_t)(o + 2); d[1] = (uint16_t)(o + 1); d[2] = o;
    d[3] = (uint16_t)(o + 3); d[4] = (uint16_t)(o + 2); d[5] = o;
    m->index_count[type] = n + 6;
}

/* The four vertices of one drawn face (WorldRender.cpp:475-553): 48 bytes at `o`. */
static void or_face_vertices(const GcOrWorld* w, int cx, int cy, int cz, const GcOrBlockInfo* info, int f, int x, int y, int z, uint8_t* o)
{
    const OrFace* face = &OR_FACES[f];
    for (int v = 0; v < 4; v++, o += 12)
    {
        const int* c = face->c[v];
        OrColor col = or_vertex_light(w, cx, cy, cz, face->axis, x, y, z, 2 * c[0] - 1, 2 * c[1] - 1, 2 * c[2] - 1);
        o[0] = (uint8_t)(x + c[0]); o[1] = (uint8_t)(y + c[1]); o[2] = (uint8_t)(z + c[2]);   /* VertexInfo, Mesh.h:37-44 */
        o[3] = OR_UV[v][0]; o[4] = OR_UV[v][1]; o[5] = info->textures[f];
        o[6] = o[7] = o[8] = (uint8_t)col.l; o[9] = (uint8_t)col.s;
        o[10] = info->alpha; o[11] = 0;
    }
}

/* BuildBlock (WorldRender.cpp:445-560). */
static int or_build_block(const GcOrWorld* w, int cx, int cy, int cz, GcOrMesh* m, int x, int y, int z, int block)
{
    const GcOrBlockInfo* info = &w->table[block];
    int added = 0;
    m->index_allocated[info->mesh_type] = 1; /* :454-459 */

    for (int f = 0; f < 6; f++)
    {
        const OrFace* face = &OR_FACES[f];
        /* GetNeighborBlocks (World.cpp:310-378) + CheckFace/CanDrawFace (WorldRender.cpp:387-402) */
        int adj = or_block(w, or_probe(w, cx, cy, cz, x + face->nx, y + face->ny, z + face->nz));
        if (!(info->cull < w->table[adj].cull && adj != block)) continue;
        added += 4;
        if (m->vert_count + 4 > GCOR_MAX_VERTICES) { m->overflow = 1; continue; }

        or_push_indices(m, info->mesh_type);
        or_face_vertices(w, cx, cy, cz, info, f, x, y, z, m->vertices + (size_t)m->vert_count * 12);
        m->vert_count += 4;
    }
    return added;
}

/* The same traversal with the reference's fixed-size MeshData lifted (the "stress" variant: no
 * 40 000-vertex cap, 32-bit indices — SetIndices' pattern o+2, o+1, o, o+3, o+2, o without the uint16 wrap).  The caller
 * provides the buffers: vertices (max_quads * 48 bytes) and one uint32 index array of max_quads * 6 elements per mesh type.
 * Returns the quad count, or -1 when a buffer is too small.  Not a reference behaviour: the reference asserts at the cap. */
int64_t gcor_build_chunk_uncapped(const GcOrWorld* w, int cx, int cy, int cz, uint8_t* vertices, uint32_t* const* indices, int64_t max_quads,
                                  int64_t* index_count)
{
    const uint16_t* blocks = w->blocks + (((int64_t)cz * w->size_x + cx) * GCOR_WORLD_CHUNKS_Y + cy) * GCOR_CHUNK_VOXELS;
    int64_t quads = 0;
    for (int t = 0; t < GCOR_MESH_TYPES; t++) index_count[t] = 0;
    for (int y = 0; y < GCOR_CHUNK_V; y++)
        for (int z = 0; z < GCOR_CHUNK_H; z++)
            for (int x = 0; x < GCOR_CHUNK_H; x++)
            {
                const int block = blocks[x + 32 * (y + 64 * z)];
                if (block == 0 || w->table[block].invisible) continue;
                const GcOrBlockInfo* info = &w->table[block];
                for (int f = 0; f < 6; f++)
                {
                    const OrFace* face = &OR_FACES[f];
                    int adj = or_block(w, or_probe(w, cx, cy, cz, x + face->nx, y + face->ny, z + face->nz));
                    if (!(info->cull < w->table[adj].cull && adj != block)) continue;
                    if (quads >= max_quads) return -1;
                    const uint32_t o = (uint32_t)(quads * 4);
                    uint32_t* d = indices[info->mesh_type] + index_count[info->mesh_type];
                    d[0] = o + 2; d[1] = o + 1; d[2] = o; d[3] = o + 3; d[4] = o + 2; d[5] = o;
                    index_count[info->mesh_type] += 6;
                    or_face_vertices(w, cx, cy, cz, info, f, x, y, z, vertices + (size_t)quads * 48);
                    quads++;
                }
            }
    return quads;
}

/* BuildChunkAsync (WorldRender.cpp:6-27): y outer, z middle, x inner; AIR and invisible blocks skipped. */
int gcor_build_chunk(const GcOrWorld* w, int cx, int cy, int cz, GcOrMesh* m)
{
    m->vert_count = 0;
    m->overflow = 0;
    for (int t = 0; t < GCOR_MESH_TYPES; t++) { m->index_count[t] = 0; m->index_allocated[t] = 0; }

    const uint16_t* blocks = w->blocks + (((int64_t)cz * w->size_x + cx) * GCOR_WORLD_CHUNKS_Y + cy) * GCOR_CHUNK_VOXELS;
    int total = 0;
    for (int y = 0; y < GCOR_CHUNK_V; y++)
        for (int z = 0; z < GCOR_CHUNK_H; z++)
            for (int x = 0; x < GCOR_CHUNK_H; x++)
            {
                int b = blocks[x + 32 * (y + 64 * z)];
                if (b == 0 || w->table[b].invisible) continue;
                total += or_build_block(w, cx, cy, cz, m, x, y, z, b);
            }
    return total;
}

/* ChunkOverflowed (WorldRender.cpp:412-439): would building this one block push totalVertices past the cap? */
int gcor_chunk_overflowed(const GcOrWorld* w, int cx, int cy, int cz, int x, int y, int z, int total_vertices)
{
    const uint16_t* blocks = w->blocks + (((int64_t)cz * w->size_x + cx) * GCOR_WORLD_CHUNKS_Y + cy) * GCOR_CHUNK_VOXELS;
    int b = blocks[x + 32 * (y + 64 * z)];
    if (b == 0) return 0;
    int added = 0;
    for (int f = 0; f < 6; f++)
    {
        int adj = or_block(w, or_probe(w, cx, cy, cz, x + OR_FACES[f].nx, y + OR_FACES[f].ny, z + OR_FACES[f].nz));
        if (w->table[b].cull < w->table[adj].cull && adj != b) added += 4;
    }
    return total_vertices + added > GCOR_MAX_VERTICES;
}

/* ---- multi-threaded timing (CPU baseline "port") ---- */
typedef struct OrJob
{
    const GcOrWorld* w;
    const int32_t* coords;
    int n;
    atomic_int* next;
    int64_t quads;
    int index;
} OrJob;

/* worker t runs on the t-th CPU this process may use (round robin): a timed run does not depend on where the scheduler
 * happens to put, or move, its threads */
static void or_pin_self(int t)
{
    cpu_set_t all, one;
    if (getenv("GCREF_NO_PIN") || sched_getaffinity(0, sizeof(all), &all) != 0) return;
    const int n = CPU_COUNT(&all);
    if (n <= 0) return;
    int want = t % n;
    for (int c = 0; c < CPU_SETSIZE; c++)
        if (CPU_ISSET(c, &all) && want-- == 0)
        {
            CPU_ZERO(&one); CPU_SET(c, &one);
            pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
            return;
        }
}

static void* or_worker(void* p)
{
    OrJob* job = (OrJob*)p;
    or_pin_self(job->index);
    GcOrMesh* m = (GcOrMesh*)malloc(sizeof(GcOrMesh));
    for (;;)
    {
        int i = atomic_fetch_add(job->next, 1);
        if (i >= job->n) break;
        gcor_build_chunk(job->w, job->coords[3 * i], job->coords[3 * i + 1], job->coords[3 * i + 2], m);
        job->quads += m->vert_count / 4;
    }
    free(m);
    return NULL;
}

double gcor_bench_build(const GcOrWorld* w, const int32_t* coords, int n, int threads, int64_t* quads_out)
{
    atomic_int next = 0;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    OrJob* jobs = (OrJob*)malloc(sizeof(OrJob) * (size_t)threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++)
    {
        jobs[t].w = w; jobs[t].coords = coords; jobs[t].n = n; jobs[t].next = &next; jobs[t].quads = 0; jobs[t].index = t;
        pthread_create(&th[t], NULL, or_worker, &jobs[t]);
    }
    int64_t quads = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); quads += jobs[t].quads; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    if (quads_out) *quads_out = quads;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

uint64_t gcor_fnv1a64(const void* data, uint64_t n, uint64_t h)
{
    const uint8_t* p = (const uint8_t*)data;
    if (h == 0) h = 14695981039346656037ull;
    for (uint64_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

/* ---- exhaustive parity check of a mesher batch against this oracle (multi-threaded) ----
 * desc: one 64-byte GcChunkOut per chunk (include/gcmesh.h: quad_base, vert_count, index_count[5], index_offset[5],
 * quads, flags), vertices / indices: the batch arenas as gcmesh_batch_download returns them.  Every chunk is rebuilt with
 * gcor_build_chunk (BuildChunkAsync, WorldRender.cpp:6-27) and compared byte for byte.
 * result[i]: 0 identical; bit 0 counts differ, bit 1 vertex bytes, bit 2 index bytes, bit 3 flags / segment outside the arena. */
typedef struct OrChunkOut
{
    uint32_t quad_base, vert_count, index_count[GCOR_MESH_TYPES], index_offset[GCOR_MESH_TYPES], quads, flags, reserved[2];
} OrChunkOut;

typedef struct OrCheckJob
{
    const GcOrWorld* w;
    const int32_t* coords;
    int n;
    const OrChunkOut* desc;
    const uint8_t* vertices;
    const uint16_t* indices;
    uint64_t arena_quads;
    uint8_t* result;
    atomic_int* next;
    int64_t bad, quads;
} OrCheckJob;

static void* or_check_worker(void* p)
{
    OrCheckJob* job = (OrCheckJob*)p;
    GcOrMesh* m = (GcOrMesh*)malloc(sizeof(GcOrMesh));
    for (;;)
    {
        int i = atomic_fetch_add(job->next, 1);
        if (i >= job->n) break;
        const OrChunkOut* d = job->desc + i;
        const int total = gcor_build_chunk(job->w, job->coords[3 * i], job->coords[3 * i + 1], job->coords[3 * i + 2], m);
        uint8_t r = 0;
        if (m->overflow)
        {
            /* over the reference's cap (an assert there, WorldRender.cpp:558): the mesher must flag it and emit nothing */
            if (!(d->flags & 1u) || d->vert_count != 0 || d->quads != (uint32_t)(total / 4)) r |= 8;
        }
        else
        {
            if (d->flags != 0) r |= 8;
            if (d->vert_count != (uint32_t)m->vert_count || d->quads != (uint32_t)(m->vert_count / 4)) r |= 1;
            for (int t = 0; t < GCOR_MESH_TYPES; t++)
                if (d->index_count[t] != (uint32_t)m->index_count[t]) r |= 1;
            if (!r && m->vert_count)
            {
                if ((uint64_t)d->quad_base + d->quads > job->arena_quads) r |= 8;
                else
                {
                    if (memcmp(job->vertices + (size_t)d->quad_base * 48, m->vertices, (size_t)m->vert_count * 12)) r |= 2;
                    for (int t = 0; t < GCOR_MESH_TYPES; t++)
                        if (m->index_count[t] &&
                            memcmp(job->indices + (size_t)d->quad_base * 6 + d->index_offset[t], m->indices[t], (size_t)m->index_count[t] * 2)) r |= 4;
                }
            }
            job->quads += m->vert_count / 4;
        }
        if (job->result) job->result[i] = r;
        if (r) job->bad++;
    }
    free(m);
    return NULL;
}

int64_t gcor_check_batch(const GcOrWorld* w, const int32_t* coords, int n, const void* desc, const void* vertices, const void* indices,
                         uint64_t arena_quads, int threads, uint8_t* result, int64_t* quads_checked)
{
    if (threads < 1) threads = 1;
    atomic_int next = 0;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    OrCheckJob* jobs = (OrCheckJob*)malloc(sizeof(OrCheckJob) * (size_t)threads);
    for (int t = 0; t < threads; t++)
    {
        OrCheckJob j = { w, coords, n, (const OrChunkOut*)desc, (const uint8_t*)vertices, (const uint16_t*)indices, arena_quads, result, &next, 0, 0 };
        jobs[t] = j;
        pthread_create(&th[t], NULL, or_check_worker, &jobs[t]);
    }
    int64_t bad = 0, quads = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; quads += jobs[t].quads; }
    free(th); free(jobs);
    if (quads_checked) *quads_checked = quads;
    return bad;
}

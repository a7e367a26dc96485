/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for the block-edit path: the decisions of SetBlock
 * (Code/World.cpp:548-582) with ChunkOverflowed (WorldRender.cpp:412-439), the surface entry of the edited column
 * (ComputeSurfaceAt, Lighting.cpp:5-17) and the chunks FlagChunkForUpdate marks (World.cpp:507-546).
 *
 * Light after an edit is not restated here as queues: the oracle of the relight is gcor_light_fill on the edited
 * blocks (gc_light_cpu.c) — the fill the reference itself runs when the column is next pre-processed.  The reference's
 * incremental RecomputeLight (Lighting.cpp:283-449) reaches that fill on ordinary terrain edits and can stop a few
 * cells short of it after a surface rise or an edit under water (tests/test_edit_oracle.py measures both against the
 * reference compiled verbatim).
 *
 * Parity status: PINNED for the decisions, the stored block, the surface entry (non-empty columns) and the 27-chunk
 * flag set, against oracle/_ref's SetBlock (tests/test_edit_oracle.py).
 */
#include "gc_oracle.h"

static int ed_block(const uint16_t* blocks, int size_x, int kill_zone, int x, int y, int z)
{
    /* GetBlockSafe (World.cpp:422-431) */
    if (y < 0) return kill_zone;
    if (y >= GCOR_WORLD_HEIGHT) return 0;
    const int64_t slot = ((int64_t)(z >> 5) * size_x + (x >> 5)) * GCOR_WORLD_CHUNKS_Y + (y >> 6);
    return blocks[slot * GCOR_CHUNK_VOXELS + (x & 31) + 32 * ((y & 63) + 64 * (z & 31))];
}

/* status codes as include/gcmesh.h GC_EDIT_* */
enum { ED_APPLIED = 0, ED_UNCHANGED = 1, ED_IGNORED = 2, ED_NOT_BUILT = 3, ED_OVERFLOW = 4 };

/* One SetBlock on window arrays (edited in place).  vert_table[slot] = chunk->totalVertices, < 0 = never built.
 * flagged[slot] is set to 1 for the chunks FlagChunkForUpdate marks (edge chunks included).  cull[id] = CullValue. */
int gcor_set_block(int size_x, int size_z, uint16_t* blocks, uint8_t* surface, const int32_t* vert_table, const uint8_t* cull,
                   int kill_zone, int wx, int wy, int wz, int block, uint8_t* flagged)
{
    (void)size_z;
    if (wy < 0 || wy >= GCOR_WORLD_HEIGHT) return ED_IGNORED;                               /* World.cpp:550 */
    const int cx = wx >> 5, cy = wy >> 6, cz = wz >> 5, rx = wx & 31, ry = wy & 63, rz = wz & 31;
    const int64_t slot = ((int64_t)cz * size_x + cx) * GCOR_WORLD_CHUNKS_Y + cy;
    uint16_t* cell = blocks + slot * GCOR_CHUNK_VOXELS + rx + 32 * (ry + 64 * rz);
    if (vert_table[slot] < 0) return ED_NOT_BUILT;                                          /* World.cpp:562-565 */
    if (*cell == block) return ED_UNCHANGED;                                                /* World.cpp:570 */
    *cell = (uint16_t)block;
    if (block != 0)
    {
        /* ChunkOverflowed (WorldRender.cpp:412-439) */
        static const int d[6][3] = { { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 }, { 1, 0, 0 }, { -1, 0, 0 } };
        int added = 0;
        for (int f = 0; f < 6; f++)
        {
            const int adj = ed_block(blocks, size_x, kill_zone, wx + d[f][0], wy + d[f][1], wz + d[f][2]);
            if (cull[block & 255] < cull[adj & 255] && adj != block) added += 4;
        }
        if (vert_table[slot] + added > GCOR_MAX_VERTICES) { *cell = 0; return ED_OVERFLOW; } /* World.cpp:572-573 */
    }
    /* ComputeSurfaceAt (Lighting.cpp:5-17).  The reference leaves the entry untouched when the column holds no block
     * at all; a from-scratch ComputeSurface of that column gives 0 (World.cpp:632), which is what is stored here. */
    int top = 0;
    for (int y = GCOR_WORLD_HEIGHT - 2; y >= 0; y--)
        if (ed_block(blocks, size_x, kill_zone, wx, y, wz) != 0) { top = y + 1; break; }
    surface[((int64_t)cz * size_x + cx) * 1024 + rz * 32 + rx] = (uint8_t)top;
    /* FlagChunkForUpdate (World.cpp:507-546) */
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++)
            {
                int tx = cx, ty = cy, tz = cz;
                if (rx + dx < 0) tx--; else if (rx + dx >= GCOR_CHUNK_H) tx++;
                if (rz + dz < 0) tz--; else if (rz + dz >= GCOR_CHUNK_H) tz++;
                if (ry + dy < 0) { if (ty - 1 >= 0) ty--; }
                else if (ry + dy >= GCOR_CHUNK_V) { if (ty + 1 < GCOR_WORLD_CHUNKS_Y) ty++; }
                flagged[((int64_t)tz * size_x + tx) * GCOR_WORLD_CHUNKS_Y + ty] = 1;
            }
    return ED_APPLIED;
}

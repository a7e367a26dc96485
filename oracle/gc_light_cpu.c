/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for the column pre-processing that produces the mesher's light inputs:
 * the surface map and the sunlight / blockLight flood fill of the reference (Code/Lighting.cpp), restated as the
 * fixpoint the reference's queues compute.
 *
 *   surface   ComputeSurfaceAt (Lighting.cpp:5-17): highest y in 0..254 whose block is not AIR, plus one; 0 if none.
 *   sunlight  SetLightNodes + ScatterSunlight (Lighting.cpp:247-281, 109-143, 81-93, 69-79): every cell at or above
 *             its column's surface reads 15 (stored value ignored).  Sky cells y in [surface, maxY] of the pre-processed
 *             (non-edge) columns are the seeds; a cell v pushes  l = eff(v) - lightStep[block(v)]  into each in-window,
 *             in-height neighbour n whose block is not opaque when  l > MIN_LIGHT (1)  and  eff(n) < l, storing l.  The
 *             result does not depend on the visiting order: it is the least fixpoint of that rule.
 *   blockLight  CheckForBlockLight + ScatterBlockLight (Lighting.cpp:231-245, 195-229, 95-107): emitters (lightEmitted
 *             > MIN_LIGHT) of the pre-processed columns at y < surface or y in [surface, maxY] are seeds, then the same
 *             push rule on the stored values.  CANONICAL FORM: the reference *assigns* the emission when it seeds a
 *             column (Lighting.cpp:240) after earlier columns may have pushed brighter light into the emitter cell, so
 *             its result depends on the column order; here a seed is max(stored, emission), i.e. the
 *             least fixpoint.  The two agree whenever no emitter lies within reach of a brighter one across a column
 *             border — tests/test_light_oracle.py pins this oracle to the verbatim reference on such worlds and on
 *             sunlight always.
 *
 * Parity status: PINNED against oracle/_ref/libgcref.so (gcref_compute_surface_all, gcref_light_all).
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CH 32
#define CV 64
#define WH 256
#define CVOX 65536
#define OPAQUE_STEP 255

static inline int64_t vox(int size_x, int lx, int ly, int lz)
{
    const int cx = lx >> 5, cz = lz >> 5;
    return (((int64_t)cz * size_x + cx) * 4 + (ly >> 6)) * CVOX + (lx & 31) + 32 * ((ly & 63) + 64 * (lz & 31));
}
static inline int surf(const uint8_t* surface, int size_x, int lx, int lz)
{
    return surface[((int64_t)(lz >> 5) * size_x + (lx >> 5)) * 1024 + (lz & 31) * CH + (lx & 31)];
}

/* ComputeSurface (Lighting.cpp:5-26) for every column of the window. */
void gcor_compute_surface(int size_x, int size_z, const uint16_t* blocks, uint8_t* surface)
{
    for (int lz = 0; lz < size_z * CH; lz++)
        for (int lx = 0; lx < size_x * CH; lx++)
        {
            int s = 0;
            for (int y = WH - 2; y >= 0; y--)
                if (blocks[vox(size_x, lx, y, lz)] != 0) { s = y + 1; break; }
            surface[((int64_t)(lz >> 5) * size_x + (lx >> 5)) * 1024 + (lz & 31) * CH + (lx & 31)] = (uint8_t)s;
        }
}

typedef struct Q { int32_t* v; size_t cap, rd, wr; } Q;
static void q_push(Q* q, int x, int y, int z)
{
    if (q->wr == q->cap)
    {
        if (q->rd > q->cap / 2) { memmove(q->v, q->v + q->rd * 3, (q->wr - q->rd) * 3 * sizeof(int32_t)); q->wr -= q->rd; q->rd = 0; }
        else { q->cap *= 2; q->v = (int32_t*)realloc(q->v, q->cap * 3 * sizeof(int32_t)); }
    }
    int32_t* p = q->v + q->wr * 3; p[0] = x; p[1] = y; p[2] = z; q->wr++;
}

static const int D6[6][3] = { {0,1,0}, {0,-1,0}, {0,0,1}, {0,0,-1}, {-1,0,0}, {1,0,0} };

/* ComputeMaxY (Lighting.cpp:28-67): the column cell's surface and its four horizontal neighbours', capped at 255 */
static int max_y(const uint8_t* surface, int size_x, int lx, int lz)
{
    int m = surf(surface, size_x, lx, lz), s;
    s = surf(surface, size_x, lx, lz + 1); if (s > m) m = s;
    s = surf(surface, size_x, lx, lz - 1); if (s > m) m = s;
    s = surf(surface, size_x, lx - 1, lz); if (s > m) m = s;
    s = surf(surface, size_x, lx + 1, lz); if (s > m) m = s;
    return m > WH - 1 ? WH - 1 : m;
}

/*
 * step[id]: lightStep (Block.h:95), 255 = opaque (INT_MAX, Block.cpp:65-68); emitted[id]: lightEmitted.
 * sun / light are overwritten (zeroed first, as freshly generated chunks are: World.cpp:632 / value-initialised).
 */
void gcor_light_fill(int size_x, int size_z, const uint16_t* blocks, const uint8_t* surface, const uint8_t* step, const uint8_t* emitted,
                     uint8_t* sun, uint8_t* light)
{
    const int64_t total = (int64_t)size_x * size_z * 4 * CVOX;
    const int wx = size_x * CH, wz = size_z * CH;
    memset(sun, 0, (size_t)total);
    memset(light, 0, (size_t)total);
    Q sq = { (int32_t*)malloc((1u << 20) * 3 * sizeof(int32_t)), 1u << 20, 0, 0 };
    Q bq = { (int32_t*)malloc((1u << 16) * 3 * sizeof(int32_t)), 1u << 16, 0, 0 };

    /* seeds of every pre-processed (non-edge) column (SetLightNodes, Lighting.cpp:247-281) */
    for (int lz = CH; lz < wz - CH; lz++)
        for (int lx = CH; lx < wx - CH; lx++)
        {
            const int s = surf(surface, size_x, lx, lz), my = max_y(surface, size_x, lx, lz);
            for (int y = 0; y <= my; y++)
            {
                if (y >= s) q_push(&sq, lx, y, lz);
                const int64_t vi = vox(size_x, lx, y, lz);
                const int e = emitted[blocks[vi] & 255];
                if (e > 1) { if (light[vi] < e) light[vi] = (uint8_t)e; q_push(&bq, lx, y, lz); }
            }
        }
    /* ScatterSunlight (Lighting.cpp:109-143) */
    while (sq.rd != sq.wr)
    {
        const int32_t* p = sq.v + sq.rd * 3; const int x = p[0], y = p[1], z = p[2]; sq.rd++;
        const int64_t vi = vox(size_x, x, y, z);
        const int st = step[blocks[vi] & 255];
        if (st == OPAQUE_STEP) continue;
        const int eff = y >= surf(surface, size_x, x, z) ? 15 : (sun[vi] ? sun[vi] : 1);   /* GetSunlight :69-79 */
        const int l = eff - st;
        if (l <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            const int nx = x + D6[d][0], ny = y + D6[d][1], nz = z + D6[d][2];
            if (ny < 0 || ny >= WH || nx < 0 || nz < 0 || nx >= wx || nz >= wz) continue;
            const int64_t ni = vox(size_x, nx, ny, nz);
            if (step[blocks[ni] & 255] == OPAQUE_STEP) continue;
            const int ne = ny >= surf(surface, size_x, nx, nz) ? 15 : (sun[ni] ? sun[ni] : 1);
            if (ne < l) { sun[ni] = (uint8_t)l; q_push(&sq, nx, ny, nz); }                  /* SetMaxSunlight :81-93 */
        }
    }
    /* ScatterBlockLight (Lighting.cpp:195-229) */
    while (bq.rd != bq.wr)
    {
        const int32_t* p = bq.v + bq.rd * 3; const int x = p[0], y = p[1], z = p[2]; bq.rd++;
        const int64_t vi = vox(size_x, x, y, z);
        const int st = step[blocks[vi] & 255];
        if (st == OPAQUE_STEP) continue;
        const int l = (int)light[vi] - st;
        if (l <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            const int nx = x + D6[d][0], ny = y + D6[d][1], nz = z + D6[d][2];
            if (ny < 0 || ny >= WH || nx < 0 || nz < 0 || nx >= wx || nz >= wz) continue;
            const int64_t ni = vox(size_x, nx, ny, nz);
            if (step[blocks[ni] & 255] == OPAQUE_STEP) continue;
            if ((int)light[ni] < l) { light[ni] = (uint8_t)l; q_push(&bq, nx, ny, nz); }   /* SetMaxBlockLight :95-107 */
        }
    }
    free(sq.v); free(bq.v);
}

/* lightStep / lightEmitted of the reference's block table (Block.cpp:139-276), 255 = opaque */
void gcor_default_light_rules(uint8_t* step, uint8_t* emitted)
{
    for (int i = 0; i < 256; i++) { step[i] = OPAQUE_STEP; emitted[i] = 0; }
    /* ids: Block.h:11-38 */
    step[0] = 1;                      /* AIR        Block.cpp:146 */
    step[8] = 2;                      /* WATER      :167 */
    step[13] = 2;                     /* ICE        :211 */
    step[14] = 1; emitted[14] = 15;   /* LANTERN    :217-218 */
    step[17] = 1; emitted[17] = 10;   /* MAGMA      :235-236 */
    step[20] = 1;                     /* KILL_ZONE  :252 */
    step[21] = 1; emitted[21] = 10;   /* LAVA       :261-262 */
    step[22] = 1;                     /* GLASS      :272 */
}

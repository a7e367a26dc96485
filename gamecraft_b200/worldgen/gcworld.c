/*
 * gcworld — deterministic synthetic-world producer for the chunk mesher (host C, pthreads).
 *
 * The mesher's inputs are a window of columns, each 4 chunk bricks (uint16 blocks, uint8 sunlight,
 * uint8 blockLight) plus a 32x32 uint8 surface map.  The reference builds them with FastNoiseSIMD +
 * MSVC rand(), neither of which is reproducible here (not vendored / thread-timing dependent), so this
 * file restates the reference's *rules* with its own seeded simplex noise and PRNG:
 *   terrain layering   Code/Generation.cpp:83-226 (forest), :228-369 (islands), :412-556 (snow), :558-645 (desert),
 *                      :647-770 (volcanic), :772-836 (flat); Code/Dungeon.cpp:17-31 (dungeon)
 *   trees              Code/Generation.cpp:50-64, :204-215
 *   island fall-off    Code/Generation.cpp:66-81
 *   surface map        Code/Lighting.cpp:5-26
 *   light flood fill   Code/Lighting.cpp:28-67, :81-143, :195-281
 * plus the worst-case patterns of BASELINE.json's config 4.  It is an INPUT PRODUCER (bench / tests /
 * examples); it is not on the meshing path and the mesher library does not link it.
 *
 * Storage (shared with libgcmesh's device world and the oracle):
 *   slot(cx,cy,cz) = (cz*size_x + cx)*4 + cy ;  voxel = slot*65536 + x + 32*(y + 64*z)
 *   surface[(cz*size_x + cx)*1024 + z*32 + x]
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define CH 32
#define CV 64
#define CVOX 65536
#define WCY 4
#define WH 256

enum
{
    B_AIR = 0, B_GRASS, B_DIRT, B_STONE, B_CRATE, B_SAND, B_STONE_BRICK, B_METAL_CRATE, B_WATER, B_WOOD,
    B_LEAVES, B_CLAY, B_SNOW, B_ICE, B_LANTERN, B_TRAMPOLINE, B_CACTUS, B_MAGMA, B_COOLED_MAGMA,
    B_OBSIDIAN, B_KILL_ZONE, B_LAVA, B_GLASS, B_SHOCKER, B_COUNT
}; /* Code/Block.h:11-38 */

typedef struct GcwWorld
{
    int32_t size_x, size_z;
    int32_t origin_cx, origin_cz; /* world-space column of window column (0,0): noise is sampled in world space */
    uint16_t* blocks;
    uint8_t* sun;
    uint8_t* light;
    uint8_t* surface;
} GcwWorld;

enum { GCW_FLAT = 0, GCW_FOREST = 1, GCW_ISLANDS = 2, GCW_CHECKER = 3, GCW_NOISE = 4, GCW_STRESS = 5, GCW_MIXED = 6, GCW_EMPTY = 7,
       GCW_SNOW = 8, GCW_DESERT = 9, GCW_VOLCANIC = 10, GCW_DUNGEON = 11 };


/* ------------------------------------------------------------------------------------------ */
/* tiny parallel-for (atomic work counter over pthreads)                                        */
/* ------------------------------------------------------------------------------------------ */
typedef void (*ParFn)(void* ctx, int index, int worker);
typedef struct ParJob { ParFn fn; void* ctx; int n; atomic_int* next; int worker; } ParJob;
static void* par_worker(void* p)
{
    ParJob* j = (ParJob*)p;
    for (;;)
    {
        int i = atomic_fetch_add(j->next, 1);
        if (i >= j->n) break;
        j->fn(j->ctx, i, j->worker);
    }
    return NULL;
}
static void par_for(int n, int threads, ParFn fn, void* ctx)
{
    if (threads > n) threads = n;
    if (threads <= 1) { for (int i = 0; i < n; i++) fn(ctx, i, 0); return; }
    atomic_int next = 0;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    ParJob* jobs = (ParJob*)malloc(sizeof(ParJob) * (size_t)threads);
    for (int t = 0; t < threads; t++)
    {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].n = n; jobs[t].next = &next; jobs[t].worker = t;
        pthread_create(&th[t], NULL, par_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------------------------------ */
/* hashing / PRNG                                                                              */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

static inline uint32_t hash3(uint32_t seed, int32_t a, int32_t b, int32_t c)
{
    uint32_t h = seed ^ 0x9E3779B9u;
    h = mix32(h + (uint32_t)a * 0x85EBCA6Bu);
    h = mix32(h + (uint32_t)b * 0xC2B2AE35u);
    h = mix32(h + (uint32_t)c * 0x27D4EB2Fu);
    return h;
}

typedef struct Rng { uint64_t s; } Rng;
static inline uint32_t rng_next(Rng* r)
{
    r->s += 0x9E3779B97F4A7C15ull;
    uint64_t z = r->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}
/* RandRange(int,int) (Code/Random.h:32-35): inclusive bounds */
static inline int rng_range(Rng* r, int lo, int hi) { return lo + (int)(rng_next(r) % (uint32_t)(hi + 1 - lo)); }

/* ------------------------------------------------------------------------------------------ */
/* simplex noise (own implementation, hashed gradients) + the three fractal shapes              */
/* ------------------------------------------------------------------------------------------ */
static inline float grad2(uint32_t h, float x, float y)
{
    static const float G[8][2] = { {1,0}, {-1,0}, {0,1}, {0,-1}, {0.70710678f,0.70710678f}, {-0.70710678f,0.70710678f}, {0.70710678f,-0.70710678f}, {-0.70710678f,-0.70710678f} };
    const float* g = G[h & 7];
    return g[0] * x + g[1] * y;
}

static float simplex2(uint32_t seed, float x, float y)
{
    const float F2 = 0.36602540f, G2 = 0.21132487f;
    float s = (x + y) * F2;
    int i = (int)floorf(x + s), j = (int)floorf(y + s);
    float t = (float)(i + j) * G2;
    float x0 = x - ((float)i - t), y0 = y - ((float)j - t);
    int i1 = x0 > y0, j1 = !i1;
    float x1 = x0 - (float)i1 + G2, y1 = y0 - (float)j1 + G2;
    float x2 = x0 - 1.0f + 2.0f * G2, y2 = y0 - 1.0f + 2.0f * G2;
    float n = 0.0f, a;
    a = 0.5f - x0 * x0 - y0 * y0; if (a > 0) { a *= a; n += a * a * grad2(hash3(seed, i, j, 0), x0, y0); }
    a = 0.5f - x1 * x1 - y1 * y1; if (a > 0) { a *= a; n += a * a * grad2(hash3(seed, i + i1, j + j1, 0), x1, y1); }
    a = 0.5f - x2 * x2 - y2 * y2; if (a > 0) { a *= a; n += a * a * grad2(hash3(seed, i + 1, j + 1, 0), x2, y2); }
    n *= 70.0f;
    return n < -1.0f ? -1.0f : (n > 1.0f ? 1.0f : n);
}

static inline float grad3(uint32_t h, float x, float y, float z)
{
    static const float G[12][3] = { {1,1,0}, {-1,1,0}, {1,-1,0}, {-1,-1,0}, {1,0,1}, {-1,0,1}, {1,0,-1}, {-1,0,-1}, {0,1,1}, {0,-1,1}, {0,1,-1}, {0,-1,-1} };
    const float* g = G[h % 12u];
    return g[0] * x + g[1] * y + g[2] * z;
}

static float simplex3(uint32_t seed, float x, float y, float z)
{
    const float F3 = 1.0f / 3.0f, G3 = 1.0f / 6.0f;
    float s = (x + y + z) * F3;
    int i = (int)floorf(x + s), j = (int)floorf(y + s), k = (int)floorf(z + s);
    float t = (float)(i + j + k) * G3;
    float x0 = x - ((float)i - t), y0 = y - ((float)j - t), z0 = z - ((float)k - t);
    int i1, j1, k1, i2, j2, k2;
    if (x0 >= y0)
    {
        if (y0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
        else if (x0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1; }
        else { i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1; }
    }
    else
    {
        if (y0 < z0) { i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1; }
        else if (x0 < z0) { i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1; }
        else { i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
    }
    float x1 = x0 - i1 + G3, y1 = y0 - j1 + G3, z1 = z0 - k1 + G3;
    float x2 = x0 - i2 + 2 * G3, y2 = y0 - j2 + 2 * G3, z2 = z0 - k2 + 2 * G3;
    float x3 = x0 - 1 + 3 * G3, y3 = y0 - 1 + 3 * G3, z3 = z0 - 1 + 3 * G3;
    float n = 0.0f, a;
    a = 0.6f - x0 * x0 - y0 * y0 - z0 * z0; if (a > 0) { a *= a; n += a * a * grad3(hash3(seed, i, j, k), x0, y0, z0); }
    a = 0.6f - x1 * x1 - y1 * y1 - z1 * z1; if (a > 0) { a *= a; n += a * a * grad3(hash3(seed, i + i1, j + j1, k + k1), x1, y1, z1); }
    a = 0.6f - x2 * x2 - y2 * y2 - z2 * z2; if (a > 0) { a *= a; n += a * a * grad3(hash3(seed, i + i2, j + j2, k + k2), x2, y2, z2); }
    a = 0.6f - x3 * x3 - y3 * y3 - z3 * z3; if (a > 0) { a *= a; n += a * a * grad3(hash3(seed, i + 1, j + 1, k + 1), x3, y3, z3); }
    n *= 32.0f;
    return n < -1.0f ? -1.0f : (n > 1.0f ? 1.0f : n);
}

static float fractal_bounding(int octaves)
{
    float amp = 0.5f, sum = 1.0f;
    for (int i = 1; i < octaves; i++) { sum += amp; amp *= 0.5f; }
    return 1.0f / sum;
}

/* ridged multifractal, 2-D */
static float ridged2(uint32_t seed, float x, float y, int octaves)
{
    float sum = 1.0f - fabsf(simplex2(seed, x, y)), amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x *= 2.0f; y *= 2.0f; amp *= 0.5f;
        sum -= (1.0f - fabsf(simplex2(seed + (uint32_t)i, x, y))) * amp;
    }
    return sum;
}

/* billow, 2-D */
static float billow2(uint32_t seed, float x, float y, int octaves)
{
    float sum = fabsf(simplex2(seed, x, y)) * 2.0f - 1.0f, amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x *= 2.0f; y *= 2.0f; amp *= 0.5f;
        sum += (fabsf(simplex2(seed + (uint32_t)i, x, y)) * 2.0f - 1.0f) * amp;
    }
    return sum * fractal_bounding(octaves);
}

/* fBm, 2-D */
static float fbm2(uint32_t seed, float x, float y, int octaves)
{
    float sum = simplex2(seed, x, y), amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x *= 2.0f; y *= 2.0f; amp *= 0.5f;
        sum += simplex2(seed + (uint32_t)i, x, y) * amp;
    }
    return sum * fractal_bounding(octaves);
}

/* fBm, 3-D */
static float fbm3(uint32_t seed, float x, float y, float z, int octaves)
{
    float sum = simplex3(seed, x, y, z), amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x *= 2.0f; y *= 2.0f; z *= 2.0f; amp *= 0.5f;
        sum += simplex3(seed + (uint32_t)i, x, y, z) * amp;
    }
    return sum * fractal_bounding(octaves);
}

/* The per-column random stream of the tree / cactus placement, for the same test infrastructure: a column's start state and the
 * next draw (rng_range(lo, hi) = lo + draw % (hi + 1 - lo), RandRange's formula). */
uint64_t gcw_column_rng(uint32_t seed, int32_t wcx, int32_t wcz) { return ((uint64_t)hash3(seed, wcx, wcz, 77) << 32) | hash3(seed, wcz, wcx, 78); }
uint32_t gcw_rng_next(uint64_t* state) { Rng r = { *state }; const uint32_t v = rng_next(&r); *state = r.s; return v; }

/* The noise fields by number, for test infrastructure that runs the reference's own generator code over them
 * (oracle/ref: its FastNoiseSIMD stand-in calls back into this): 0 simplex (x, y), 1 billow (x, y), 2 ridged (x, y),
 * 3 fBm (x, y), 4 fBm (x, y, z). */
float gcw_noise(uint32_t seed, int fn, int octaves, float x, float y, float z)
{
    switch (fn)
    {
        case 0: return simplex2(seed, x, y);
        case 1: return billow2(seed, x, y, octaves);
        case 2: return ridged2(seed, x, y, octaves);
        case 3: return fbm2(seed, x, y, octaves);
        default: return fbm3(seed, x, y, z, octaves);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* column helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */
static inline int64_t col_index(const GcwWorld* w, int cx, int cz) { return (int64_t)cz * w->size_x + cx; }
static inline int64_t voxel_index(const GcwWorld* w, int cx, int cz, int x, int wy, int z)
{
    return (col_index(w, cx, cz) * WCY + (wy >> 6)) * CVOX + x + 32 * ((wy & 63) + 64 * z);
}
static inline void set_block(GcwWorld* w, int cx, int cz, int x, int wy, int z, int b)
{
    if (wy < 0 || wy >= WH) return;
    w->blocks[voxel_index(w, cx, cz, x, wy, z)] = (uint16_t)b;
}

static void box(GcwWorld* w, int cx, int cz, int x0, int y0, int z0, int x1, int y1, int z1, int b)
{
    for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) set_block(w, cx, cz, x, y, z, b);
}

/* CreateTree (Generation.cpp:50-64): trunk then five leaf boxes */
static void tree(GcwWorld* w, int cx, int cz, Rng* r, int bx, int by, int bz)
{
    int h = rng_range(r, 3, 5);
    for (int j = 0; j < h; j++) set_block(w, cx, cz, bx, by + j, bz, B_WOOD);
    int sy = by + h;
    box(w, cx, cz, bx - 1, sy, bz - 1, bx + 1, sy, bz + 1, B_LEAVES);
    box(w, cx, cz, bx - 2, sy + 1, bz - 1, bx + 2, sy + 2, bz + 1, B_LEAVES);
    box(w, cx, cz, bx - 1, sy + 1, bz + 2, bx + 1, sy + 2, bz + 2, B_LEAVES);
    box(w, cx, cz, bx - 1, sy + 1, bz - 2, bx + 1, sy + 2, bz - 2, B_LEAVES);
    box(w, cx, cz, bx - 1, sy + 3, bz - 1, bx + 1, sy + 3, bz + 1, B_LEAVES);
}

/* IsWithinIsland (Generation.cpp:66-81) */
static int within_island(int wx, int wz, int radius, int falloff, float* p)
{
    int v = (int)sqrt((double)wx * wx + (double)wz * wz);
    if (v < radius)
    {
        *p = 1.0f;
        if (v > falloff) *p = 1.0f - ((float)(v - falloff) / (float)(radius - falloff));
        return 1;
    }
    return 0;
}

static inline float scurve3(float a) { return a * a * (3.0f - 2.0f * a); }

/* The five noise terrains (Generation.cpp:83-226 forest, :228-369 islands, :412-556 snow, :558-645 desert, :647-770 volcanic)
 * for one column: a height field blended from two fractal fields by a biome field (desert: one field), the island fall-off,
 * stone pockets where a 3-D fBm drops to 0.2 or below from four layers under the surface (not in the desert), the biome's
 * layering, then its decorations (trees / cacti) from the column's own PRNG stream. */
enum { T_FOREST = 0, T_ISLANDS = 1, T_SNOW = 2, T_DESERT = 3, T_VOLCANIC = 4 };
static void gen_terrain_column(GcwWorld* w, int cx, int cz, uint32_t seed, int biome_kind, int radius, int falloff)
{
    static const int SEA[5] = { 10, 20, 20, 10, 12 };
    const int sea = SEA[biome_kind];
    int wcx = w->origin_cx + cx, wcz = w->origin_cz + cz;
    int height[CH * CH];
    uint8_t ice[CH * CH];
    int max_y = 0;
    memset(ice, 0, sizeof(ice));

    for (int z = 0; z < CH; z++)
    {
        for (int x = 0; x < CH; x++)
        {
            int wx = wcx * CH + x, wz = wcz * CH + z;
            float p;
            int h = 0, m = sea;
            if (within_island(wx, wz, radius, falloff, &p))
            {
                float bx = (float)wx, bz = (float)wz;
                float terrain;
                if (biome_kind == T_DESERT)
                {
                    /* one billow field, frequency 0.05 sampled at scale 0.25, the library's default three octaves (:566-569,:585-586) */
                    float base = (billow2(seed + 202u, bx * 0.05f * 0.25f, bz * 0.05f * 0.25f, 3) + 1.0f) * 0.5f;
                    terrain = base * 0.2f * 30.0f + 20.0f;
                }
                else if (biome_kind == T_SNOW)
                {
                    /* biome: 3-octave fBm at 0.005; sea floor from the billow field, land from the ridged one, blended over
                     * -0.4 .. 0; the upper part of the blend band freezes the sea surface (:431-479) */
                    float biome = fbm2(seed + 101u, bx * 0.005f, bz * 0.005f, 3);
                    float floor_ = (billow2(seed + 202u, bx * 0.025f * 0.5f, bz * 0.025f * 0.5f, 4) + 1.0f) * 0.5f;
                    floor_ = floor_ * 0.2f * 20.0f + 5.0f;
                    float land = (ridged2(seed + 303u, bx * 0.015f * 0.5f, bz * 0.015f * 0.5f, 4) + 1.0f) * 0.5f;
                    land = land * 0.2f * 10.0f + 20.0f;
                    const float lower = -0.4f, upper = 0.0f;
                    if (biome < lower) terrain = floor_;
                    else if (biome > upper) terrain = land;
                    else
                    {
                        if (biome >= -0.25f && biome <= 0.0f) ice[z * CH + x] = 1;
                        terrain = floor_ + scurve3((biome - lower) / (upper - lower)) * (land - floor_);
                    }
                }
                else
                {
                    const float mscale = biome_kind == T_FOREST ? 30.0f : 60.0f, mbase = biome_kind == T_FOREST ? 10.0f : 20.0f;
                    const float bfreq = biome_kind == T_VOLCANIC ? 0.005f : 0.01f;   /* (:103-104 forest, :248-249 islands, :667-668 volcanic) */
                    float biome = simplex2(seed + 101u, bx * bfreq, bz * bfreq);
                    float flat = (billow2(seed + 202u, bx * 0.025f * 0.5f, bz * 0.025f * 0.5f, 4) + 1.0f) * 0.5f;
                    flat = flat * 0.2f * 30.0f + 10.0f;
                    float mount = (ridged2(seed + 303u, bx * 0.015f * 0.5f, bz * 0.015f * 0.5f, 4) + 1.0f) * 0.5f;
                    if (mount < 0.0f) mount = 0.0f;
                    mount = powf(mount, 3.5f) * mscale + mbase;
                    if (biome < 0.0f) terrain = flat;
                    else if (biome > 0.6f) terrain = mount;
                    else terrain = flat + scurve3(biome / 0.6f) * (mount - flat);
                }
                h = (int)(terrain * p);
                if (h > WH - 12) h = WH - 12;
                if (h > m) m = h;
            }
            height[z * CH + x] = h;
            if (m > max_y) max_y = m;
        }
    }

    for (int z = 0; z < CH; z++)
    {
        for (int x = 0; x < CH; x++)
        {
            int h = height[z * CH + x];
            int wz = wcz * CH + z;
            for (int wy = 0; wy <= max_y; wy++)
            {
                int b = B_AIR;
                int stone = 0;
                if (biome_kind != T_DESERT && wy <= h - 4)
                {
                    /* What the reference's code does, not what it means: the 3-D noise set has maxY + 1 layers per x
                     * (GetNoise3D(noise, x, z, maxY + 1), Generation.cpp:157-161) and is read with a stride of maxY
                     * (GetNoiseValue3D(comp, x, wY, z, maxY), :16-18, :179), so cell (x, wY) gets the sample of
                     * (x', y') with x' * (maxY + 1) + y' = x * maxY + wY. */
                    const int q = x * max_y + wy, sx = q / (max_y + 1), sy = q % (max_y + 1);
                    float comp = (fbm3(seed + 404u, (float)(wcx * CH + sx) * 0.015f * 0.2f, (float)sy * 0.015f * 0.2f, (float)wz * 0.015f * 0.2f, 2) + 1.0f) * 0.5f;
                    stone = comp <= 0.2f;
                }
                if (stone) b = B_STONE;
                else if (biome_kind == T_SNOW)
                {
                    /* (:525-539) the sea-level layer is tested before "below the surface": a column taller than the sea holds
                     * water (or ice) at y = 20 inside its dirt, as the reference's branch order has it */
                    if (wy == h) b = B_SNOW;
                    else if (wy > h && wy < sea) b = B_WATER;
                    else if (wy == sea) b = ice[z * CH + x] ? B_ICE : B_WATER;
                    else if (wy < h) b = B_DIRT;
                }
                else if (biome_kind == T_DESERT)
                {
                    if (wy <= h) b = B_SAND;                      /* (:610-613) */
                    else if (wy <= sea) b = B_WATER;
                }
                else if (biome_kind == T_VOLCANIC)
                {
                    if (wy <= h) b = B_OBSIDIAN;                  /* (:750-753) */
                    else if (wy <= sea) b = B_LAVA;
                }
                else if (wy == h) b = B_GRASS;
                else if (wy > h && wy <= sea) b = B_WATER;
                else if (wy < h) b = B_DIRT;
                if (b != B_AIR) set_block(w, cx, cz, x, wy, z, b);
            }
        }
    }

    Rng r = { ((uint64_t)hash3(seed, wcx, wcz, 77) << 32) | hash3(seed, wcz, wcx, 78) };
    if (biome_kind == T_FOREST || biome_kind == T_ISLANDS)
    {
        int trees = biome_kind == T_ISLANDS ? rng_range(&r, 1, 3) : rng_range(&r, 3, 6);
        for (int i = 0; i < trees; i++)
        {
            int rx = rng_range(&r, 3, CH - 4), rz = rng_range(&r, 3, CH - 4);
            int s = height[rz * CH + rx];
            if (s > sea) tree(w, cx, cz, &r, rx, s + 1, rz);
        }
    }
    else if (biome_kind == T_DESERT)
    {
        /* zero to two cacti anywhere in the column, two to four blocks on top of the surface entry — under water too (:630-641) */
        int cacti = rng_range(&r, 0, 2);
        for (int i = 0; i < cacti; i++)
        {
            int rx = rng_range(&r, 0, CH - 1), rz = rng_range(&r, 0, CH - 1);
            int ch = rng_range(&r, 2, 4);
            for (int j = 1; j <= ch; j++) set_block(w, cx, cz, rx, height[rz * CH + rx] + j, rz, B_CACTUS);
        }
    }
}

/* GenerateDungeon (Dungeon.cpp:17-31): in the column's lowest chunk a stone floor and four stone walls eleven blocks high
 * along the column's border; no noise, no randomness. */
static void gen_dungeon_column(GcwWorld* w, int cx, int cz)
{
    const int wall = 10, lim = CH - 1;
    box(w, cx, cz, 0, 0, 0, lim, 0, lim, B_STONE);
    box(w, cx, cz, 0, 0, 0, lim, wall, 0, B_STONE);
    box(w, cx, cz, 0, 0, 0, 0, wall, lim, B_STONE);
    box(w, cx, cz, 0, 0, lim, lim, wall, lim, B_STONE);
    box(w, cx, cz, lim, 0, 0, lim, wall, lim, B_STONE);
}

/* GenerateFlatTerrain (Generation.cpp:772-836): sea level 12; inside the island grass at (int)(12 * p) over dirt, water up
 * to sea level; outside it height 0.  With radius -> infinity: dirt 0..11, grass 12. */
static void gen_flat_column(GcwWorld* w, int cx, int cz, int radius, int falloff)
{
    const int sea = 12;
    int wcx = w->origin_cx + cx, wcz = w->origin_cz + cz;
    for (int z = 0; z < CH; z++) for (int x = 0; x < CH; x++)
    {
        float p;
        int h = 0;
        if (within_island(wcx * CH + x, wcz * CH + z, radius, falloff, &p)) h = (int)((float)sea * p);
        for (int wy = 0; wy <= sea; wy++)
        {
            if (wy == h) set_block(w, cx, cz, x, wy, z, B_GRASS);
            else if (wy > h) set_block(w, cx, cz, x, wy, z, B_WATER);
            else set_block(w, cx, cz, x, wy, z, B_DIRT);
        }
    }
}

/* Blocks used to hit every cull class / mesh type / light behaviour (Block.cpp:139-276). */
static const uint8_t MIXED_PALETTE[16] = {
    B_STONE, B_GLASS, B_ICE, B_WATER, B_MAGMA, B_LANTERN, B_LAVA, B_LEAVES,
    B_GRASS, B_CACTUS, B_WOOD, B_SNOW, B_SAND, B_OBSIDIAN, B_GLASS, B_ICE
};

/*
 * Config 4(i): checkerboard in y-layers 1..3 of every chunk: 1536 isolated voxels, 9216 quads <= 10000 cap
 * (Mesh.h:5-6).  Block ids cycle through the palette so all five index streams are hit.
 */
static void gen_checker_column(GcwWorld* w, int cx, int cz, uint32_t seed)
{
    int wcx = w->origin_cx + cx, wcz = w->origin_cz + cz;
    for (int cy = 0; cy < WCY; cy++)
        for (int y = 1; y <= 3; y++)
            for (int z = 0; z < CH; z++)
                for (int x = 0; x < CH; x++)
                    if (((x + y + z) & 1) == 0)
                        set_block(w, cx, cz, x, cy * CV + y, z, MIXED_PALETTE[hash3(seed, wcx * CH + x, cy * CV + y, wcz * CH + z) & 15]);
}

/* Config 4(i) second pattern: seeded scattered voxels, <= 1666 per chunk (<= 9996 quads). */
static void gen_noise_column(GcwWorld* w, int cx, int cz, uint32_t seed)
{
    int wcx = w->origin_cx + cx, wcz = w->origin_cz + cz;
    for (int cy = 0; cy < WCY; cy++)
    {
        Rng r = { ((uint64_t)hash3(seed, wcx, cy, wcz) << 32) | 0x1234567u };
        int n = 1200 + (int)(rng_next(&r) % 467u); /* 1200..1666 */
        for (int i = 0; i < n; i++)
        {
            uint32_t v = rng_next(&r);
            int x = v & 31, z = (v >> 5) & 31, y = (v >> 10) & 63;
            int wy = cy * CV + y;
            if (wy >= WH - 1) continue; /* keep y=255 free: ComputeSurfaceAt ignores it (Lighting.cpp:9) */
            set_block(w, cx, cz, x, wy, z, MIXED_PALETTE[(v >> 16) & 15]);
        }
    }
}

/* Config 4(ii): full 3-D checkerboard — 196608 quads per chunk, far over the cap (stress, no parity). */
static void gen_stress_column(GcwWorld* w, int cx, int cz)
{
    for (int wy = 0; wy < WH; wy++) for (int z = 0; z < CH; z++) for (int x = 0; x < CH; x++)
        if (((x + wy + z) & 1) == 0) set_block(w, cx, cz, x, wy, z, B_STONE);
}

/*
 * "Mixed": islands-rule terrain with blocks of every class injected on chunk borders, corners, the y=63/64
 * seam and the world top — the validation pattern of SURVEY.md A.10.
 */
static void gen_mixed_column(GcwWorld* w, int cx, int cz, uint32_t seed, int radius, int falloff)
{
    gen_terrain_column(w, cx, cz, seed, 1, radius, falloff);
    int wcx = w->origin_cx + cx, wcz = w->origin_cz + cz;
    Rng r = { ((uint64_t)hash3(seed, wcx, 991, wcz) << 32) | 0xABCDEFu };
    static const int edge[6] = { 0, 1, 15, 16, 30, 31 };
    static const int ys[12] = { 0, 1, 30, 62, 63, 64, 65, 127, 128, 250, 254, 255 };
    for (int i = 0; i < 260; i++)
    {
        uint32_t v = rng_next(&r);
        int x = (v & 1) ? edge[(v >> 1) % 6] : (int)((v >> 1) & 31);
        int z = (v & 0x100) ? edge[(v >> 9) % 6] : (int)((v >> 9) & 31);
        int y = (v & 0x10000) ? ys[(v >> 17) % 12] : (int)((v >> 17) & 127);
        static const uint8_t pal[12] = { B_GLASS, B_ICE, B_MAGMA, B_LANTERN, B_LAVA, B_WATER, B_STONE, B_LEAVES, B_CACTUS, B_WOOD, B_SNOW, B_AIR };
        set_block(w, cx, cz, x, y, z, pal[(v >> 24) % 12]);
    }
}

typedef struct GenCtx { GcwWorld* w; int kind; uint32_t seed; int radius, falloff; } GenCtx;
static void gen_column(void* p, int c, int worker)
{
    (void)worker;
    GenCtx* g = (GenCtx*)p;
    GcwWorld* w = g->w;
    int cx = c % w->size_x, cz = c / w->size_x;
    memset(w->blocks + (size_t)c * WCY * CVOX, 0, (size_t)WCY * CVOX * 2);
    memset(w->sun + (size_t)c * WCY * CVOX, 0, (size_t)WCY * CVOX);
    memset(w->light + (size_t)c * WCY * CVOX, 0, (size_t)WCY * CVOX);
    switch (g->kind)
    {
        case GCW_FLAT: gen_flat_column(w, cx, cz, g->radius, g->falloff); break;
        case GCW_FOREST: gen_terrain_column(w, cx, cz, g->seed, T_FOREST, g->radius, g->falloff); break;
        case GCW_ISLANDS: gen_terrain_column(w, cx, cz, g->seed, T_ISLANDS, g->radius, g->falloff); break;
        case GCW_SNOW: gen_terrain_column(w, cx, cz, g->seed, T_SNOW, g->radius, g->falloff); break;
        case GCW_DESERT: gen_terrain_column(w, cx, cz, g->seed, T_DESERT, g->radius, g->falloff); break;
        case GCW_VOLCANIC: gen_terrain_column(w, cx, cz, g->seed, T_VOLCANIC, g->radius, g->falloff); break;
        case GCW_DUNGEON: gen_dungeon_column(w, cx, cz); break;
        case GCW_CHECKER: gen_checker_column(w, cx, cz, g->seed); break;
        case GCW_NOISE: gen_noise_column(w, cx, cz, g->seed); break;
        case GCW_STRESS: gen_stress_column(w, cx, cz); break;
        case GCW_MIXED: gen_mixed_column(w, cx, cz, g->seed, g->radius, g->falloff); break;
        default: break;
    }
}

void gcw_generate(GcwWorld* w, int kind, uint32_t seed, int radius, int falloff, int threads)
{
    int ncol = w->size_x * w->size_z;
    memset(w->surface, 0, (size_t)ncol * 1024);
    GenCtx g = { w, kind, seed, radius, falloff };
    par_for(ncol, threads, gen_column, &g);
}

/* ComputeSurface (Lighting.cpp:5-26): highest non-air y in 0..254, plus one; 0 for an empty column. */
static void surface_column(void* p, int c, int worker)
{
    (void)worker;
    GcwWorld* w = (GcwWorld*)p;
    int cx = c % w->size_x, cz = c / w->size_x;
    for (int z = 0; z < CH; z++) for (int x = 0; x < CH; x++)
    {
        int s = 0;
        for (int y = WH - 2; y >= 0; y--)
            if (w->blocks[voxel_index(w, cx, cz, x, y, z)] != B_AIR) { s = y + 1; break; }
        w->surface[(size_t)c * 1024 + z * CH + x] = (uint8_t)s;
    }
}

void gcw_compute_surface(GcwWorld* w, int threads)
{
    par_for(w->size_x * w->size_z, threads, surface_column, w);
}

/* ------------------------------------------------------------------------------------------ */
/* light flood fill                                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct GcwLightRules
{
    int32_t step[256];     /* lightStep (Block.h:95); INT_MAX = opaque (Block.cpp:65-68) */
    int32_t emitted[256];  /* lightEmitted */
} GcwLightRules;

void gcw_default_light_rules(GcwLightRules* r)
{
    for (int i = 0; i < 256; i++) { r->step[i] = INT_MAX; r->emitted[i] = 0; }
    r->step[B_AIR] = 1;                                   /* Block.cpp:146 */
    r->step[B_WATER] = 2;                                 /* :167 */
    r->step[B_ICE] = 2;                                   /* :211 */
    r->step[B_LANTERN] = 1; r->emitted[B_LANTERN] = 15;   /* :217-218 */
    r->step[B_MAGMA] = 1; r->emitted[B_MAGMA] = 10;       /* :235-236 */
    r->step[B_KILL_ZONE] = 1;                             /* :252 */
    r->step[B_LAVA] = 1; r->emitted[B_LAVA] = 10;         /* :261-262 */
    r->step[B_GLASS] = 1;                                 /* :272 */
}

typedef struct NodeQ { int32_t* v; size_t cap, rd, wr; } NodeQ; /* world-voxel coords (x,y,z) packed 3 ints */
static void q_init(NodeQ* q) { q->cap = 1u << 16; q->v = (int32_t*)malloc(q->cap * 3 * sizeof(int32_t)); q->rd = q->wr = 0; }
static void q_free(NodeQ* q) { free(q->v); }
static inline int q_empty(const NodeQ* q) { return q->rd == q->wr; }
static void q_push(NodeQ* q, int x, int y, int z)
{
    if (q->wr == q->cap)
    {
        if (q->rd > 0)
        {
            memmove(q->v, q->v + q->rd * 3, (q->wr - q->rd) * 3 * sizeof(int32_t));
            q->wr -= q->rd; q->rd = 0;
        }
        if (q->wr * 2 > q->cap) { q->cap *= 2; q->v = (int32_t*)realloc(q->v, q->cap * 3 * sizeof(int32_t)); }
    }
    int32_t* p = q->v + q->wr * 3; p[0] = x; p[1] = y; p[2] = z; q->wr++;
}
static inline void q_pop(NodeQ* q, int* x, int* y, int* z) { int32_t* p = q->v + q->rd * 3; *x = p[0]; *y = p[1]; *z = p[2]; q->rd++; }

static inline int64_t wvox(const GcwWorld* w, int lx, int ly, int lz) { return voxel_index(w, lx >> 5, lz >> 5, lx & 31, ly, lz & 31); }
static inline int surf_at(const GcwWorld* w, int lx, int lz) { return w->surface[col_index(w, lx >> 5, lz >> 5) * 1024 + (lz & 31) * CH + (lx & 31)]; }

/* GetSunlight (Lighting.cpp:69-79) */
static inline int sun_at(const GcwWorld* w, int lx, int ly, int lz, int64_t vi)
{
    if (ly >= surf_at(w, lx, lz)) return 15;
    int l = w->sun[vi];
    return l == 0 ? 1 : l;
}

static const int DIRS6[6][3] = { {0,1,0}, {0,-1,0}, {0,0,1}, {0,0,-1}, {-1,0,0}, {1,0,0} }; /* Utils.h:106-111 */

/* ScatterSunlight (Lighting.cpp:109-143) */
static void scatter_sun(GcwWorld* w, const GcwLightRules* rules, NodeQ* q)
{
    int wx_max = w->size_x * CH, wz_max = w->size_z * CH;
    while (!q_empty(q))
    {
        int x, y, z; q_pop(q, &x, &y, &z);
        int64_t vi = wvox(w, x, y, z);
        int step = rules->step[w->blocks[vi] & 255];
        if (step == INT_MAX) continue; /* sun - INT_MAX <= MIN_LIGHT */
        int l = sun_at(w, x, y, z, vi) - step;
        if (l <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            int nx = x + DIRS6[d][0], ny = y + DIRS6[d][1], nz = z + DIRS6[d][2];
            if (ny < 0 || ny >= WH) continue;
            if (nx < 0 || nz < 0 || nx >= wx_max || nz >= wz_max) continue; /* window edge (reference asserts) */
            int64_t ni = wvox(w, nx, ny, nz);
            if (rules->step[w->blocks[ni] & 255] == INT_MAX) continue;
            if (sun_at(w, nx, ny, nz, ni) < l) { w->sun[ni] = (uint8_t)l; q_push(q, nx, ny, nz); } /* SetMaxSunlight :81-93 */
        }
    }
}

/* ScatterBlockLight (Lighting.cpp:195-229) */
static void scatter_block(GcwWorld* w, const GcwLightRules* rules, NodeQ* q)
{
    int wx_max = w->size_x * CH, wz_max = w->size_z * CH;
    while (!q_empty(q))
    {
        int x, y, z; q_pop(q, &x, &y, &z);
        int64_t vi = wvox(w, x, y, z);
        int step = rules->step[w->blocks[vi] & 255];
        if (step == INT_MAX) continue;
        int l = (int)w->light[vi] - step;
        if (l <= 1) continue;
        for (int d = 0; d < 6; d++)
        {
            int nx = x + DIRS6[d][0], ny = y + DIRS6[d][1], nz = z + DIRS6[d][2];
            if (ny < 0 || ny >= WH) continue;
            if (nx < 0 || nz < 0 || nx >= wx_max || nz >= wz_max) continue;
            int64_t ni = wvox(w, nx, ny, nz);
            if (rules->step[w->blocks[ni] & 255] == INT_MAX) continue;
            if ((int)w->light[ni] < l) { w->light[ni] = (uint8_t)l; q_push(q, nx, ny, nz); } /* SetMaxBlockLight :95-107 */
        }
    }
}

/* SetLightNodes (Lighting.cpp:247-281) for one non-edge column. */
static void light_column(GcwWorld* w, const GcwLightRules* rules, int cx, int cz, NodeQ* sunq, NodeQ* blkq)
{
    sunq->rd = sunq->wr = 0;
    blkq->rd = blkq->wr = 0;
    for (int z = 0; z < CH; z++)
    {
        int lz = cz * CH + z;
        for (int x = 0; x < CH; x++)
        {
            int lx = cx * CH + x;
            int surface = surf_at(w, lx, lz);
            /* ComputeMaxY (Lighting.cpp:28-67): max with the four horizontal neighbours' surface, capped at 255 */
            int my = surface;
            int s;
            s = surf_at(w, lx, lz + 1); if (s > my) my = s;
            s = surf_at(w, lx, lz - 1); if (s > my) my = s;
            s = surf_at(w, lx - 1, lz); if (s > my) my = s;
            s = surf_at(w, lx + 1, lz); if (s > my) my = s;
            if (my > WH - 1) my = WH - 1;

            for (int y = 0; y < surface; y++)
            {
                int64_t vi = wvox(w, lx, y, lz);
                int e = rules->emitted[w->blocks[vi] & 255];
                if (e > 1) { w->light[vi] = (uint8_t)e; q_push(blkq, lx, y, lz); } /* CheckForBlockLight :231-245 */
            }
            for (int y = surface; y <= my; y++)
            {
                q_push(sunq, lx, y, lz);
                int64_t vi = wvox(w, lx, y, lz);
                int e = rules->emitted[w->blocks[vi] & 255];
                if (e > 1) { w->light[vi] = (uint8_t)e; q_push(blkq, lx, y, lz); }
            }
        }
    }
    scatter_sun(w, rules, sunq);
    scatter_block(w, rules, blkq);
}

/*
 * Light every non-edge column.  threads == 1: z-major then x, one column at a time — the order the oracle
 * harness uses for the reference.  threads > 1: columns three apart never touch the same bricks (light
 * travels < 32 voxels), so nine colour classes run in parallel; sunlight is an order-independent fixpoint.
 */
typedef struct LightCtx { GcwWorld* w; const GcwLightRules* rules; int ox, oz, nx; NodeQ* sunq; NodeQ* blkq; } LightCtx;
static void light_task(void* p, int i, int worker)
{
    LightCtx* c = (LightCtx*)p;
    int cx = 1 + c->ox + 3 * (i % c->nx), cz = 1 + c->oz + 3 * (i / c->nx);
    if (cx < c->w->size_x - 1 && cz < c->w->size_z - 1) light_column(c->w, c->rules, cx, cz, &c->sunq[worker], &c->blkq[worker]);
}

void gcw_light(GcwWorld* w, const GcwLightRules* rules, int threads)
{
    GcwLightRules def;
    if (!rules) { gcw_default_light_rules(&def); rules = &def; }
    if (threads <= 1)
    {
        NodeQ a, b; q_init(&a); q_init(&b);
        for (int cz = 1; cz < w->size_z - 1; cz++) for (int cx = 1; cx < w->size_x - 1; cx++) light_column(w, rules, cx, cz, &a, &b);
        q_free(&a); q_free(&b);
        return;
    }
    NodeQ* sunq = (NodeQ*)malloc(sizeof(NodeQ) * (size_t)threads);
    NodeQ* blkq = (NodeQ*)malloc(sizeof(NodeQ) * (size_t)threads);
    for (int t = 0; t < threads; t++) { q_init(&sunq[t]); q_init(&blkq[t]); }
    for (int colour = 0; colour < 9; colour++)
    {
        int ox = colour % 3, oz = colour / 3;
        int nx = (w->size_x - 2 - ox + 2) / 3, nz = (w->size_z - 2 - oz + 2) / 3;
        if (nx <= 0 || nz <= 0) continue;
        LightCtx c = { w, rules, ox, oz, nx, sunq, blkq };
        par_for(nx * nz, threads, light_task, &c);
    }
    for (int t = 0; t < threads; t++) { q_free(&sunq[t]); q_free(&blkq[t]); }
    free(sunq); free(blkq);
}

/* Convenience: generate + surface + light. */
void gcw_build(GcwWorld* w, int kind, uint32_t seed, int radius, int falloff, int threads)
{
    gcw_generate(w, kind, seed, radius, falloff, threads);
    gcw_compute_surface(w, threads);
    gcw_light(w, NULL, threads);
}

int gcw_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

// gcgen — terrain generation on the device (see gcgen_kernel.cuh).
//
// Three kernels per window:
//   heights  one thread per (column, z, x): island test, biome / flat / mountain noise, terrain height; column maximum
//   layers   one thread per (column, z, x) walking up to the column maximum: stone pockets (3-D fBm), then the biome's
//            layering (grass / dirt / water; snow / ice; sand; obsidian / lava)
//   trees    one thread per column: the column's own PRNG stream places trunks and leaf boxes (forest, islands) or cacti
//            (desert) in order (later ones overwrite earlier ones, as in the host producer)
// Every float operation is an explicitly rounded single operation (__fmul_rn / __fadd_rn / ...), never a fused
// multiply-add, in the order the host producer evaluates it (compiled with -ffp-contract=off), so both give the same
// heights; powf is taken through double precision.
#include "gcgen_kernel.cuh"

namespace gc
{
namespace
{
enum { B_AIR = 0, B_GRASS = 1, B_DIRT = 2, B_STONE = 3, B_SAND = 5, B_METAL_CRATE = 7, B_WATER = 8, B_WOOD = 9, B_LEAVES = 10,
       B_SNOW = 12, B_ICE = 13, B_CACTUS = 16, B_OBSIDIAN = 19, B_LAVA = 21 };   // Code/Block.h:11-38
enum { T_FOREST = 0, T_ISLANDS = 1, T_SNOW = 2, T_DESERT = 3, T_VOLCANIC = 4 };   // GenArgs::biome
__device__ __forceinline__ int sea_level(int biome) { return biome == T_FOREST || biome == T_DESERT ? 10 : (biome == T_VOLCANIC ? 12 : 20); }
constexpr int kIceBit = 0x4000;   // height scratch: "the sea surface freezes over this cell" (snow)

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t hash3(uint32_t seed, int32_t a, int32_t b, int32_t c)
{
    uint32_t h = seed ^ 0x9E3779B9u;
    h = mix32(h + (uint32_t)a * 0x85EBCA6Bu);
    h = mix32(h + (uint32_t)b * 0xC2B2AE35u);
    h = mix32(h + (uint32_t)c * 0x27D4EB2Fu);
    return h;
}
struct Rng { uint64_t s; };
__device__ __forceinline__ uint32_t rng_next(Rng& r)
{
    r.s += 0x9E3779B97F4A7C15ull;
    uint64_t z = r.s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}
__device__ __forceinline__ int rng_range(Rng& r, int lo, int hi) { return lo + (int)(rng_next(r) % (uint32_t)(hi + 1 - lo)); }

__device__ float grad2(uint32_t h, float x, float y)
{
    const float d = 0.70710678f;
    float gx, gy;
    switch (h & 7u)
    {
        case 0: gx = 1; gy = 0; break;
        case 1: gx = -1; gy = 0; break;
        case 2: gx = 0; gy = 1; break;
        case 3: gx = 0; gy = -1; break;
        case 4: gx = d; gy = d; break;
        case 5: gx = -d; gy = d; break;
        case 6: gx = d; gy = -d; break;
        default: gx = -d; gy = -d; break;
    }
    return add(mul(gx, x), mul(gy, y));
}

__device__ float corner2(uint32_t seed, int i, int j, float x, float y)
{
    // a corner out of reach contributes +0.0f: adding it leaves n as the host's skipped statement does
    float a = sub(sub(0.5f, mul(x, x)), mul(y, y));
    if (!(a > 0)) return 0.0f;
    a = mul(a, a);
    return mul(mul(a, a), grad2(hash3(seed, i, j, 0), x, y));
}

__device__ float simplex2(uint32_t seed, float x, float y)
{
    const float F2 = 0.36602540f, G2 = 0.21132487f;
    const float s = mul(add(x, y), F2);
    const int i = (int)floorf(add(x, s)), j = (int)floorf(add(y, s));
    const float t = mul((float)(i + j), G2);
    const float x0 = sub(x, sub((float)i, t)), y0 = sub(y, sub((float)j, t));
    const int i1 = x0 > y0, j1 = !i1;
    const float x1 = add(sub(x0, (float)i1), G2), y1 = add(sub(y0, (float)j1), G2);
    const float g22 = mul(2.0f, G2);
    const float x2 = add(sub(x0, 1.0f), g22), y2 = add(sub(y0, 1.0f), g22);
    float n = 0.0f;
    n = add(n, corner2(seed, i, j, x0, y0));
    n = add(n, corner2(seed, i + i1, j + j1, x1, y1));
    n = add(n, corner2(seed, i + 1, j + 1, x2, y2));
    n = mul(n, 70.0f);
    return n < -1.0f ? -1.0f : (n > 1.0f ? 1.0f : n);
}

__device__ float grad3(uint32_t h, float x, float y, float z)
{
    // G[h % 12] = (gx, gy, gz) with entries in {-1, 0, 1}: g0*x + g1*y + g2*z, zero terms included (x*0 = +-0 adds exactly)
    float gx, gy, gz;
    switch (h % 12u)
    {
        case 0: gx = 1; gy = 1; gz = 0; break;
        case 1: gx = -1; gy = 1; gz = 0; break;
        case 2: gx = 1; gy = -1; gz = 0; break;
        case 3: gx = -1; gy = -1; gz = 0; break;
        case 4: gx = 1; gy = 0; gz = 1; break;
        case 5: gx = -1; gy = 0; gz = 1; break;
        case 6: gx = 1; gy = 0; gz = -1; break;
        case 7: gx = -1; gy = 0; gz = -1; break;
        case 8: gx = 0; gy = 1; gz = 1; break;
        case 9: gx = 0; gy = -1; gz = 1; break;
        case 10: gx = 0; gy = 1; gz = -1; break;
        default: gx = 0; gy = -1; gz = -1; break;
    }
    return add(add(mul(gx, x), mul(gy, y)), mul(gz, z));
}

__device__ float corner3(uint32_t seed, int i, int j, int k, float x, float y, float z)
{
    float a = sub(sub(sub(0.6f, mul(x, x)), mul(y, y)), mul(z, z));
    if (!(a > 0)) return 0.0f;
    a = mul(a, a);
    return mul(mul(a, a), grad3(hash3(seed, i, j, k), x, y, z));
}

__device__ float simplex3(uint32_t seed, float x, float y, float z)
{
    const float F3 = __fdiv_rn(1.0f, 3.0f), G3 = __fdiv_rn(1.0f, 6.0f);
    const float s = mul(add(add(x, y), z), F3);
    const int i = (int)floorf(add(x, s)), j = (int)floorf(add(y, s)), k = (int)floorf(add(z, s));
    const float t = mul((float)(i + j + k), G3);
    const float x0 = sub(x, sub((float)i, t)), y0 = sub(y, sub((float)j, t)), z0 = sub(z, sub((float)k, t));
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
    const float g2 = mul(2.0f, G3), g3 = mul(3.0f, G3);
    const float x1 = add(sub(x0, (float)i1), G3), y1 = add(sub(y0, (float)j1), G3), z1 = add(sub(z0, (float)k1), G3);
    const float x2 = add(sub(x0, (float)i2), g2), y2 = add(sub(y0, (float)j2), g2), z2 = add(sub(z0, (float)k2), g2);
    const float x3 = add(sub(x0, 1.0f), g3), y3 = add(sub(y0, 1.0f), g3), z3 = add(sub(z0, 1.0f), g3);
    float n = 0.0f;
    n = add(n, corner3(seed, i, j, k, x0, y0, z0));
    n = add(n, corner3(seed, i + i1, j + j1, k + k1, x1, y1, z1));
    n = add(n, corner3(seed, i + i2, j + j2, k + k2, x2, y2, z2));
    n = add(n, corner3(seed, i + 1, j + 1, k + 1, x3, y3, z3));
    n = mul(n, 32.0f);
    return n < -1.0f ? -1.0f : (n > 1.0f ? 1.0f : n);
}

__device__ float fractal_bounding(int octaves)
{
    float amp = 0.5f, sum = 1.0f;
    for (int i = 1; i < octaves; i++) { sum = add(sum, amp); amp = mul(amp, 0.5f); }
    return __fdiv_rn(1.0f, sum);
}

__device__ float ridged2(uint32_t seed, float x, float y, int octaves)
{
    float sum = sub(1.0f, fabsf(simplex2(seed, x, y))), amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x = mul(x, 2.0f); y = mul(y, 2.0f); amp = mul(amp, 0.5f);
        sum = sub(sum, mul(sub(1.0f, fabsf(simplex2(seed + (uint32_t)i, x, y))), amp));
    }
    return sum;
}

__device__ float billow2(uint32_t seed, float x, float y, int octaves)
{
    float sum = sub(mul(fabsf(simplex2(seed, x, y)), 2.0f), 1.0f), amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x = mul(x, 2.0f); y = mul(y, 2.0f); amp = mul(amp, 0.5f);
        sum = add(sum, mul(sub(mul(fabsf(simplex2(seed + (uint32_t)i, x, y)), 2.0f), 1.0f), amp));
    }
    return mul(sum, fractal_bounding(octaves));
}

__device__ float fbm2(uint32_t seed, float x, float y, int octaves)
{
    float sum = simplex2(seed, x, y), amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x = mul(x, 2.0f); y = mul(y, 2.0f); amp = mul(amp, 0.5f);
        sum = add(sum, mul(simplex2(seed + (uint32_t)i, x, y), amp));
    }
    return mul(sum, fractal_bounding(octaves));
}

__device__ float fbm3(uint32_t seed, float x, float y, float z, int octaves)
{
    float sum = simplex3(seed, x, y, z), amp = 1.0f;
    for (int i = 1; i < octaves; i++)
    {
        x = mul(x, 2.0f); y = mul(y, 2.0f); z = mul(z, 2.0f); amp = mul(amp, 0.5f);
        sum = add(sum, mul(simplex3(seed + (uint32_t)i, x, y, z), amp));
    }
    return mul(sum, fractal_bounding(octaves));
}

__device__ __forceinline__ float scurve3(float a) { return mul(mul(a, a), sub(3.0f, mul(2.0f, a))); }

__device__ __forceinline__ void set_block(const GenArgs& a, int col, int x, int wy, int z, int b)
{
    if (wy < 0 || wy >= GC_WORLD_BLOCK_HEIGHT) return;
    a.blocks[((size_t)col * 4 + (wy >> 6)) * GC_CHUNK_SIZE_3 + x + 32 * ((wy & 63) + 64 * z)] = (uint16_t)b;
}

// ---- heights (Generation.cpp:83-130 / :228-275 as gcworld.c gen_terrain_column restates them) ----
__global__ void __launch_bounds__(256) gc_gen_height_kernel(GenArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.size_x * a.size_z * 1024;
    if (i >= n) return;
    const int col = i >> 10, z = (i >> 5) & 31, x = i & 31;
    const int cx = col % a.size_x, cz = col / a.size_x;
    const int sea = sea_level(a.biome);
    const int wx = (a.origin_cx + cx) * 32 + x, wz = (a.origin_cz + cz) * 32 + z;
    int h = 0, m = sea, ice = 0;
    // IsWithinIsland (Generation.cpp:66-81)
    const int v = (int)sqrt((double)wx * wx + (double)wz * wz);
    if (v < a.radius)
    {
        float p = 1.0f;
        if (v > a.falloff) p = sub(1.0f, __fdiv_rn((float)(v - a.falloff), (float)(a.radius - a.falloff)));
        const float bx = (float)wx, bz = (float)wz;
        float terrain;
        if (a.biome == T_DESERT)
        {
            // one billow field of three octaves (Generation.cpp:566-569, :585-586)
            const float base = mul(add(billow2(a.seed + 202u, mul(mul(bx, 0.05f), 0.25f), mul(mul(bz, 0.05f), 0.25f), 3), 1.0f), 0.5f);
            terrain = add(mul(mul(base, 0.2f), 30.0f), 20.0f);
        }
        else if (a.biome == T_SNOW)
        {
            // sea floor and land blended over the biome band -0.4 .. 0, whose upper part freezes the sea (Generation.cpp:431-479)
            const float biome = fbm2(a.seed + 101u, mul(bx, 0.005f), mul(bz, 0.005f), 3);
            float floor_ = mul(add(billow2(a.seed + 202u, mul(mul(bx, 0.025f), 0.5f), mul(mul(bz, 0.025f), 0.5f), 4), 1.0f), 0.5f);
            floor_ = add(mul(mul(floor_, 0.2f), 20.0f), 5.0f);
            float land = mul(add(ridged2(a.seed + 303u, mul(mul(bx, 0.015f), 0.5f), mul(mul(bz, 0.015f), 0.5f), 4), 1.0f), 0.5f);
            land = add(mul(mul(land, 0.2f), 10.0f), 20.0f);
            const float lower = -0.4f, upper = 0.0f;
            if (biome < lower) terrain = floor_;
            else if (biome > upper) terrain = land;
            else
            {
                if (biome >= -0.25f && biome <= 0.0f) ice = kIceBit;
                terrain = add(floor_, mul(scurve3(__fdiv_rn(sub(biome, lower), sub(upper, lower))), sub(land, floor_)));
            }
        }
        else
        {
            const float mscale = a.biome == T_FOREST ? 30.0f : 60.0f, mbase = a.biome == T_FOREST ? 10.0f : 20.0f;
            const float bfreq = a.biome == T_VOLCANIC ? 0.005f : 0.01f;
            const float biome = simplex2(a.seed + 101u, mul(bx, bfreq), mul(bz, bfreq));
            float flat = mul(add(billow2(a.seed + 202u, mul(mul(bx, 0.025f), 0.5f), mul(mul(bz, 0.025f), 0.5f), 4), 1.0f), 0.5f);
            flat = add(mul(mul(flat, 0.2f), 30.0f), 10.0f);
            float mount = mul(add(ridged2(a.seed + 303u, mul(mul(bx, 0.015f), 0.5f), mul(mul(bz, 0.015f), 0.5f), 4), 1.0f), 0.5f);
            if (mount < 0.0f) mount = 0.0f;
            mount = add(mul((float)pow((double)mount, 3.5), mscale), mbase);
            if (biome < 0.0f) terrain = flat;
            else if (biome > 0.6f) terrain = mount;
            else terrain = add(flat, mul(scurve3(__fdiv_rn(biome, 0.6f)), sub(mount, flat)));
        }
        h = (int)mul(terrain, p);
        if (h > GC_WORLD_BLOCK_HEIGHT - 12) h = GC_WORLD_BLOCK_HEIGHT - 12;
        if (h > m) m = h;
    }
    a.height[i] = (int16_t)(h | ice);
    // ComputeSurface's entry for this cell (Lighting.cpp:5-17) before trees: the layering fills 0..h and water up to sea level
    a.surface[i] = (uint8_t)(max(h, sea) + 1);
    // column maximum of m: warp first, one atomic per warp (a warp is one z-row of one column)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(a.max_y + col, m);
}

// ---- layers (Generation.cpp:132-200): stone pockets, grass on top, dirt below, water up to sea level ----
// A CTA owns eight z-rows of one column (a warp = one row of 32 x) and walks up to the column's highest layer.
__global__ void __launch_bounds__(256) gc_gen_layers_kernel(GenArgs a)
{
    const int col = blockIdx.x >> 2;
    const int z = (blockIdx.x & 3) * 8 + (threadIdx.x >> 5), x = threadIdx.x & 31;
    const int cx = col % a.size_x, cz = col / a.size_x;
    const int sea = sea_level(a.biome);
    const int hraw = a.height[col * 1024 + z * 32 + x];
    const int h = hraw & (kIceBit - 1);
    const bool ice = (hraw & kIceBit) != 0;
    const int top = a.max_y[col];
    const int wz = (a.origin_cz + cz) * 32 + z;
    const float fz = mul(mul((float)wz, 0.015f), 0.2f);
    for (int wy = 0; wy <= top; wy++)
    {
        int b = B_AIR;
        bool stone = false;
        if (a.biome != T_DESERT && wy <= h - 4)
        {
            // what the reference's code does, not what it means: its 3-D noise set has maxY + 1 layers per x and is read with a
            // stride of maxY (Generation.cpp:16-18 against :157-161), so cell (x, wY) gets the sample of (x', y') with
            // x' * (maxY + 1) + y' = x * maxY + wY
            const int q = x * top + wy, sx = q / (top + 1), sy = q % (top + 1);
            const float fx = mul(mul((float)((a.origin_cx + cx) * 32 + sx), 0.015f), 0.2f);
            const float comp = mul(add(fbm3(a.seed + 404u, fx, mul(mul((float)sy, 0.015f), 0.2f), fz, 2), 1.0f), 0.5f);
            stone = comp <= 0.2f;
        }
        if (stone) b = B_STONE;
        else if (a.biome == T_SNOW)
        {
            // (Generation.cpp:525-539) the sea-level layer is tested before "below the surface": a column taller than the sea
            // holds water (or ice) at y = 20 inside its dirt
            if (wy == h) b = B_SNOW;
            else if (wy > h && wy < sea) b = B_WATER;
            else if (wy == sea) b = ice ? B_ICE : B_WATER;
            else if (wy < h) b = B_DIRT;
        }
        else if (a.biome == T_DESERT) b = wy <= h ? B_SAND : (wy <= sea ? B_WATER : B_AIR);          // (:610-613)
        else if (a.biome == T_VOLCANIC) b = wy <= h ? B_OBSIDIAN : (wy <= sea ? B_LAVA : B_AIR);     // (:750-753)
        else if (wy == h) b = B_GRASS;
        else if (wy > h && wy <= sea) b = B_WATER;
        else if (wy < h) b = B_DIRT;
        if (b != B_AIR) set_block(a, col, x, wy, z, b);
    }
}

__device__ void box(const GenArgs& a, int col, int x0, int y0, int z0, int x1, int y1, int z1, int b)
{
    for (int z = z0; z <= z1; z++)
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) set_block(a, col, x, y, z, b);
}

// surface = highest non-AIR y in 0..254, plus one (Lighting.cpp:5-17): cells under a box whose top layer is `top`
__device__ void raise_surface(const GenArgs& a, int col, int x0, int z0, int x1, int z1, int top)
{
    if (top > GC_WORLD_BLOCK_HEIGHT - 2) top = GC_WORLD_BLOCK_HEIGHT - 2;
    if (top < 0) return;
    for (int z = z0; z <= z1; z++)
        for (int x = x0; x <= x1; x++)
        {
            uint8_t* s = a.surface + (size_t)col * 1024 + z * 32 + x;
            if (*s < top + 1) *s = (uint8_t)(top + 1);
        }
}

// ---- trees (Generation.cpp:50-64, :204-215): the column's own PRNG stream, in order ----
__global__ void gc_gen_trees_kernel(GenArgs a)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= a.size_x * a.size_z) return;
    const int cx = col % a.size_x, cz = col / a.size_x;
    const int wcx = a.origin_cx + cx, wcz = a.origin_cz + cz;
    const int sea = sea_level(a.biome);
    Rng r = { ((uint64_t)hash3(a.seed, wcx, wcz, 77) << 32) | hash3(a.seed, wcz, wcx, 78) };
    if (a.biome == T_DESERT)
    {
        // zero to two cacti anywhere in the column, two to four blocks on top of the surface entry — under water too (Generation.cpp:630-641)
        const int cacti = rng_range(r, 0, 2);
        for (int i = 0; i < cacti; i++)
        {
            const int rx = rng_range(r, 0, 31);
            const int rz = rng_range(r, 0, 31);
            const int ch = rng_range(r, 2, 4);
            const int s = a.height[col * 1024 + rz * 32 + rx];
            for (int j = 1; j <= ch; j++) set_block(a, col, rx, s + j, rz, B_CACTUS);
            raise_surface(a, col, rx, rz, rx, rz, s + ch);
        }
        return;
    }
    if (a.biome != T_FOREST && a.biome != T_ISLANDS) return;   // snow and volcanic columns carry no decoration
    const int trees = a.biome == T_ISLANDS ? rng_range(r, 1, 3) : rng_range(r, 3, 6);
    for (int i = 0; i < trees; i++)
    {
        const int rx = rng_range(r, 3, 32 - 4);
        const int rz = rng_range(r, 3, 32 - 4);
        const int s = a.height[col * 1024 + rz * 32 + rx];
        if (s <= sea) continue;
        // CreateTree: trunk then five leaf boxes
        const int by = s + 1;
        const int th = rng_range(r, 3, 5);
        for (int j = 0; j < th; j++) set_block(a, col, rx, by + j, rz, B_WOOD);
        const int sy = by + th;
        box(a, col, rx - 1, sy, rz - 1, rx + 1, sy, rz + 1, B_LEAVES);
        box(a, col, rx - 2, sy + 1, rz - 1, rx + 2, sy + 2, rz + 1, B_LEAVES);
        box(a, col, rx - 1, sy + 1, rz + 2, rx + 1, sy + 2, rz + 2, B_LEAVES);
        box(a, col, rx - 1, sy + 1, rz - 2, rx + 1, sy + 2, rz - 2, B_LEAVES);
        box(a, col, rx - 1, sy + 3, rz - 1, rx + 1, sy + 3, rz + 1, B_LEAVES);
        // the surface entries under the crown (the trunk's cell is under its top box); a tree tops out below y = 254
        raise_surface(a, col, rx - 1, rz - 1, rx + 1, rz + 1, sy + 3);
        raise_surface(a, col, rx - 2, rz - 1, rx + 2, rz + 1, sy + 2);
        raise_surface(a, col, rx - 1, rz + 2, rx + 1, rz + 2, sy + 2);
        raise_surface(a, col, rx - 1, rz - 2, rx + 1, rz - 2, sy + 2);
    }
}

// GenerateDungeon (Dungeon.cpp:17-31): in the column's lowest chunk a stone floor and four stone walls eleven blocks high along
// the column's border.  One thread per (column, z, x).
__global__ void __launch_bounds__(256) gc_gen_dungeon_kernel(GenArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.size_x * a.size_z * 1024) return;
    const int col = i >> 10, z = (i >> 5) & 31, x = i & 31;
    const bool wall = x == 0 || x == 31 || z == 0 || z == 31;
    const int top = wall ? 10 : 0;
    for (int y = 0; y <= top; y++) set_block(a, col, x, y, z, B_STONE);
    a.surface[i] = (uint8_t)(top + 1);
}

// ---- the reference's three generators that use no noise: reproduced exactly (pinned against the reference's own code) ----
// IsWithinIsland (Generation.cpp:66-81)
__device__ __forceinline__ bool within_island(const GenArgs& a, int wx, int wz, float& p)
{
    const int v = (int)sqrt((double)(wx * wx) + (double)(wz * wz));          // Square() is integer arithmetic
    if (v >= a.radius) return false;
    p = 1.0f;
    if (v > a.falloff) p = sub(1.0f, __fdiv_rn((float)(v - a.falloff), (float)(a.radius - a.falloff)));
    return true;
}

// GenerateFlatTerrain (Generation.cpp:772-836): sea level 12; inside the island grass at (int)(12 * p) with dirt below and
// water up to sea level, outside it height 0 (one grass layer under the sea)
__global__ void __launch_bounds__(256) gc_gen_flat_kernel(GenArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.size_x * a.size_z * 1024 * 13;
    if (i >= n) return;
    const int col = i / (1024 * 13), r = i % (1024 * 13);
    const int wy = r >> 10, z = (r >> 5) & 31, x = r & 31;
    const int cx = col % a.size_x, cz = col / a.size_x;
    const int sea = 12;
    float p;
    int height = 0;
    if (within_island(a, (a.origin_cx + cx) * 32 + x, (a.origin_cz + cz) * 32 + z, p)) height = (int)mul((float)sea, p);
    int b = B_AIR;
    if (wy == height) b = B_GRASS;
    else if (wy > height && wy <= sea) b = B_WATER;
    else if (wy < height) b = B_DIRT;
    if (b != B_AIR) set_block(a, col, x, wy, z, b);
    if (wy == 0) a.surface[(size_t)col * 1024 + z * 32 + x] = (uint8_t)(max(height, sea) + 1);
}

// GenerateGridTerrain (Generation.cpp:371-410): in each of the column's four chunks, metal crates on the even lattice
// inside the radius, a 3 x 3 bundle of pillars and a mast in the middle.  The mast's last crate is stored at chunk-local
// y = 64: BlockIndex (World.cpp:278-281) carries it to (16, 0, 17) of the same chunk, which is where it is put here.
__global__ void __launch_bounds__(256) gc_gen_grid_kernel(GenArgs a)
{
    const int brick = blockIdx.x;                             // one CTA per chunk
    const int col = brick >> 2;
    const int cx = col % a.size_x, cz = col / a.size_x;
    uint16_t* chunk = a.blocks + (size_t)brick * GC_CHUNK_SIZE_3;
    const int t = threadIdx.x;                                // 16 x 16 lattice points
    const int x = (t & 15) * 2, z = (t >> 4) * 2;
    const int wx = (a.origin_cx + cx) * 32 + x, wz = (a.origin_cz + cz) * 32 + z;
    if ((int)sqrt((double)(wx * wx) + (double)(wz * wz)) < a.radius)
        for (int y = 2; y < 14; y += 2) chunk[x + 32 * (y + 64 * z)] = B_METAL_CRATE;
    __syncthreads();
    if (t < 9)
    {
        const int px = (t % 3) * 2 - 2 + 16, pz = (t / 3) * 2 - 2 + 16;
        for (int y = 14; y < 40; y += 2) chunk[px + 32 * (y + 64 * pz)] = B_METAL_CRATE;
    }
    if (t == 9)
        for (int y = 40; y <= 64; y += 2) chunk[16 + 32 * (y + 64 * 16)] = B_METAL_CRATE;
}

// GenerateVoidTerrain (Generation.cpp:838-852): a grass plate at y = 40 in the column at the world origin, nothing else
__global__ void gc_gen_void_kernel(GenArgs a)
{
    const int cx = -a.origin_cx, cz = -a.origin_cz;
    if (cx < 0 || cx >= a.size_x || cz < 0 || cz >= a.size_z) return;
    const int col = cz * a.size_x + cx;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) set_block(a, col, i & 31, 40, i >> 5, B_GRASS);
}
}  // namespace

cudaError_t launch_generate_terrain(const GenArgs& a, cudaStream_t s)
{
    const int n_cols = a.size_x * a.size_z;
    cudaError_t e = cudaMemsetAsync(a.max_y, 0, sizeof(int32_t) * n_cols, s);
    if (e != cudaSuccess) return e;
    gc_gen_height_kernel<<<(n_cols * 1024 + 255) / 256, 256, 0, s>>>(a);
    gc_gen_layers_kernel<<<n_cols * 4, 256, 0, s>>>(a);
    gc_gen_trees_kernel<<<(n_cols + 63) / 64, 64, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_generate_dungeon(const GenArgs& a, cudaStream_t s)
{
    const int n = a.size_x * a.size_z * 1024;
    gc_gen_dungeon_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_generate_flat(const GenArgs& a, cudaStream_t s)
{
    const int n = a.size_x * a.size_z * 1024 * 13;
    gc_gen_flat_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_generate_grid(const GenArgs& a, cudaStream_t s)
{
    gc_gen_grid_kernel<<<a.size_x * a.size_z * 4, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_generate_void(const GenArgs& a, cudaStream_t s)
{
    gc_gen_void_kernel<<<1, 256, 0, s>>>(a);
    return cudaGetLastError();
}
}  // namespace gc

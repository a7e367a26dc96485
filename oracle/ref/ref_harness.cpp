// TEST INFRASTRUCTURE ONLY — the "gold" oracle: the reference's own mesher, compiled verbatim.
//
// This translation unit #includes the reference's hot-path headers and sources where they
// lie under /root/reference/Code (never copied into this repository) in the order of the
// reference's unity build (Code/Main.cpp:92-143), minus the Win32/GL/imgui/audio bodies,
// behind the stand-ins in shim_platform.h / shim_glm.h.  It then exposes a small C ABI so
// tests (ctypes) and bench.py's `--impl reference` / `cpu_baseline` legs can drive
//     generate (flat only) -> ComputeSurface -> SetLightNodes -> BuildChunkAsync
// exactly as the reference does.  Built by oracle/Makefile into oracle/_ref/libgcref.so.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.

#include "shim_platform.h"

// Stand-ins for headers whose bodies need Win32 / imgui (Code/Main.cpp:92-100).
#define DEBUG_SERVICES 0
#define TIMED_FUNCTION
#define BEGIN_BLOCK(ID)
#define END_BLOCK(ID)
#define DRAW_CHUNK_OUTLINE(chunk)
#define TRACK_MESH
#define RESET_TRACKED_MESHES

#include "Utils.h"

template <typename T>
struct ObjectPool
{
    T* Get() { return new T(); }
    void Return(T* item) { delete item; }
};

#include "Containers.h"
// The reference draws tree and cactus positions with rand() (Random.h:32-35, RandRange = min + rand() % n).  Tests may plug a
// stream in (gcref_set_noise): rand() then returns draw % 1248 — 1248 = lcm of every n the generators use (3, 4, 26, 32), so
// rand() % n equals draw % n and the reference's own placement code reproduces the host producer's per-column stream.  Without a
// stream it is the C library's rand().
static int (*const g_libc_rand)(void) = rand;
static uint32_t (*g_rng_next)(uint64_t*) = nullptr;
static uint64_t g_rng_state = 0;
static int gcref_rand() { return g_rng_next ? (int)(g_rng_next(&g_rng_state) % 1248u) : g_libc_rand(); }
#define rand gcref_rand
#include "Random.h"

struct UI { int unused; };
struct AudioEngine { int unused; };

#include "AssetBuilder.h"
#include "Assets.h"
#include "Input.h"
#include "Mesh.h"
#include "Block.h"
#include "Particles.h"
#include "Environment.h"
#include "Generation.h"
#include "World.h"
#include "WorldIO.h"
static void SaveWorld(GameState*, World*);
#include "Renderer.h"
#include "Lighting.h"
#include "WorldRender.h"
#include "Simulation.h"
#include "Async.h"
#include "Commands.h"
#include "GameState.h"

// Link-time stubs for symbols the kept sources name but the mesher never reaches.
static Sound GetSound(GameState*, SoundID) { return Sound{}; }
static void PlaySound(Sound) {}
static void IgnoreCollideFunc(GameState*, World*, vec3&, vec3, Block) {}
static void BounceCollideFunc(GameState*, World*, vec3&, vec3, Block) {}
static void DefaultCollideFunc(GameState*, World*, vec3&, vec3, Block) {}
static void OverTimeDamageCollideFunc(GameState*, World*, vec3&, vec3, Block) {}
static void KillCollideFunc(GameState*, World*, vec3&, vec3, Block) {}
static bool OverlapsBlock(Player*, int, int, int) { return false; }
static void QueueAsync(GameState*, AsyncFunc, World*, void*, AsyncCallback) {}
static FrustumVisibility TestFrustum(Camera*, vec3, vec3) { return FRUSTUM_VISIBLE; }
static void DeleteRegions(World*) {}
static void SpawnPlayer(GameState*, World*, Player*, Rectf) {}
static void TeleportPlayer(GameState*, WorldLocation) {}
static void CreateBiomes(GameState*, World*) {}
static void ResetEnvironment(GameState*, World*) {}
static void GenerateDungeonTerrain(World*, ChunkGroup*) {}
static void MoveCamera(Camera*, vec3) {}
static void UpdateCameraVectors(Camera*) {}
static void UpdateViewMatrix(Camera*) {}
static void GetCameraPlanes(Camera*) {}
static void DeleteDirectory(char*) {}

#include "Mesh.cpp"
#include "Block.cpp"
#include "World.cpp"
// WorldIO.cpp:29 returns `false` from a function returning Region* (accepted by MSVC); g++ needs a null pointer constant there
#define false 0
#include "WorldIO.cpp"
#undef false
#include "Lighting.cpp"
#include "WorldRender.cpp"
#include "Generation.cpp"
#include "Dungeon.cpp"

#include <chrono>
#include <thread>
#include <pthread.h>
#include <sched.h>

// ---------------------------------------------------------------------------------------
// C ABI of the gold oracle.
// ---------------------------------------------------------------------------------------
struct GcRefWorld
{
    GameState* state;
    World* world;
    int size;
};

static ChunkGroup* RefGroup(GcRefWorld* w, int cx, int cz) { return w->world->groups[cz * w->size + cx]; }

static void RefCopyOut(MeshData* md, uint8_t* verts, uint16_t* const* idx, int32_t* counts)
{
    counts[0] = md->vertCount;
    if (verts) memcpy(verts, md->vertices, (size_t)md->vertCount * sizeof(VertexInfo));
    for (int t = 0; t < MESH_TYPE_COUNT; t++)
    {
        MeshIndexData* d = md->indices[t];
        counts[1 + t] = d ? d->count : 0;
        // counts[6 + t]: was the index array allocated at all (WorldRender.cpp:454-459)?
        counts[6 + t] = d ? 1 : 0;
        if (d && idx && idx[t]) memcpy(idx[t], d->data, (size_t)d->count * sizeof(uint16_t));
    }
}

static void RefResetMeshData(MeshData* md)
{
    md->vertCount = 0;
    for (int t = 0; t < MESH_TYPE_COUNT; t++)
    {
        if (md->indices[t]) { delete md->indices[t]; md->indices[t] = nullptr; }
    }
}

extern "C" {

int gcref_sizeof(int what)
{
    switch (what)
    {
        case 0: return (int)sizeof(VertexInfo);
        case 1: return (int)sizeof(MeshData);
        case 2: return (int)sizeof(MeshIndexData);
        case 3: return (int)sizeof(Chunk);
        case 4: return (int)sizeof(ChunkGroup);
        case 5: return (int)sizeof(BlockData);
        case 6: return (int)offsetof(VertexInfo, pos);
        case 7: return (int)offsetof(VertexInfo, uv);
        case 8: return (int)offsetof(VertexInfo, color);
        case 9: return (int)offsetof(VertexInfo, alpha);
        case 10: return (int)offsetof(VertexInfo, reserved);
        case 11: return MAX_VERTICES;
        case 12: return MAX_INDICES;
        case 13: return BLOCK_COUNT;
    }
    return -1;
}

// Column window like World::groups (World.cpp:153-160): size x size columns, all present.
GcRefWorld* gcref_world_create(int size, int seed, int radius, int falloffRadius)
{
    GcRefWorld* w = new GcRefWorld;
    w->state = new GameState();
    w->world = new World();
    w->size = size;
    World* world = w->world;
    world->size = size;
    world->loadRange = size / 2;
    world->totalGroups = size * size;
    world->groups = (ChunkGroup**)calloc((size_t)world->totalGroups, sizeof(ChunkGroup*));
    world->properties.seed = seed;
    world->properties.radius = radius;
    world->properties.biome = BIOME_FLAT;
    world->falloffRadius = falloffRadius;
    CreateBlockData(w->state, world->blockData);

    for (int cz = 0; cz < size; cz++)
    {
        for (int cx = 0; cx < size; cx++)
        {
            // As CreateChunkGroup (World.cpp:624-643) without the async load.
            ChunkGroup* group = (ChunkGroup*)calloc(1, sizeof(ChunkGroup));
            group->pos = ivec3(cx, 0, cz);
            group->active = true;
            for (int i = 0; i < WORLD_CHUNK_HEIGHT; i++)
            {
                Chunk* chunk = group->chunks + i;
                chunk->lcPos = ivec3(cx, i, cz);
                chunk->lwPos = LChunkToLWorldP(chunk->lcPos);
                chunk->group = group;
            }
            group->state = GROUP_LOADED;
            world->groups[cz * size + cx] = group;
        }
    }
    return w;
}

void gcref_world_destroy(GcRefWorld* w)
{
    for (int i = 0; i < w->world->totalGroups; i++) free(w->world->groups[i]);
    free(w->world->groups);
    delete w->world;
    delete w->state;
    delete w;
}

// Dump the fields of World::blockData the mesher reads (Block.cpp:5-68).
// out[b*12 + 0..5] textures, 6 meshType, 7 cull, 8 alpha, 9 invisible, 10 opaque, 11 lightEmitted
void gcref_block_table(GcRefWorld* w, int32_t* out)
{
    for (int b = 0; b < BLOCK_COUNT; b++)
    {
        BlockData& d = w->world->blockData[b];
        for (int f = 0; f < 6; f++) out[b * 12 + f] = d.textures[f];
        out[b * 12 + 6] = (int)d.meshType;
        out[b * 12 + 7] = d.cull;
        out[b * 12 + 8] = d.alpha;
        out[b * 12 + 9] = d.invisible ? 1 : 0;
        out[b * 12 + 10] = IsOpaque(w->world, (Block)b) ? 1 : 0;
        out[b * 12 + 11] = d.lightEmitted;
    }
}

void gcref_block_light_steps(GcRefWorld* w, int32_t* out)
{
    for (int b = 0; b < BLOCK_COUNT; b++) out[b] = w->world->blockData[b].lightStep;
}

void gcref_block_light_emitted(GcRefWorld* w, int32_t* out)
{
    for (int b = 0; b < BLOCK_COUNT; b++) out[b] = w->world->blockData[b].lightEmitted;
}

void gcref_set_chunk(GcRefWorld* w, int cx, int cy, int cz, const uint16_t* blocks, const uint8_t* sun, const uint8_t* light)
{
    Chunk* c = RefGroup(w, cx, cz)->chunks + cy;
    if (blocks) memcpy(c->blocks, blocks, sizeof(c->blocks));
    if (sun) memcpy(c->sunlight, sun, sizeof(c->sunlight));
    if (light) memcpy(c->blockLight, light, sizeof(c->blockLight));
}

void gcref_get_chunk(GcRefWorld* w, int cx, int cy, int cz, uint16_t* blocks, uint8_t* sun, uint8_t* light)
{
    Chunk* c = RefGroup(w, cx, cz)->chunks + cy;
    if (blocks) memcpy(blocks, c->blocks, sizeof(c->blocks));
    if (sun) memcpy(sun, c->sunlight, sizeof(c->sunlight));
    if (light) memcpy(light, c->blockLight, sizeof(c->blockLight));
}

void gcref_set_surface(GcRefWorld* w, int cx, int cz, const uint8_t* surface) { memcpy(RefGroup(w, cx, cz)->surface, surface, CHUNK_SIZE_2); }
void gcref_get_surface(GcRefWorld* w, int cx, int cz, uint8_t* surface) { memcpy(surface, RefGroup(w, cx, cz)->surface, CHUNK_SIZE_2); }

// GenerateFlatTerrain (Generation.cpp:772-836) on every column; no noise involved when radius is huge.
void gcref_generate_flat(GcRefWorld* w)
{
    for (int i = 0; i < w->world->totalGroups; i++) GenerateFlatTerrain(w->world, w->world->groups[i]);
}

// The reference's generators that need no FastNoiseSIMD, on every column of the window (group->pos = window position):
// 0 GenerateFlatTerrain (Generation.cpp:772-836), 3 GenerateGridTerrain (:371-410), 4 GenerateVoidTerrain (:838-852),
// 8 GenerateDungeon (Dungeon.cpp:17-31).
int gcref_generate(GcRefWorld* w, int kind)
{
    for (int i = 0; i < w->world->totalGroups; i++)
    {
        ChunkGroup* g = w->world->groups[i];
        if (kind == 0) GenerateFlatTerrain(w->world, g);
        else if (kind == 3) GenerateGridTerrain(w->world, g);
        else if (kind == 4) GenerateVoidTerrain(w->world, g);
        else if (kind == 8) GenerateDungeon(w->world, g);
        else return -1;
    }
    return 0;
}

// ---- the noise terrains over a plugged-in noise field ----
// `fn`: float (*)(uint32_t seed, int field, int octaves, float x, float y, float z) — gamecraft_b200/worldgen's gcw_noise.  The
// generators set one seed for all their fields; the host producer keeps its fields apart by an offset on the seed (biome + 101,
// billow + 202, ridged + 303, 3-D fBm + 404), which is put back here from what the generator asked for.
typedef float (*GcwNoiseFn)(uint32_t, int, int, float, float, float);
static GcwNoiseFn g_gcw_noise = nullptr;
static float NoiseBridge(int seed, int type, int ftype, int octaves, int dims, float x, float y, float z)
{
    const bool fractal = type == (int)FastNoiseSIMD::SimplexFractal;
    if (dims == 3) return g_gcw_noise((uint32_t)seed + 404u, 4, octaves, x, y, z);
    if (!fractal) return g_gcw_noise((uint32_t)seed + 101u, 0, 1, x, z, 0.0f);                         // plain simplex: the biome field
    if (ftype == (int)FastNoiseSIMD::Billow) return g_gcw_noise((uint32_t)seed + 202u, 1, octaves, x, z, 0.0f);
    if (ftype == (int)FastNoiseSIMD::RigidMulti) return g_gcw_noise((uint32_t)seed + 303u, 2, octaves, x, z, 0.0f);
    return g_gcw_noise((uint32_t)seed + 101u, 3, octaves, x, z, 0.0f);                                  // 2-D fBm: the snow biome field
}
typedef uint64_t (*GcwColumnRngFn)(uint32_t, int32_t, int32_t);
static GcwColumnRngFn g_column_rng = nullptr;
void gcref_set_noise(void* fn, void* column_rng, void* rng_next)
{
    g_gcw_noise = (GcwNoiseFn)fn;
    FastNoiseSIMD::hook() = fn ? NoiseBridge : nullptr;
    g_column_rng = (GcwColumnRngFn)column_rng;
    g_rng_next = (uint32_t (*)(uint64_t*))rng_next;
}
void gcref_set_seed(GcRefWorld* w, int seed) { w->world->properties.seed = seed; }

// The reference's five noise terrains (Generation.cpp:83-226, :228-369, :412-556, :558-645, :647-770) on every column, over the
// plugged-in noise: 1 forest, 2 islands, 5 snow, 6 desert, 7 volcanic (the numbering of GC_GEN_*).  Trees and cacti come from
// rand() as in the reference — the plugged-in per-column stream when there is one (see gcref_rand above).
int gcref_generate_noise_terrain(GcRefWorld* w, int kind)
{
    if (!g_gcw_noise) return -2;
    for (int i = 0; i < w->world->totalGroups; i++)
    {
        ChunkGroup* g = w->world->groups[i];
        if (g_column_rng) g_rng_state = g_column_rng((uint32_t)w->world->properties.seed, g->pos.x, g->pos.z);   // the column's own stream
        if (kind == 1) GenerateForestTerrain(w->world, g);
        else if (kind == 2) GenerateIslandsTerrain(w->world, g);
        else if (kind == 5) GenerateSnowTerrain(w->world, g);
        else if (kind == 6) GenerateDesertTerrain(w->world, g);
        else if (kind == 7) GenerateVolcanicTerrain(w->world, g);
        else return -1;
    }
    return 0;
}

// ComputeSurface (Lighting.cpp:19-26) on one column / all columns.
void gcref_compute_surface(GcRefWorld* w, int cx, int cz)
{
    ChunkGroup* g = RefGroup(w, cx, cz);
    memset(g->surface, 0, sizeof(g->surface));
    ComputeSurface(g);
}

void gcref_compute_surface_all(GcRefWorld* w)
{
    for (int cz = 0; cz < w->size; cz++) for (int cx = 0; cx < w->size; cx++) gcref_compute_surface(w, cx, cz);
}

// SetLightNodes (Lighting.cpp:247-281) for all non-edge columns, z-major then x (a fixed order;
// the reference's order is thread-timing dependent, WorldRender.cpp:255-260).
void gcref_light_all(GcRefWorld* w)
{
    for (int cz = 1; cz < w->size - 1; cz++)
    {
        for (int cx = 1; cx < w->size - 1; cx++)
        {
            ChunkGroup* g = RefGroup(w, cx, cz);
            Queue<ivec3> sunNodes(MAX_LIGHT_NODES);
            Queue<ivec3> lightNodes(MAX_LIGHT_NODES);
            SetLightNodes(w->world, g, sunNodes, lightNodes);
            g->state = GROUP_PREPROCESSED;
        }
    }
}

// SaveGroup (WorldIO.cpp:245-307) on one column: the four chunks' RLE records as the reference stores them in
// region->chunks[] ([offset][items][(count, block) ...]).  Records are concatenated into `out` (capacity `cap` words);
// sizes[cy] = words of chunk cy's record.  Returns the total, or -1 when `out` is too small.
int gcref_rle_save_group(GcRefWorld* w, int cx, int cz, uint16_t* out, int cap, int32_t* sizes)
{
    World* world = w->world;
    ChunkGroup* group = RefGroup(w, cx, cz);
    Region* region = GetOrLoadRegion(world, ChunkToRegionP(group->pos));   // no file behind it: an empty region from the pool
    for (int y = 0; y < WORLD_CHUNK_HEIGHT; y++) group->chunks[y].modified = true;
    SaveGroup(w->state, world, group);
    int total = 0;
    for (int y = 0; y < WORLD_CHUNK_HEIGHT; y++)
    {
        ivec3 local = ivec3(group->pos.x & REGION_MASK, y, group->pos.z & REGION_MASK);
        List<uint16_t>& store = region->chunks[RegionIndex(local)];
        sizes[y] = store.size;
        if (total + store.size > cap) return -1;
        memcpy(out + total, store.items, (size_t)store.size * sizeof(uint16_t));
        total += store.size;
    }
    return total;
}

// A window whose column (0, 0) is world column (ox, oz): the groups' world positions as World::CreateChunkGroup would have
// set them.  Only the region functions below look at group->pos (the mesher walks World::groups by window index).
void gcref_world_set_origin(GcRefWorld* w, int ox, int oz)
{
    for (int cz = 0; cz < w->size; cz++)
        for (int cx = 0; cx < w->size; cx++) RefGroup(w, cx, cz)->pos = ivec3(cx + ox, 0, cz + oz);
}

// What LoadRegionFile (WorldIO.cpp:15-80) does with one record read from a region file: word 0 (`position`) picks the
// list of the region — here the region of column (cx, cz) — that receives [position][items][items words].
// Returns the position, or -1 for a position outside the region.
int gcref_rle_store_record(GcRefWorld* w, int cx, int cz, const uint16_t* rec, int words)
{
    World* world = w->world;
    ChunkGroup* group = RefGroup(w, cx, cz);
    Region* region = GetOrLoadRegion(world, ChunkToRegionP(group->pos));
    if (words < 2) return -1;
    const uint16_t position = rec[0], items = rec[1];
    if (position >= REGION_SIZE_3) return -1;
    List<uint16_t>* chunk = region->chunks + position;
    chunk->Clear();
    chunk->Reserve(words);
    chunk->Add(position);
    chunk->Add(items);
    for (int i = 2; i < words; i++) chunk->Add(rec[i]);
    region->hasData = true;
    return position;
}

// Empties every list of the region of column (cx, cz).
void gcref_rle_clear_region(GcRefWorld* w, int cx, int cz)
{
    Region* region = GetOrLoadRegion(w->world, ChunkToRegionP(RefGroup(w, cx, cz)->pos));
    for (int i = 0; i < REGION_SIZE_3; i++) region->chunks[i].Clear();
}

// LoadGroupFromDisk (WorldIO.cpp:167-221): the records currently stored for column (cx, cz) decoded into its chunks'
// block arrays (which the caller may have cleared in between).  Returns 1 when all four chunks had data.
int gcref_rle_load_group(GcRefWorld* w, int cx, int cz)
{
    return LoadGroupFromDisk(w->world, RefGroup(w, cx, cz)) ? 1 : 0;
}

// BuildChunkAsync (WorldRender.cpp:6-27) on one chunk with a fresh MeshData.
// verts: >= 40000*12 bytes (or null), idx[t]: >= 60000 u16 each (or null),
// counts[11]: vertCount, count[5], allocated[5].  Returns chunk->totalVertices.
int gcref_build_chunk(GcRefWorld* w, int cx, int cy, int cz, uint8_t* verts, uint16_t* const* idx, int32_t* counts)
{
    Chunk* chunk = RefGroup(w, cx, cz)->chunks + cy;
    MeshData* md = new MeshData();
    md->vertCount = 0;
    chunk->meshData = md;
    BuildChunkAsync(w->state, w->world, chunk);
    RefCopyOut(md, verts, idx, counts);
    RefResetMeshData(md);
    delete md;
    chunk->meshData = nullptr;
    return chunk->totalVertices;
}

// ChunkOverflowed (WorldRender.cpp:412-439).
int gcref_chunk_overflowed(GcRefWorld* w, int cx, int cy, int cz, int x, int y, int z, int totalVertices)
{
    Chunk* chunk = RefGroup(w, cx, cz)->chunks + cy;
    chunk->totalVertices = totalVertices;
    return ChunkOverflowed(w->world, chunk, x, y, z) ? 1 : 0;
}

// SetBlock (World.cpp:548-582) as the interactive caller runs it: ChunkOverflowed pre-check, RecomputeLight,
// FlagChunkForUpdate.  The player-overlap test is stubbed to "no overlap".  Returns 1 when the stored block differs from
// what it was before the call.
int gcref_set_block(GcRefWorld* w, int lwX, int lwY, int lwZ, int block)
{
    if (lwY < 0 || lwY >= WORLD_BLOCK_HEIGHT) { SetBlock(w->world, lwX, lwY, lwZ, (Block)block); return 0; }
    const Block before = GetBlock(w->world, lwX, lwY, lwZ);
    SetBlock(w->world, lwX, lwY, lwZ, (Block)block);
    return GetBlock(w->world, lwX, lwY, lwZ) != before ? 1 : 0;
}

// chunk->state / totalVertices / pendingUpdate as the edit path reads and writes them (World.h:74-99)
void gcref_set_chunk_built(GcRefWorld* w, int cx, int cy, int cz, int built, int totalVertices)
{
    Chunk* chunk = RefGroup(w, cx, cz)->chunks + cy;
    chunk->state = built ? CHUNK_BUILT : CHUNK_LOADED_DATA;
    chunk->totalVertices = totalVertices;
}

// out[(cz * size + cx) * 4 + cy] = pendingUpdate; clear != 0 resets the flags afterwards
void gcref_pending_updates(GcRefWorld* w, uint8_t* out, int clear)
{
    for (int cz = 0; cz < w->size; cz++)
        for (int cx = 0; cx < w->size; cx++)
            for (int cy = 0; cy < WORLD_CHUNK_HEIGHT; cy++)
            {
                Chunk* chunk = RefGroup(w, cx, cz)->chunks + cy;
                out[(cz * w->size + cx) * 4 + cy] = chunk->pendingUpdate ? 1 : 0;
                if (clear) { chunk->pendingUpdate = false; chunk->modified = false; }
            }
}

// Multi-threaded mesh-build timing: the reference's one-chunk-per-job queue (Async.cpp:5-35) restated as
// an atomic counter; each worker owns a private MeshData (the pool hands one per job, WorldRender.cpp:61-67).
// coords: n x (cx,cy,cz).  Returns seconds; *quadsOut = total quads.
double gcref_bench_build(GcRefWorld* w, const int32_t* coords, int n, int threads, int64_t* quadsOut)
{
    std::atomic<int> next(0);
    std::atomic<int64_t> quads(0);
    std::vector<MeshData*> mds((size_t)threads);
    for (int t = 0; t < threads; t++) { mds[t] = new MeshData(); }

    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
    {
        pool.emplace_back([&, t]()
        {
            // worker t runs on the t-th CPU this process may use (round robin): a timed run does not depend on where the
            // scheduler happens to put, or move, its threads
            {
                cpu_set_t all, one;
                if (!getenv("GCREF_NO_PIN") && sched_getaffinity(0, sizeof(all), &all) == 0 && CPU_COUNT(&all) > 0)
                {
                    int want = t % CPU_COUNT(&all);
                    for (int c = 0; c < CPU_SETSIZE; c++)
                        if (CPU_ISSET(c, &all) && want-- == 0)
                        {
                            CPU_ZERO(&one); CPU_SET(c, &one);
                            pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
                            break;
                        }
                }
            }
            MeshData* md = mds[t];
            int64_t local = 0;
            for (;;)
            {
                int i = next.fetch_add(1);
                if (i >= n) break;
                Chunk* chunk = RefGroup(w, coords[i * 3 + 0], coords[i * 3 + 2])->chunks + coords[i * 3 + 1];
                md->vertCount = 0;
                chunk->meshData = md;
                BuildChunkAsync(w->state, w->world, chunk);
                local += md->vertCount / 4;
                RefResetMeshData(md);
                chunk->meshData = nullptr;
            }
            quads += local;
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();

    for (int t = 0; t < threads; t++) delete mds[t];
    if (quadsOut) *quadsOut = quads.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Exhaustive parity check of a mesher batch against the reference's own BuildChunkAsync (multi-threaded): desc = one 64-byte
// GcChunkOut per chunk (include/gcmesh.h), vertices / indices = the batch arenas.  result[i]: 0 identical; bit 0 counts,
// bit 1 vertex bytes, bit 2 index bytes, bit 3 flags / segment outside the arena.  Chunks over the reference's cap (an
// assert in the reference, WorldRender.cpp:558) cannot be built by it: the caller must not pass them (flags != 0 is reported).
struct RefChunkOut { uint32_t quad_base, vert_count, index_count[MESH_TYPE_COUNT], index_offset[MESH_TYPE_COUNT], quads, flags, reserved[2]; };
int64_t gcref_check_batch(GcRefWorld* w, const int32_t* coords, int n, const void* descv, const void* verticesv, const void* indicesv,
                          uint64_t arenaQuads, int threads, uint8_t* result, int64_t* quadsChecked)
{
    const RefChunkOut* desc = (const RefChunkOut*)descv;
    const uint8_t* vertices = (const uint8_t*)verticesv;
    const uint16_t* indices = (const uint16_t*)indicesv;
    if (threads < 1) threads = 1;
    std::atomic<int> next(0);
    std::atomic<int64_t> bad(0), quads(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
    {
        pool.emplace_back([&]()
        {
            MeshData* md = new MeshData();
            for (;;)
            {
                int i = next.fetch_add(1);
                if (i >= n) break;
                const RefChunkOut& d = desc[i];
                uint8_t r = 0;
                if (d.flags != 0) r |= 8;
                else
                {
                    Chunk* chunk = RefGroup(w, coords[i * 3 + 0], coords[i * 3 + 2])->chunks + coords[i * 3 + 1];
                    md->vertCount = 0;
                    chunk->meshData = md;
                    BuildChunkAsync(w->state, w->world, chunk);
                    if (d.vert_count != (uint32_t)md->vertCount || d.quads * 4 != (uint32_t)md->vertCount) r |= 1;
                    for (int m = 0; m < MESH_TYPE_COUNT; m++)
                        if (d.index_count[m] != (uint32_t)(md->indices[m] ? md->indices[m]->count : 0)) r |= 1;
                    if (!r && md->vertCount)
                    {
                        if ((uint64_t)d.quad_base + d.quads > arenaQuads) r |= 8;
                        else
                        {
                            if (memcmp(vertices + (size_t)d.quad_base * 48, md->vertices, (size_t)md->vertCount * sizeof(VertexInfo))) r |= 2;
                            for (int m = 0; m < MESH_TYPE_COUNT; m++)
                                if (md->indices[m] && md->indices[m]->count &&
                                    memcmp(indices + (size_t)d.quad_base * 6 + d.index_offset[m], md->indices[m]->data, (size_t)md->indices[m]->count * 2)) r |= 4;
                        }
                    }
                    quads += md->vertCount / 4;
                    RefResetMeshData(md);
                    chunk->meshData = nullptr;
                }
                if (result) result[i] = r;
                if (r) bad++;
            }
            delete md;
        });
    }
    for (auto& th : pool) th.join();
    if (quadsChecked) *quadsChecked = quads.load();
    return bad.load();
}

}  // extern "C"

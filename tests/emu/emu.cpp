// TEST INFRASTRUCTURE — sequential host walk over the kernel's phases.
//
// Runs the exact per-lane functions of gamecraft_b200/csrc/mesh_core.cuh (compiled for the host) in the
// order the sm_100a kernel runs them, on ordinary memory instead of shared memory.  This checks the
// bit-level logic (class planes, face masks, ranks, light gathering, vertex packing) against the CPU oracle
// on the build box, where there is no GPU.  It is NOT a fallback: nothing in the product links it.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../gamecraft_b200/csrc/mesh_core.cuh"
#include "../../include/gcmesh.h"

using namespace gc;

namespace
{
struct EmuWorld
{
    int size_x, size_z;
    const uint16_t* blocks;
    const uint8_t* sun;
    const uint8_t* light;
    const uint8_t* surface;
};

const FaceDesc kFaces[6] = GC_FACE_TABLE;
}

extern "C" int gcemu_build_chunk(int size_x, int size_z, const uint16_t* blocks, const uint8_t* sun, const uint8_t* light,
                                 const uint8_t* surface, const GcBlockInfo* table, int ntab, int kill_zone,
                                 int cx, int cy, int cz, uint8_t* verts_out, uint16_t* idx_out, GcChunkOut* out)
{
    EmuWorld w = { size_x, size_z, blocks, sun, light, surface };
    std::vector<LutEntry> lut(256);
    std::vector<uint8_t> cls8(256);
    std::vector<MatEntry> mat(256);
    for (int i = 0; i < 256; i++)
    {
        const GcBlockInfo& b = table[i < ntab ? i : 0];
        int emits = (i != 0) && !b.invisible && i < ntab;
        uint32_t c = class_byte(b.cull, b.opaque, b.mesh_type, emits);
        cls8[i] = (uint8_t)c;
        lut[i] = spread_class(c);
        mat[i].lo = b.textures[0] | (b.textures[1] << 8) | (b.textures[2] << 16) | ((uint32_t)b.textures[3] << 24);
        mat[i].hi = b.textures[4] | (b.textures[5] << 8) | (b.alpha << 16);
    }

    std::vector<uint32_t> planes_store((size_t)kPlaneWords + 4, 0xDEADBEEFu);
    struct AlignedView { uint32_t* p; uint32_t* data() { return p; } } planes{ (uint32_t*)(((uintptr_t)planes_store.data() + 15) & ~(uintptr_t)15) };
    std::vector<uint16_t> hx(kHaloRows, 0xFFFF);
    uint8_t* hx_b = (uint8_t*)hx.data();

    // ---- phase 1: centre-x columns of every row (class planes + light), like the kernel's class / light warps ----
    std::vector<uint32_t> bvm(2 * kBvWords, 0);  // [side][layer]: 64-bit masks of rows whose x = 0 / x = 31 voxel emits
    for (int tz = -1; tz <= kTZ; tz++)
    {
        int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
        int zz = tz & 31;
        int64_t col = (int64_t)(cz + dcz) * w.size_x + cx;
        for (int ty = -1; ty <= kTY; ty++)
        {
            int wy = cy * 64 + ty;
            int hr = hrow(ty, tz);
            if (wy < 0 || wy >= 256)
            {
                uint32_t out[8];
                p1_class_row_const(wy < 0 ? cls8[kill_zone] : cls8[0], out);
                store_row_planes(planes.data(), hr, out);
                continue;
            }
            int64_t base = (col * 4 + (wy >> 6)) * 65536 + 32 * ((wy & 63) + 64 * zz);
            const uint32_t* bw = (const uint32_t*)(w.blocks + base);
            const bool centre = ty >= 0 && ty < kTY && tz >= 0 && tz < kTZ;
            {
                // the kernel reads the four octets rotated by a lane-dependent amount and un-rotates the plane words
                const int rot = (ty + tz + 2) & 3;
                Ids8 q[4];
                for (int k = 0; k < 4; k++)
                {
                    const uint32_t* o = bw + 4 * ((k + rot) & 3);
                    q[k].x = o[0]; q[k].y = o[1]; q[k].z = o[2]; q[k].w = o[3];
                }
                uint32_t out[8];
                p1_class_row(q, row_uniformity(q), lut.data(), out);
                for (int k = 0; k < 8; k++) out[k] = rotl_bytes(out[k], rot_selector(rot));
                store_row_planes(planes.data(), hr, out);
                uint32_t vis = centre ? row_emit_mask(out) : 0u;
                const int sl = tz + 1;
                if (vis & 1u) bvm[(ty + 1) * 2 + (sl >> 5)] |= 1u << (sl & 31);
                if (vis >> 31) bvm[kBvWords + (ty + 1) * 2 + (sl >> 5)] |= 1u << (sl & 31);
            }
        }
    }
    // ---- phase 1x: only the x-halo cells something will read; the others stay poisoned (0xFFFF) ----
    for (int hr = 0; hr < kHaloRows; hr++)
    {
        int ty = hr / kRowsZ - 1, tz = hr % kRowsZ - 1;
        int wy = cy * 64 + ty;
        for (int side = 0; side < 2; side++)
        {
            if (wy < 0) { p1_xhalo_cell(cls8[kill_zone], hx_b, hr, side); continue; }
            if (wy >= 256) { p1_xhalo_cell(cls8[0], hx_b, hr, side); continue; }
            if (!xhalo_needed(bvm.data() + side * kBvWords, ty, tz)) continue;
            int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
            int zz = tz & 31, xx = side ? 0 : 31;
            int64_t col = (int64_t)(cz + dcz) * w.size_x + (cx + (side ? 1 : -1));
            int64_t vox = (col * 4 + (wy >> 6)) * 65536 + xx + 32 * ((wy & 63) + 64 * zz);
            p1_xhalo_cell(cls8[w.blocks[vox] & 255u], hx_b, hr, side);
        }
    }

    // ---- phase 2 ----
    std::vector<uint8_t> rowCnt(kCenterRows);
    std::vector<uint32_t> rowTyp(kCenterRows);
    uint32_t total = 0, ttot[kMeshTypes] = { 0, 0, 0, 0, 0 };
    for (int R = 0; R < kCenterRows; R++)
    {
        RowFaces rf;
        p2_row_faces(planes.data(), hx.data(), R >> 5, R & 31, rf);
        int c; uint32_t t4;
        p2_row_counts(rf, c, t4);
        rowCnt[R] = (uint8_t)c; rowTyp[R] = t4;
        total += c;
        for (int t = 1; t < kMeshTypes; t++) ttot[t] += (t4 >> (8 * (t - 1))) & 255u;
    }
    ttot[0] = total - ttot[1] - ttot[2] - ttot[3] - ttot[4];

    memset(out, 0, sizeof(*out));
    out->quads = total;
    if (total > (uint32_t)kMaxQuads) { out->flags = GC_CHUNK_OVERFLOW; return 0; }
    out->vert_count = total * 4;
    uint32_t type_off[kMeshTypes];
    uint32_t acc = 0;
    for (int t = 0; t < kMeshTypes; t++) { type_off[t] = acc; out->index_offset[t] = acc * 6; out->index_count[t] = ttot[t] * 6; acc += ttot[t]; }

    // ---- phase 3a ----
    std::vector<uint32_t> list(total + 1);
    std::vector<uint16_t> tlist(total + 1);
    uint32_t rank = 0, trank[kMeshTypes] = { 0, 0, 0, 0, 0 };
    for (int R = 0; R < kCenterRows; R++)
    {
        RowFaces rf;
        p2_row_faces(planes.data(), hx.data(), R >> 5, R & 31, rf);
        p3a_row(rf, R >> 5, R & 31, rank, trank, list.data(), tlist.data());
        uint32_t t4 = rowTyp[R], c = rowCnt[R], nz = 0;
        for (int t = 1; t < kMeshTypes; t++) { uint32_t k = (t4 >> (8 * (t - 1))) & 255u; trank[t] += k; nz += k; }
        trank[0] += c - nz;
        rank += c;
    }

    // ---- phase 3b: the surface rows the kernel's fetch warp stages, then one gather per quad ----
    std::vector<uint8_t> surf(kSurfTileBytes);
    for (int tz = -1; tz <= kTZ; tz++)
        for (int x = -1; x <= kTX; x++)
        {
            int dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0), dcx = x < 0 ? -1 : (x > 31 ? 1 : 0);
            surf[(tz + 1) * kSurfStride + x + 1] = w.surface[((int64_t)(cz + dcz) * w.size_x + cx + dcx) * 1024 + (tz & 31) * 32 + (x & 31)];
        }
    LightGather lg;
    lg.sun = w.sun; lg.light = w.light; lg.surf = surf.data();
    lg.own = (((int64_t)cz * w.size_x + cx) * 4 + cy) * 65536;
    lg.size_x = w.size_x; lg.cx = cx; lg.cy = cy; lg.cz = cz;
    const uint16_t* cb = w.blocks + (((int64_t)cz * w.size_x + cx) * 4 + cy) * 65536;
    for (uint32_t i = 0; i < total; i++)
    {
        uint32_t e = list[i];
        int x = (e >> 3) & 31, tz = (e >> 8) & 31, ty = (e >> 13) & 63;
        uint32_t id = cb[x + 32 * (ty + 64 * tz)] & 255u;
        uint32_t v[12], iw[3];
        p3b_quad(e, make_face_plan(kFaces[e & 7]), lg, planes.data() + 4 * kHaloRows, hx.data(), mat[id], v);
        memcpy(verts_out + (size_t)i * 48, v, 48);
        quad_index_words(i, iw);
        memcpy((uint32_t*)idx_out + (size_t)(type_off[(e >> 19) & 7] + tlist[i]) * 3, iw, 12);
    }
    return 0;
}


// ---- the relight replay (gamecraft_b200/csrc/relight_core.cuh) on the host: the device kernel runs the same functions ----
#include "../../gamecraft_b200/csrc/relight_core.cuh"

extern "C" void gcemu_recompute_light(int size_x, int size_z, const uint16_t* blocks, uint8_t* sun, uint8_t* light, uint8_t* surface,
                                      const uint8_t* step, const uint8_t* emitted, const int32_t* vert_table, uint8_t* dirty, int x, int y, int z)
{
    gc::RelightWorld w = { blocks, sun, light, surface, step, emitted, vert_table, dirty, size_x, size_z };
    std::vector<uint64_t> a((size_t)gc::kRelightQueueNodes), b((size_t)gc::kRelightQueueNodes);
    gc::RelightQueue qa = { a.data(), 0, 0, 0 }, qb = { b.data(), 0, 0, 0 };
    gc::rl_recompute_light(w, x, y, z, qa, qb);
}

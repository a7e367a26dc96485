// gcmesh core: the per-lane building blocks of the chunk-mesh kernel.
//
// Everything here is a small __host__ __device__ function over plain pointers, so the SAME code runs
//   * inside the sm_100a kernel (gcmesh_kernel.cu), on shared-memory tiles, and
//   * inside tests/emu (a sequential host walk over the same phases) where it is checked bit-for-bit
//     against the CPU oracle before any GPU time is spent.
// The kernel adds only the orchestration: TMA staging, mbarriers, warp roles, scans, arena reservation.
//
// On-chip form of one chunk + 1-voxel halo (tile cell (x,ty,tz), x in -1..32, ty in -1..64, tz in -1..32):
//   row index      hr(ty,tz) = (ty+1)*34 + (tz+1)                      2244 rows, z fastest
//   planes         two arrays of 16-byte records, one record per row: G[hr] = {plane 0..3}, M[hr] = {plane 4..7} (M follows G),
//                  so a row's four cull planes (or opaque + type planes) move as one 128-bit shared-memory access;
//                  plane word k: bit x = class bit k of cell (x,ty,tz), x in 0..31
//                  k: 0..3 = cull >= 1..4 (thermometer code of CullValue, Block.h:40-47)
//                     4    = opaque (Block.cpp:65-68)
//                     5..7 = mesh type bits (Mesh.h:8-16); 7 (all ones) = never emits (AIR / invisible)
//   hx[hr]         uint16: low byte = the 8 class bits of cell x = -1, high byte = of cell x = 32
// Light is NOT staged: only cells in front of a drawn face are ever read (VertexLight, WorldRender.cpp:329-383), so
// each quad gathers its 3x3 light cells straight from the sunlight / blockLight arrays (LightGather below), as
//   (sun' << 4) | blockLight with GetSunlight's rules applied (Lighting.cpp:69-79);
//   0x00 below the world, 0xFF above it (GetFinalLight, WorldRender.cpp:316-319)
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GC_HD __host__ __device__ __forceinline__
#else
#define GC_HD static inline
#endif

namespace gc
{
constexpr int kTX = 32, kTY = 64, kTZ = 32;
constexpr int kRowsY = kTY + 2, kRowsZ = kTZ + 2;
constexpr int kHaloRows = kRowsY * kRowsZ;   // 2244
constexpr int kCenterRows = kTY * kTZ;       // 2048
constexpr int kPlanes = 8;
constexpr int kMaxQuads = 10000;             // MAX_VERTICES / 4, Mesh.h:5
constexpr int kMeshTypes = 5;
constexpr int kNoEmitType = 7;

enum { P_G1 = 0, P_G2 = 1, P_G3 = 2, P_G4 = 3, P_OPAQUE = 4, P_M0 = 5, P_M1 = 6, P_M2 = 7 };

GC_HD int hrow(int ty, int tz) { return (ty + 1) * kRowsZ + (tz + 1); }

struct alignas(16) Planes4 { uint32_t x, y, z, w; };
constexpr int kPlaneWords = kPlanes * kHaloRows;                     // the whole tile, in 32-bit words
GC_HD const Planes4* planes_g(const uint32_t* planes) { return reinterpret_cast<const Planes4*>(planes); }
GC_HD const Planes4* planes_m(const uint32_t* planes) { return reinterpret_cast<const Planes4*>(planes) + kHaloRows; }
GC_HD void store_row_planes(uint32_t* planes, int hr, const uint32_t o[8])
{
    reinterpret_cast<Planes4*>(planes)[hr] = Planes4{ o[0], o[1], o[2], o[3] };
    reinterpret_cast<Planes4*>(planes)[kHaloRows + hr] = Planes4{ o[4], o[5], o[6], o[7] };
}

// ---- small portable intrinsics ----
GC_HD int popc(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
GC_HD int lowbit(uint32_t v)  // index of lowest set bit, v != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
// per byte: a >= b ? 0xFF : 0x00
GC_HD uint32_t cmpge4(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vcmpgeu4(a, b);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 255u) >= ((b >> (8 * i)) & 255u)) r |= 255u << (8 * i);
    return r;
#endif
}
// per byte max
GC_HD uint32_t max4(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vmaxu4(a, b);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
    {
        uint32_t x = (a >> (8 * i)) & 255u, y = (b >> (8 * i)) & 255u;
        r |= (x > y ? x : y) << (8 * i);
    }
    return r;
#endif
}

// PRMT: result byte i = byte (sel nibble i) of the 8-byte pool {a bytes 0..3, b bytes 4..7}
GC_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t pool = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7))) & 255u) << (8 * i);
    return r;
#endif
}
// 4x4 byte transpose: out[j] byte i = in[i] byte j
GC_HD void transpose4x4(const uint32_t in[4], uint32_t out[4])
{
    const uint32_t t0 = byte_perm(in[0], in[1], 0x5140), t1 = byte_perm(in[2], in[3], 0x5140);
    const uint32_t t2 = byte_perm(in[0], in[1], 0x7362), t3 = byte_perm(in[2], in[3], 0x7362);
    out[0] = byte_perm(t0, t1, 0x5410); out[1] = byte_perm(t0, t1, 0x7632);
    out[2] = byte_perm(t2, t3, 0x5410); out[3] = byte_perm(t2, t3, 0x7632);
}
// rotate left by r bytes (r in 0..3) as one PRMT: sel = rot_selector(r), computed once per lane
GC_HD uint32_t rot_selector(int r) { return (0x32103210u >> (4 * ((4 - r) & 3))) & 0xFFFFu; }
GC_HD uint32_t rotl_bytes(uint32_t v, uint32_t sel) { return byte_perm(v, v, sel); }

// ---- block table in kernel form ----
struct LutEntry { uint32_t lo, hi; };  // class bit k at bit position 8k (lo: k 0..3, hi: k 4..7)

// class byte of one block type: the 8 class bits packed (see file header)
GC_HD uint32_t class_byte(int cull, int opaque, int mesh_type, int emits)
{
    uint32_t c = 0;
    for (int k = 1; k <= 4; k++)
        if (cull >= k) c |= 1u << (k - 1);
    if (opaque) c |= 1u << P_OPAQUE;
    c |= (uint32_t)(emits ? mesh_type : kNoEmitType) << P_M0;
    return c;
}
GC_HD LutEntry spread_class(uint32_t cls)
{
    LutEntry e;
    e.lo = (cls & 1u) | ((cls & 2u) << 7) | ((cls & 4u) << 14) | ((cls & 8u) << 21);
    uint32_t h = cls >> 4;
    e.hi = (h & 1u) | ((h & 2u) << 7) | ((h & 4u) << 14) | ((h & 8u) << 21);
    return e;
}

// =====================================================================================================
// Phase 1: raw voxels -> on-chip form
// =====================================================================================================

// Eight consecutive voxels -> their 8 class bits, one byte per plane: lo = planes 0..3, hi = planes 4..7
// (byte k bit i = class bit k of voxel i).  w0..w3 hold the 8 uint16 block ids.
GC_HD void p1_class_octet(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, const LutEntry* lut, uint32_t& lo, uint32_t& hi)
{
    const bool uniform = (w0 == w1) && (w2 == w3) && (w0 == w2) && ((w0 >> 16) == (w0 & 0xFFFFu));
    if (uniform)
    {
        const LutEntry e = lut[w0 & 255u];
        lo = e.lo * 255u;
        hi = e.hi * 255u;
    }
    else
    {
        const LutEntry a0 = lut[w0 & 255u], b0 = lut[(w0 >> 16) & 255u], a1 = lut[w1 & 255u], b1 = lut[(w1 >> 16) & 255u];
        const LutEntry a2 = lut[w2 & 255u], b2 = lut[(w2 >> 16) & 255u], a3 = lut[w3 & 255u], b3 = lut[(w3 >> 16) & 255u];
        lo = a0.lo + (b0.lo << 1) + (a1.lo << 2) + (b1.lo << 3) + (a2.lo << 4) + (b2.lo << 5) + (a3.lo << 6) + (b3.lo << 7);
        hi = a0.hi + (b0.hi << 1) + (a1.hi << 2) + (b1.hi << 3) + (a2.hi << 4) + (b2.hi << 5) + (a3.hi << 6) + (b3.hi << 7);
    }
}

// One row of 32 voxels -> the row's word of each of the 8 class planes.  q[k] = octet k (x = 8k .. 8k+7) as 8 uint16.
// The caller may pass the octets rotated (q[k] = octet (k + r) & 3, to read shared memory without bank conflicts)
// and undo it with rotl_bytes(out[j], r).
struct Ids8 { uint32_t x, y, z, w; };
GC_HD bool octet_uniform(const Ids8& q) { return (q.x == q.y) && (q.z == q.w) && (q.x == q.z) && ((q.x >> 16) == (q.x & 0xFFFFu)); }
// How uniform is a row?  2: one block id throughout (air, deep stone: the bulk of a world), 1: every octet holds a
// single id, 0: mixed.  Reads all sixteen words whatever the answer.
GC_HD int row_uniformity(const Ids8 q[4])
{
    const uint32_t w = q[0].x;
    const uint32_t diff = (q[0].y ^ w) | (q[0].z ^ w) | (q[0].w ^ w) | (q[1].x ^ w) | (q[1].y ^ w) | (q[1].z ^ w) | (q[1].w ^ w)
                        | (q[2].x ^ w) | (q[2].y ^ w) | (q[2].z ^ w) | (q[2].w ^ w) | (q[3].x ^ w) | (q[3].y ^ w) | (q[3].z ^ w) | (q[3].w ^ w);
    if ((diff | ((w >> 16) ^ (w & 0xFFFFu))) == 0) return 2;
    return (octet_uniform(q[0]) & octet_uniform(q[1]) & octet_uniform(q[2]) & octet_uniform(q[3])) ? 1 : 0;
}
GC_HD void p1_class_row(const Ids8 q[4], int uniformity, const LutEntry* lut, uint32_t out[8])
{
    if (uniformity == 2)
    {
        // all-zero / all-one plane words: nothing to transpose (or to un-rotate)
        const LutEntry e = lut[q[0].x & 255u];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            out[k] = 0u - ((e.lo >> (8 * k)) & 1u);
            out[4 + k] = 0u - ((e.hi >> (8 * k)) & 1u);
        }
        return;
    }
    uint32_t lo[4], hi[4];
    if (uniformity == 1)
    {
        // four single-block octets (water over sand ...): four independent table lookups, no branches between them
        const LutEntry e0 = lut[q[0].x & 255u], e1 = lut[q[1].x & 255u], e2 = lut[q[2].x & 255u], e3 = lut[q[3].x & 255u];
        lo[0] = e0.lo * 255u; lo[1] = e1.lo * 255u; lo[2] = e2.lo * 255u; lo[3] = e3.lo * 255u;
        hi[0] = e0.hi * 255u; hi[1] = e1.hi * 255u; hi[2] = e2.hi * 255u; hi[3] = e3.hi * 255u;
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 4; k++) p1_class_octet(q[k].x, q[k].y, q[k].z, q[k].w, lut, lo[k], hi[k]);
    }
    transpose4x4(lo, out);
    transpose4x4(hi, out + 4);
}
GC_HD void p1_class_row_const(uint32_t cls, uint32_t out[8])
{
#pragma unroll
    for (int k = 0; k < 8; k++) out[k] = ((cls >> k) & 1u) ? 0xFFFFFFFFu : 0u;
}
// emitting voxels of a row (mesh type != 7) from its type planes
GC_HD uint32_t row_emit_mask(const uint32_t out[8]) { return ~(out[P_M0] & out[P_M1] & out[P_M2]); }

// One cell of the x = -1 / x = 32 halo columns.  side 0: x = -1, side 1: x = 32.
GC_HD void p1_xhalo_cell(uint32_t cls, uint8_t* hx_b, int hr, int side) { hx_b[hr * 2 + side] = (uint8_t)cls; }

// Is halo cell (side, ty, tz) ever read?  Only by faces / corner-opacity tests of an emitting voxel at x = 0 (side 0) or
// x = 31 (side 1) in one of the 3x3 rows around it.  bvm[(ty + 1) * 2 + w]: bit (tz + 1) of the 64-bit mask of layer ty
// says "the boundary voxel of row (ty, tz) emits" (set for centre rows only).
constexpr int kBvWords = kRowsY * 2;   // per side
GC_HD uint64_t xhalo_needed_mask(const uint32_t* bvm, int ty)   // bit tz + 1 = halo cell (ty, tz) of this side may be read
{
    uint64_t m = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
    {
        const int r = ty + 1 + dy;
        if (r >= 0 && r < kRowsY) m |= (uint64_t)bvm[r * 2] | ((uint64_t)bvm[r * 2 + 1] << 32);
    }
    return m | (m << 1) | (m >> 1);
}
GC_HD uint32_t xhalo_needed(const uint32_t* bvm, int ty, int tz) { return (uint32_t)(xhalo_needed_mask(bvm, ty) >> (tz + 1)) & 1u; }

// One cell's light as the packed pair VertexLight sums, blockLight | sun' << 16, with GetSunlight's rules
// (Lighting.cpp:69-79): sun' = 15 at or above the column's surface, else the stored value with 0 read as MIN_LIGHT = 1
GC_HD uint32_t light_pair_of(uint32_t sun, uint32_t bl, uint32_t surf, int wy)
{
    uint32_t s = sun & 15u;
    if (s == 0) s = 1;
    if ((uint32_t)wy >= surf) s = 15;
    return (bl & 15u) | (s << 16);
}

// =====================================================================================================
// Phase 2: face masks of one centre row
// =====================================================================================================
struct RowFaces
{
    uint32_t f[6];        // up(+y), down(-y), front(+z), back(-z), right(+x), left(-x): BuildBlock's order, WorldRender.cpp:468-556
    uint32_t m0, m1, m2;  // mesh-type bit planes of the row
};

// drawn iff emits(v) && cull(v) < cull(n)   (CanDrawFace, WorldRender.cpp:387-391; "n != v" is implied by the
// strict inequality because cull is a function of the block id).  With the thermometer code g_k = cull >= k:
// cull(v) < cull(n)  <=>  OR_k (~g_k(v) & g_k(n)).
GC_HD uint32_t lt_mask(const uint32_t g[4], uint32_t n0, uint32_t n1, uint32_t n2, uint32_t n3)
{
    return (~g[0] & n0) | (~g[1] & n1) | (~g[2] & n2) | (~g[3] & n3);
}

GC_HD void p2_row_faces(const uint32_t* planes, const uint16_t* hx, int ty, int tz, RowFaces& r)
{
    const int hr = hrow(ty, tz);
    const Planes4* G = planes_g(planes);
    const Planes4 gv = G[hr], mv = planes_m(planes)[hr];
    const uint32_t g[4] = { gv.x, gv.y, gv.z, gv.w };
    r.m0 = mv.y; r.m1 = mv.z; r.m2 = mv.w;
    const uint32_t vis = ~(r.m0 & r.m1 & r.m2);
    const uint32_t h = hx[hr];

    const Planes4 up = G[hr + kRowsZ], dn = G[hr - kRowsZ], fr = G[hr + 1], bk = G[hr - 1];
    r.f[0] = vis & lt_mask(g, up.x, up.y, up.z, up.w);
    r.f[1] = vis & lt_mask(g, dn.x, dn.y, dn.z, dn.w);
    r.f[2] = vis & lt_mask(g, fr.x, fr.y, fr.z, fr.w);
    r.f[3] = vis & lt_mask(g, bk.x, bk.y, bk.z, bk.w);
    // +x neighbour of voxel x is voxel x+1 (bit x+1), voxel 31's is the x = 32 halo cell (hx high byte)
    r.f[4] = vis & lt_mask(g, (g[0] >> 1) | (((h >> 8) & 1u) << 31), (g[1] >> 1) | (((h >> 9) & 1u) << 31),
                           (g[2] >> 1) | (((h >> 10) & 1u) << 31), (g[3] >> 1) | (((h >> 11) & 1u) << 31));
    r.f[5] = vis & lt_mask(g, (g[0] << 1) | (h & 1u), (g[1] << 1) | ((h >> 1) & 1u),
                           (g[2] << 1) | ((h >> 2) & 1u), (g[3] << 1) | ((h >> 3) & 1u));
}

// total quads of the row and, packed 4 x 8 bit, the quads of mesh types 1..4 (type 0 = total - sum)
GC_HD void p2_row_counts(const RowFaces& r, int& total, uint32_t& typ4)
{
    total = popc(r.f[0]) + popc(r.f[1]) + popc(r.f[2]) + popc(r.f[3]) + popc(r.f[4]) + popc(r.f[5]);
    typ4 = 0;
    const uint32_t any = r.f[0] | r.f[1] | r.f[2] | r.f[3] | r.f[4] | r.f[5];
    if (any & (r.m0 | r.m1 | r.m2))
    {
        const uint32_t t1 = r.m0 & ~r.m1 & ~r.m2, t2 = ~r.m0 & r.m1 & ~r.m2, t3 = r.m0 & r.m1 & ~r.m2, t4 = ~r.m0 & ~r.m1 & r.m2;
        int c1 = 0, c2 = 0, c3 = 0, c4 = 0;
#pragma unroll
        for (int d = 0; d < 6; d++)
        {
            c1 += popc(r.f[d] & t1); c2 += popc(r.f[d] & t2); c3 += popc(r.f[d] & t3); c4 += popc(r.f[d] & t4);
        }
        typ4 = (uint32_t)c1 | ((uint32_t)c2 << 8) | ((uint32_t)c3 << 16) | ((uint32_t)c4 << 24);
    }
}

// =====================================================================================================
// Phase 3a: one row -> quad list entries in emission order (x ascending, then BuildBlock's face order)
// =====================================================================================================
// list entry: bits 0..2 face, 3..7 x, 8..12 tz, 13..18 ty, 19..21 mesh type; tlist = the quad's rank within its mesh type
GC_HD uint32_t quad_entry(int ty, int tz, int x, int d) { return (uint32_t)d | ((uint32_t)x << 3) | ((uint32_t)tz << 8) | ((uint32_t)ty << 13); }

// rank0: chunk-relative rank of the row's first quad; trank0[t]: rank of the row's first quad within mesh type t.
GC_HD void p3a_row(const RowFaces& r, int ty, int tz, uint32_t rank0, const uint32_t trank0[kMeshTypes], uint32_t* list, uint16_t* tlist)
{
    uint32_t any = r.f[0] | r.f[1] | r.f[2] | r.f[3] | r.f[4] | r.f[5];
    uint32_t rank = rank0;
    const uint32_t base = quad_entry(ty, tz, 0, 0);
    if (!(any & (r.m0 | r.m1 | r.m2)))
    {
        // every emitting voxel of the row has mesh type 0 (the common case)
        uint32_t tr = trank0[0];
        while (any)
        {
            const int x = lowbit(any);
            any &= any - 1;
            const uint32_t e = base | ((uint32_t)x << 3);
            // the voxel's drawn faces as a 6-bit set, then one step per face (most surface voxels draw one)
            uint32_t fb = ((r.f[0] >> x) & 1u) | (((r.f[1] >> x) & 1u) << 1) | (((r.f[2] >> x) & 1u) << 2)
                        | (((r.f[3] >> x) & 1u) << 3) | (((r.f[4] >> x) & 1u) << 4) | (((r.f[5] >> x) & 1u) << 5);
            do
            {
                const int d = lowbit(fb);
                fb &= fb - 1;
                list[rank] = e | (uint32_t)d; tlist[rank] = (uint16_t)tr; rank++; tr++;
            } while (fb);
        }
        return;
    }
    uint32_t tr[kMeshTypes];
#pragma unroll
    for (int t = 0; t < kMeshTypes; t++) tr[t] = trank0[t];
    while (any)
    {
        const int x = lowbit(any);
        any &= any - 1;
        const uint32_t t = ((r.m0 >> x) & 1u) | (((r.m1 >> x) & 1u) << 1) | (((r.m2 >> x) & 1u) << 2);
        // this voxel's running position in its stream, without dynamic register indexing
        uint32_t pos = tr[0];
        pos = t == 1 ? tr[1] : pos; pos = t == 2 ? tr[2] : pos; pos = t == 3 ? tr[3] : pos; pos = t == 4 ? tr[4] : pos;
        const uint32_t e = base | ((uint32_t)x << 3) | (t << 19);
        uint32_t n = 0;
#pragma unroll
        for (int d = 0; d < 6; d++)
            if ((r.f[d] >> x) & 1u) { list[rank] = e | (uint32_t)d; tlist[rank] = (uint16_t)(pos + n); rank++; n++; }
        tr[0] += t == 0 ? n : 0; tr[1] += t == 1 ? n : 0; tr[2] += t == 2 ? n : 0; tr[3] += t == 3 ? n : 0; tr[4] += t == 4 ? n : 0;
    }
}

// SetIndices (Mesh.cpp:32-50): o+2, o+1, o, o+3, o+2, o with o = (uint16)vertCount = 4 * rank, as three 32-bit words
GC_HD void quad_index_words(uint32_t rank, uint32_t w[3])
{
    const uint32_t o = (rank * 4u) & 0xFFFFu;
    w[0] = ((o + 2u) & 0xFFFFu) | (((o + 1u) & 0xFFFFu) << 16);
    w[1] = o | (((o + 3u) & 0xFFFFu) << 16);
    w[2] = ((o + 2u) & 0xFFFFu) | (o << 16);
}

// =====================================================================================================
// Phase 3b: one quad -> 4 vertices (48 bytes)
// =====================================================================================================
// Per face: the cell in front (a), the two in-plane axes (u, v) as (lt byte stride, row stride, x stride),
// and the four corners (BuildBlock's vertex order) as bits.  Packed so one table word serves a lane.
struct FaceDesc
{
    int8_t nx, ny, nz;       // outward normal
    int8_t ux, uy, uz;       // first in-plane axis
    int8_t vx, vy, vz;       // second in-plane axis
    uint8_t corners[4];      // per vertex: bit0 = +x corner, bit1 = +y, bit2 = +z   (WorldRender.cpp:475-553)
};

// up, down, front, back, right, left.  In-plane axes as VertexLight picks them (WorldRender.cpp:333-355):
// AXIS_Y -> (x, z); AXIS_Z -> (x, y); AXIS_X -> (y, z).
#define GC_FACE_TABLE { \
    { 0, 1, 0,  1, 0, 0,  0, 0, 1,  { 3, 7, 6, 2 } }, \
    { 0,-1, 0,  1, 0, 0,  0, 0, 1,  { 0, 4, 5, 1 } }, \
    { 0, 0, 1,  1, 0, 0,  0, 1, 0,  { 4, 6, 7, 5 } }, \
    { 0, 0,-1,  1, 0, 0,  0, 1, 0,  { 1, 3, 2, 0 } }, \
    { 1, 0, 0,  0, 1, 0,  0, 0, 1,  { 5, 7, 3, 1 } }, \
    {-1, 0, 0,  0, 1, 0,  0, 0, 1,  { 0, 2, 6, 4 } } }

// material word of a block type: bytes 0..5 texture layer per face, byte 6 alpha
struct MatEntry { uint32_t lo, hi; };

// opaque bit of cell (x, row hr), x in -1..32
// (plane_o points at the M records: the opaque word of row hr is plane_o[hr * 4])
GC_HD uint32_t opaque_at(const uint32_t* plane_o, const uint16_t* hx, int hr, int x)
{
    const uint32_t h = hx[hr];
    const uint64_t w = ((uint64_t)plane_o[hr * 4] << 1) | ((h >> P_OPAQUE) & 1u) | ((uint64_t)((h >> (8 + P_OPAQUE)) & 1u) << 33);
    return (uint32_t)(w >> (x + 1)) & 1u;
}

// VertexLight (WorldRender.cpp:329-383) on packed sums: fields are 17*nibble sums, so
//   4 samples: (17*S) >> 2          3 samples: (17*S)/3 = 5*S + floor(2*S/3), floor(n/3) = (n*171) >> 9 for n < 512
GC_HD uint32_t corner_color(uint32_t a, uint32_t b, uint32_t c, uint32_t d, bool open)
{
    if (open) return (((a + b + c + d) * 17u) >> 2) & 0x00FF00FFu;
    const uint32_t s = a + b + c;
    return s * 5u + (((s * 2u * 171u) >> 9) & 0x001F001Fu);
}

// ---- light of one cell, gathered from the world arrays ----
// What a quad needs to find the light of any cell of its chunk's halo tile (x in -1..32, ty in -1..64, tz in -1..32).
//   surf      [34][34] surface heights under the tile: row tz + 1 (columns cz-1 | cz | cz+1), byte x + 1 (cx-1 | cx | cx+1)
struct LightGather
{
    const uint8_t* sun;       // sunlight arena (slot-major bricks)
    const uint8_t* light;     // blockLight arena
    const uint8_t* surf;
    long long own;            // voxel offset of the chunk's own brick in the arenas
    int size_x, cx, cy, cz;   // window width (columns) and the chunk's position
};
constexpr int kSurfStride = kTX + 2;
constexpr int kSurfTileBytes = kRowsZ * kSurfStride;
constexpr int kColumnVoxels = 4 * kTX * kTY * kTZ;   // a column's four bricks are consecutive slots

GC_HD uint32_t load_u8(const uint8_t* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// voxel offset of tile coordinate x in -1..32 relative to the start of its row in the chunk's own brick
GC_HD int x_term(int x) { return (x < 0 ? -kColumnVoxels : (x > 31 ? kColumnVoxels : 0)) + (x & 31); }

// any cell of the halo tile as blockLight | sun' << 16 (slow path: picks the neighbour brick / column)
GC_HD uint32_t gather_cell(const LightGather& g, int x, int ty, int tz)
{
    const int wy = g.cy * kTY + ty;
    if (wy < 0) return 0u;                       // (0, 0, 0, 0) below the world
    if (wy >= 4 * kTY) return 0x000F000Fu;       // full light above it
    const int dcx = x < 0 ? -1 : (x > 31 ? 1 : 0), dcz = tz < 0 ? -1 : (tz > 31 ? 1 : 0);
    const long long slot = ((long long)(g.cz + dcz) * g.size_x + (g.cx + dcx)) * 4 + (wy >> 6);
    const long long vox = slot * (kTX * kTY * kTZ) + (x & 31) + 32 * ((wy & 63) + 64 * (tz & 31));
    return light_pair_of(load_u8(g.sun + vox), load_u8(g.light + vox), g.surf[(tz + 1) * kSurfStride + x + 1], wy);
}

// A face's geometry pre-digested for p3b_quad (built once per face from FaceDesc):
//   w0  row offset of the cell in front (ny*34 + nz) as int16 | nx as int8 << 16 | ny as int8 << 24
//   w1  row stride of in-plane axis u (uy*34 + uz) | ux << 16 | uy << 24           w2  the same for axis v
//   w3  PRMT selector: nibble k = which of the four footprint corners (bit 0: +u side, bit 1: +v side) vertex k sits on
//   w4  voxel-offset stride of u (ux + 32*uy + 2048*uz) as int16 | of v << 16
//   w5  surface-row stride of u (ux + 32*uz) as int16 | of v << 16
//   w6  footprint extent per axis: bit 0 x, bit 1 y, bit 2 z (1 = some in-plane axis runs along it)
//   off[k]  vertex k's first output word minus the voxel position: corner offset (x | y << 8 | z << 16) | u texcoord << 24
// (v never runs along x: the in-plane axes are (x,z), (x,y) or (y,z), see GC_FACE_TABLE.)
struct FacePlan { uint32_t w0, w1, w2, w3, w4, w5, w6, w7, off[4]; };
GC_HD FacePlan make_face_plan(const FaceDesc& f)
{
    FacePlan p;
    p.w0 = (uint32_t)(uint16_t)(int16_t)(f.ny * kRowsZ + f.nz) | ((uint32_t)(uint8_t)f.nx << 16) | ((uint32_t)(uint8_t)f.ny << 24);
    p.w1 = (uint32_t)(uint16_t)(int16_t)(f.uy * kRowsZ + f.uz) | ((uint32_t)(uint8_t)f.ux << 16) | ((uint32_t)(uint8_t)f.uy << 24);
    p.w2 = (uint32_t)(uint16_t)(int16_t)(f.vy * kRowsZ + f.vz) | ((uint32_t)(uint8_t)f.vx << 16) | ((uint32_t)(uint8_t)f.vy << 24);
    p.w3 = 0;
    for (int k = 0; k < 4; k++)
    {
        const uint32_t c = f.corners[k];
        const uint32_t cxb = c & 1u, cyb = (c >> 1) & 1u, czb = (c >> 2) & 1u;
        const uint32_t su = f.ux ? cxb : (f.uy ? cyb : czb);   // the corner bit of the axis u (v) points along
        const uint32_t sv = f.vx ? cxb : (f.vy ? cyb : czb);
        p.w3 |= (su | (sv << 1)) << (4 * k);
        const uint32_t u = (uint32_t)(k >> 1) & 1u;            // uv of the four vertices: (0,1) (0,0) (1,0) (1,1)  (WorldRender.cpp:475-553)
        p.off[k] = cxb | (cyb << 8) | (czb << 16) | (u << 24);
    }
    p.w4 = (uint32_t)(uint16_t)(int16_t)(f.ux + 32 * f.uy + 2048 * f.uz) | ((uint32_t)(uint16_t)(int16_t)(f.vx + 32 * f.vy + 2048 * f.vz) << 16);
    p.w5 = (uint32_t)(uint16_t)(int16_t)(f.ux + kSurfStride * f.uz) | ((uint32_t)(uint16_t)(int16_t)(f.vx + kSurfStride * f.vz) << 16);
    p.w6 = (uint32_t)((f.ux | f.vx) ? 1 : 0) | ((uint32_t)((f.uy | f.vy) ? 1 : 0) << 1) | ((uint32_t)((f.uz | f.vz) ? 1 : 0) << 2);
    p.w7 = 0;
    return p;
}

// out: 12 words = 4 vertices x 12 bytes {pos u8x3, uv u8x3, color u8x4 (L,L,L,S), alpha, 0}  (VertexInfo, Mesh.h:37-44)
GC_HD void p3b_quad(uint32_t entry, const FacePlan& f, const LightGather& g, const uint32_t* plane_o, const uint16_t* hx,
                    MatEntry mat, uint32_t out[12])
{
    const int d = entry & 7, x = (entry >> 3) & 31, tz = (entry >> 8) & 31, ty = (entry >> 13) & 63;
    const int ux = (int8_t)(f.w1 >> 16), vx = (int8_t)(f.w2 >> 16);
    const int uy = (int8_t)(f.w1 >> 24), vy = (int8_t)(f.w2 >> 24);
    const int hr_u = (int16_t)(f.w1 & 0xFFFFu), hr_v = (int16_t)(f.w2 & 0xFFFFu);   // row strides of the in-plane axes
    const int hr_n = (int16_t)(f.w0 & 0xFFFFu);
    const int ny = (int8_t)(f.w0 >> 24);
    const int ax = x + (int8_t)(f.w0 >> 16), ay = ty + ny, az = tz + (hr_n - ny * kRowsZ);
    const int hr_a = hrow(ty, tz) + hr_n;

    // 3x3 light neighbourhood in the layer in front of the face: cell (i, j) = a + i*u + j*v, i, j in -1..1,
    // and the opacity of the four edge cells (only b and c are ever tested, WorldRender.cpp:361-362)
    uint32_t l00, lm0, lp0, l0m, l0p, lmm, lmp, lpm, lpp, o_um, o_up, o_vm, o_vp;
    const int ey = (f.w6 >> 1) & 1u, ez = (f.w6 >> 2) & 1u;
    if ((unsigned)(ay - ey) <= (unsigned)(63 - 2 * ey) && (unsigned)(az - ez) <= (unsigned)(31 - 2 * ez))
    {
        // the footprint stays inside the chunk's own y / z range: constant strides along y and z; along x (axis u or
        // the normal, never v) each of the up to three positions picks its own column — the brick at the same height of
        // column cx-1 / cx+1 lies a fixed number of voxels away — so a row's quads at x = 0 and x = 31 take this path too
        const int x0 = ax - ux, x2 = ax + ux;
        const int dm = x_term(x0), d0 = x_term(ax), dp = x_term(x2);
        const int du = (int16_t)(f.w4 & 0xFFFFu) - ux, dv = (int16_t)(f.w4 >> 16);       // strides without the x part
        const int fu = (int16_t)(f.w5 & 0xFFFFu) - ux, fv = (int16_t)(f.w5 >> 16);
        const long long base = g.own + 32 * (ay + 64 * az);
        const uint8_t* ps = g.sun + base;
        const uint8_t* pl = g.light + base;
        const uint8_t* sf = g.surf + (az + 1) * kSurfStride + 1;
        const int om = dm - du, o0 = d0, op = dp + du;          // voxel offsets of the cells (-1, 0), (0, 0), (+1, 0)
        const int gm = x0 - fu, g0 = ax, gp = x2 + fu;          // their surface offsets
        const int wy = g.cy * kTY + ay;
        // issue the eighteen loads before any of them is consumed
        const uint32_t s00 = load_u8(ps + o0), sm0 = load_u8(ps + om), sp0 = load_u8(ps + op), s0m = load_u8(ps + o0 - dv), s0p = load_u8(ps + o0 + dv);
        const uint32_t smm = load_u8(ps + om - dv), smp = load_u8(ps + om + dv), spm = load_u8(ps + op - dv), spp = load_u8(ps + op + dv);
        const uint32_t b00 = load_u8(pl + o0), bm0 = load_u8(pl + om), bp0 = load_u8(pl + op), b0m = load_u8(pl + o0 - dv), b0p = load_u8(pl + o0 + dv);
        const uint32_t bmm = load_u8(pl + om - dv), bmp = load_u8(pl + om + dv), bpm = load_u8(pl + op - dv), bpp = load_u8(pl + op + dv);
        l00 = light_pair_of(s00, b00, sf[g0], wy);
        lm0 = light_pair_of(sm0, bm0, sf[gm], wy - uy); lp0 = light_pair_of(sp0, bp0, sf[gp], wy + uy);
        l0m = light_pair_of(s0m, b0m, sf[g0 - fv], wy - vy); l0p = light_pair_of(s0p, b0p, sf[g0 + fv], wy + vy);
        lmm = light_pair_of(smm, bmm, sf[gm - fv], wy - uy - vy); lmp = light_pair_of(smp, bmp, sf[gm + fv], wy - uy + vy);
        lpm = light_pair_of(spm, bpm, sf[gp - fv], wy + uy - vy); lpp = light_pair_of(spp, bpp, sf[gp + fv], wy + uy + vy);
    }
    else
    {
        const int uz = hr_u - uy * kRowsZ, vz = hr_v - vy * kRowsZ;
        l00 = gather_cell(g, ax, ay, az);
        lm0 = gather_cell(g, ax - ux, ay - uy, az - uz); lp0 = gather_cell(g, ax + ux, ay + uy, az + uz);
        l0m = gather_cell(g, ax - vx, ay - vy, az - vz); l0p = gather_cell(g, ax + vx, ay + vy, az + vz);
        lmm = gather_cell(g, ax - ux - vx, ay - uy - vy, az - uz - vz); lmp = gather_cell(g, ax - ux + vx, ay - uy + vy, az - uz + vz);
        lpm = gather_cell(g, ax + ux - vx, ay + uy - vy, az + uz - vz); lpp = gather_cell(g, ax + ux + vx, ay + uy + vy, az + uz + vz);
    }
    if ((unsigned)(ax - 1) < 30u)
    {
        // x in 1..30: no edge cell touches the x halo
        o_um = (plane_o[(hr_a - hr_u) * 4] >> (ax - ux)) & 1u; o_up = (plane_o[(hr_a + hr_u) * 4] >> (ax + ux)) & 1u;
        o_vm = (plane_o[(hr_a - hr_v) * 4] >> (ax - vx)) & 1u; o_vp = (plane_o[(hr_a + hr_v) * 4] >> (ax + vx)) & 1u;
    }
    else
    {
        o_um = opaque_at(plane_o, hx, hr_a - hr_u, ax - ux); o_up = opaque_at(plane_o, hx, hr_a + hr_u, ax + ux);
        o_vm = opaque_at(plane_o, hx, hr_a - hr_v, ax - vx); o_vp = opaque_at(plane_o, hx, hr_a + hr_v, ax + vx);
    }

    const uint32_t tex = (d < 4 ? (mat.lo >> (8 * d)) : (mat.hi >> (8 * (d - 4)))) & 255u;
    const uint32_t w1c = tex << 8, w2c = ((mat.hi >> 16) & 255u) << 16;   // constant parts of vertex words 1 and 2
    const uint32_t pos = (uint32_t)x | ((uint32_t)ty << 8) | ((uint32_t)tz << 16);

    // the four footprint corners in (side along u, side along v) order, then the face's vertex order by one PRMT each:
    // byte k of lv / sv = block light / sunlight colour of vertex k
    const uint32_t c0 = corner_color(l00, lm0, l0m, lmm, !(o_um & o_vm)), c1 = corner_color(l00, lp0, l0m, lpm, !(o_up & o_vm));
    const uint32_t c2 = corner_color(l00, lm0, l0p, lmp, !(o_um & o_vp)), c3 = corner_color(l00, lp0, l0p, lpp, !(o_up & o_vp));
    const uint32_t l4 = byte_perm(byte_perm(c0, c1, 0x0040), byte_perm(c2, c3, 0x0040), 0x5410);
    const uint32_t s4 = byte_perm(byte_perm(c0, c1, 0x0062), byte_perm(c2, c3, 0x0062), 0x5410);
    const uint32_t lv = byte_perm(l4, 0u, f.w3), sv = byte_perm(s4, 0u, f.w3);
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const uint32_t l = (lv >> (8 * k)) & 255u, sn = (sv >> (8 * k)) & 255u;
        const uint32_t v = (k == 0 || k == 3) ? 1u : 0u;
        out[3 * k + 0] = pos + f.off[k];             // corner offsets never carry (x <= 31, y <= 63, z <= 31)
        out[3 * k + 1] = v | w1c | (l << 16) | (l << 24);
        out[3 * k + 2] = l | (sn << 8) | w2c;
    }
}

}  // namespace gc

"""Shared helpers of the test-suite (test infrastructure)."""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

from gamecraft_b200.worldgen import HostWorld
from oracle import binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

B = dict(AIR=0, GRASS=1, DIRT=2, STONE=3, CRATE=4, SAND=5, STONE_BRICK=6, METAL_CRATE=7, WATER=8, WOOD=9, LEAVES=10, CLAY=11,
         SNOW=12, ICE=13, LANTERN=14, TRAMPOLINE=15, CACTUS=16, MAGMA=17, COOLED_MAGMA=18, OBSIDIAN=19, KILL_ZONE=20, LAVA=21,
         GLASS=22, SHOCKER=23)


def load_golden():
    with open(os.path.join(GOLDEN, "golden_meshes.json")) as f:
        return json.load(f)


def golden_world(entry, threads=1):
    kw = entry["kwargs"]
    size = entry["size"]
    origin = (-(size // 2), -(size // 2)) if kw else (0, 0)
    return HostWorld(size, size, origin=origin).build(entry["kind"], entry["seed"], threads=threads, **kw)


def input_digest(w):
    return [ob.fnv1a64(w.blocks), ob.fnv1a64(w.sun), ob.fnv1a64(w.light), ob.fnv1a64(w.surface)]


def mesh_signature(m):
    return [m.vert_count, [len(i) for i in m.indices], ob.fnv1a64(m.vertices) if m.vert_count else "",
            [ob.fnv1a64(np.ascontiguousarray(i)) if len(i) else "" for i in m.indices]]


def assert_same_mesh(got, want, where=""):
    assert got.vert_count == want.vert_count, "%s vertCount %d != %d" % (where, got.vert_count, want.vert_count)
    assert [len(i) for i in got.indices] == [len(i) for i in want.indices], "%s index counts" % where
    if got.vert_count:
        rows = np.nonzero((np.asarray(got.vertices) != np.asarray(want.vertices)).any(1))[0]
        assert len(rows) == 0, "%s %d vertices differ, first %d: got %s want %s" % (where, len(rows), rows[0], got.vertices[rows[0]], want.vertices[rows[0]])
    for t in range(5):
        assert np.array_equal(got.indices[t], want.indices[t]), "%s index stream %d differs" % (where, t)


def light_world(w):
    """ComputeSurface + light fill in the reference's column order (threads=1)."""
    w.compute_surface(threads=1)
    w.compute_light(threads=1)
    return w


def empty_world(size=3):
    return HostWorld(size, size)


def cap_edge_world(pairs: int, singles: int, block=3):
    """3x3 window whose centre chunk (1,1,1) holds `singles` isolated voxels (6 quads each) and `pairs` adjacent pairs
    (10 quads each), on a sparse lattice so nothing else touches."""
    w = HostWorld(3, 3)
    cells = [(x, y, z) for y in range(1, 63, 2) for z in range(1, 31, 2) for x in range(1, 29, 3)]
    assert pairs + singles <= len(cells)
    for i in range(pairs + singles):
        x, y, z = cells[i]
        w.set_block(32 + x, 64 + y, 32 + z, block)
        if i < pairs:
            w.set_block(32 + x + 1, 64 + y, 32 + z, block)
    return light_world(w)


# ---- host emulator of the kernel phases (tests/emu) ----
class _ChunkOut(ctypes.Structure):
    _fields_ = [("quad_base", ctypes.c_uint32), ("vert_count", ctypes.c_uint32), ("index_count", ctypes.c_uint32 * 5),
                ("index_offset", ctypes.c_uint32 * 5), ("quads", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 2)]


_emu = None


def emu_build_chunk(w, cx, cy, cz, table=None, kill_zone=20):
    global _emu
    if _emu is None:
        _emu = ctypes.CDLL(os.path.join(HERE, "emu", "libgcemu.so"))
    if table is None:
        arr, n, kill_zone = ob.default_block_table()
    else:
        arr = (ob.BlockInfo * 256)(*table)
        n = len(table)
    verts = np.zeros(10000 * 48, np.uint8)
    idx = np.zeros(10000 * 6, np.uint16)
    co = _ChunkOut()
    vp = ctypes.c_void_p
    _emu.gcemu_build_chunk(w.size_x, w.size_z, vp(w.blocks.ctypes.data), vp(w.sun.ctypes.data), vp(w.light.ctypes.data),
                           vp(w.surface.ctypes.data), arr, n, kill_zone, int(cx), int(cy), int(cz), vp(verts.ctypes.data), vp(idx.ctypes.data),
                           ctypes.byref(co))
    nv = co.vert_count
    m = ob.Mesh(verts[: nv * 12].reshape(nv, 12).copy(),
                [idx[co.index_offset[t]: co.index_offset[t] + co.index_count[t]].copy() for t in range(5)], co.flags & 1)
    m.quads_total = co.quads
    return m


def custom_table():
    """A block table unlike the reference's: more ids, cull / mesh type / opacity shuffled, invisible block with a low cull."""
    base, n, kz = ob.default_block_table()
    infos = [base[i] for i in range(n)]
    for i in range(n, 40):
        b = ob.BlockInfo()
        for f in range(6):
            b.textures[f] = (i * 7 + f * 3) % 61
        b.mesh_type = i % 5
        b.cull = (i * 3) % 5
        b.alpha = 255 - i
        b.invisible = 1 if i % 11 == 0 else 0
        b.opaque = 1 if i % 2 else 0
        infos.append(b)
    infos[3].cull = 1      # stone culls like glass
    infos[22].cull = 0     # glass culls like stone
    infos[8].mesh_type = 2
    infos[10].opaque = 0   # leaves let light through the corner test
    return infos, 25


def arbitrary_world(seed, fill=0.03, size=3):
    """Inputs no generator would produce: random block ids (all 24 kinds), random light nibbles and a random surface map that
    has nothing to do with the blocks.  The mesher is a function of its arrays, whatever they hold."""
    rng = np.random.default_rng(seed)
    w = HostWorld(size, size)
    w.blocks[:] = rng.integers(1, 24, w.blocks.shape, dtype=np.uint16) * (rng.random(w.blocks.shape) < fill)
    w.sun[:] = rng.integers(0, 16, w.sun.shape, dtype=np.uint8)
    w.light[:] = rng.integers(0, 16, w.light.shape, dtype=np.uint8)
    w.surface[:] = rng.integers(0, 256, w.surface.shape, dtype=np.uint8)
    return w

"""TEST INFRASTRUCTURE ONLY — ctypes bindings of the two CPU oracles (see ``oracle/__init__.py``)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "libgcoracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libgcref.so")
REFERENCE_ROOT = "/root/reference"

MAX_VERTICES = 40000
MAX_INDICES = 60000
MESH_TYPES = 5
MAX_BLOCKS = 256


def build(with_ref: bool | None = None) -> None:
    """Compile libgcoracle.so, and _ref/libgcref.so when the reference tree is present."""
    subprocess.check_call(["make", "-C", _HERE, "libgcoracle.so"], stdout=subprocess.DEVNULL)
    if with_ref is None:
        with_ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "Code"))
    if with_ref:
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


class BlockInfo(ctypes.Structure):
    _fields_ = [("textures", ctypes.c_uint8 * 6), ("mesh_type", ctypes.c_uint8), ("cull", ctypes.c_uint8),
                ("alpha", ctypes.c_uint8), ("invisible", ctypes.c_uint8), ("opaque", ctypes.c_uint8),
                ("reserved", ctypes.c_uint8)]


class _OrWorld(ctypes.Structure):
    _fields_ = [("size_x", ctypes.c_int), ("size_z", ctypes.c_int),
                ("blocks", ctypes.c_void_p), ("sun", ctypes.c_void_p), ("light", ctypes.c_void_p),
                ("surface", ctypes.c_void_p), ("table", BlockInfo * MAX_BLOCKS),
                ("table_count", ctypes.c_int), ("kill_zone_id", ctypes.c_int)]


class _OrMesh(ctypes.Structure):
    _fields_ = [("vertices", ctypes.c_uint8 * (MAX_VERTICES * 12)),
                ("indices", (ctypes.c_uint16 * MAX_INDICES) * MESH_TYPES),
                ("vert_count", ctypes.c_int32), ("index_count", ctypes.c_int32 * MESH_TYPES),
                ("index_allocated", ctypes.c_int32 * MESH_TYPES), ("overflow", ctypes.c_int32)]


class Mesh:
    """One chunk's mesh as plain numpy: ``vertices`` (n,12) u8, ``indices[t]`` u16, ``overflow``."""

    def __init__(self, vertices, indices, overflow=False, allocated=None):
        self.vertices = vertices
        self.indices = indices
        self.overflow = bool(overflow)
        self.allocated = allocated

    @property
    def vert_count(self):
        return len(self.vertices)

    @property
    def quads(self):
        return len(self.vertices) // 4

    def same_as(self, other) -> bool:
        return (self.vertices.shape == other.vertices.shape and np.array_equal(self.vertices, other.vertices)
                and all(np.array_equal(a, b) for a, b in zip(self.indices, other.indices)))


def fnv1a64(data: bytes | np.ndarray) -> str:
    L = oracle_lib()
    b = np.ascontiguousarray(data).view(np.uint8).reshape(-1) if isinstance(data, np.ndarray) else np.frombuffer(data, np.uint8)
    return "%016x" % L.gcor_fnv1a64(b.ctypes.data_as(ctypes.c_void_p), b.size, 0)


_olib = None
_rlib = None


def oracle_lib() -> ctypes.CDLL:
    global _olib
    if _olib is None:
        if not os.path.exists(ORACLE_LIB):
            build(with_ref=False)
        L = ctypes.CDLL(ORACLE_LIB)
        L.gcor_world_init.argtypes = [ctypes.POINTER(_OrWorld), ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.gcor_build_chunk.argtypes = [ctypes.POINTER(_OrWorld)] + [ctypes.c_int] * 3 + [ctypes.POINTER(_OrMesh)]
        L.gcor_chunk_overflowed.argtypes = [ctypes.POINTER(_OrWorld)] + [ctypes.c_int] * 7
        L.gcor_bench_build.argtypes = [ctypes.POINTER(_OrWorld), ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        L.gcor_bench_build.restype = ctypes.c_double
        L.gcor_check_batch.argtypes = [ctypes.POINTER(_OrWorld), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
        L.gcor_check_batch.restype = ctypes.c_int64
        L.gcor_build_chunk_uncapped.argtypes = [ctypes.POINTER(_OrWorld)] + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.gcor_build_chunk_uncapped.restype = ctypes.c_int64
        L.gcor_fnv1a64.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        L.gcor_fnv1a64.restype = ctypes.c_uint64
        L.gcor_default_block_table.argtypes = [ctypes.POINTER(BlockInfo), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.gcor_compute_surface.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.gcor_light_fill.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6
        L.gcor_default_light_rules.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.gcor_rle_encode_chunk.argtypes = [ctypes.c_void_p, ctypes.c_uint16, ctypes.c_void_p]
        L.gcor_rle_decode_chunk.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.gcor_set_block.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
        _olib = L
    return _olib


def ref_lib() -> ctypes.CDLL:
    global _rlib
    if _rlib is None:
        L = ctypes.CDLL(REF_LIB)
        L.gcref_world_create.restype = ctypes.c_void_p
        L.gcref_world_create.argtypes = [ctypes.c_int] * 4
        L.gcref_world_destroy.argtypes = [ctypes.c_void_p]
        L.gcref_block_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.gcref_block_light_steps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.gcref_block_light_emitted.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.gcref_rle_save_group.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.gcref_rle_load_group.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.gcref_world_set_origin.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.gcref_set_noise.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.gcref_set_seed.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.gcref_generate_noise_terrain.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.gcref_rle_store_record.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.gcref_rle_clear_region.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.gcref_set_chunk.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
        L.gcref_get_chunk.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
        L.gcref_set_surface.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.gcref_get_surface.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.gcref_generate_flat.argtypes = [ctypes.c_void_p]
        L.gcref_generate.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.gcref_compute_surface.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.gcref_compute_surface_all.argtypes = [ctypes.c_void_p]
        L.gcref_light_all.argtypes = [ctypes.c_void_p]
        L.gcref_build_chunk.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
        L.gcref_chunk_overflowed.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 7
        L.gcref_set_block.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4
        L.gcref_set_chunk_built.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5
        L.gcref_pending_updates.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.gcref_bench_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        L.gcref_bench_build.restype = ctypes.c_double
        L.gcref_check_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
        L.gcref_check_batch.restype = ctypes.c_int64
        _rlib = L
    return _rlib


def default_light_rules():
    """(step u8[256] with 255 = opaque, emitted u8[256]) as restated in gc_light_cpu.c from Code/Block.cpp:139-276."""
    step, emitted = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
    oracle_lib().gcor_default_light_rules(step.ctypes.data, emitted.ctypes.data)
    return step, emitted


def compute_surface(host) -> np.ndarray:
    """gcor_compute_surface over a HostWorld's blocks -> a new surface array shaped like host.surface."""
    out = np.zeros_like(host.surface)
    oracle_lib().gcor_compute_surface(host.size_x, host.size_z, host.blocks.ctypes.data, out.ctypes.data)
    return out


def light_fill(host, surface=None, step=None, emitted=None):
    """gcor_light_fill over a HostWorld's blocks (+ surface) -> (sun, light) arrays shaped like host.sun / host.light."""
    if step is None:
        step, emitted = default_light_rules()
    surface = host.surface if surface is None else surface
    sun, light = np.zeros_like(host.sun), np.zeros_like(host.light)
    oracle_lib().gcor_light_fill(host.size_x, host.size_z, host.blocks.ctypes.data, np.ascontiguousarray(surface).ctypes.data,
                                 np.ascontiguousarray(step, np.uint8).ctypes.data, np.ascontiguousarray(emitted, np.uint8).ctypes.data,
                                 sun.ctypes.data, light.ctypes.data)
    return sun, light


def rle_encode_chunk(blocks: np.ndarray, offset: int = 0) -> np.ndarray:
    """gcor_rle_encode_chunk: one chunk's block ids (65536 uint16) -> its region record."""
    out = np.zeros(2 + 2 * 65536, np.uint16)
    n = oracle_lib().gcor_rle_encode_chunk(np.ascontiguousarray(blocks, np.uint16).ctypes.data, offset, out.ctypes.data)
    return out[:n].copy()


def rle_decode_chunk(record: np.ndarray):
    """gcor_rle_decode_chunk: a region record -> (status, 65536 uint16)."""
    rec = np.ascontiguousarray(record, np.uint16)
    out = np.zeros(65536, np.uint16)
    st = oracle_lib().gcor_rle_decode_chunk(rec.ctypes.data, len(rec), out.ctypes.data)
    return st, out


def set_block(host, vert_table, wx, wy, wz, block, flagged, cull=None, kill_zone=20) -> int:
    """SetBlock's decisions on a HostWorld's blocks / surface (edited in place) -> GC_EDIT_* status; marks flagged[slot]."""
    if cull is None:
        cull = np.zeros(256, np.uint8)
        table, n, kill_zone = default_block_table()
        for i in range(n):
            cull[i] = table[i].cull
    vert_table = np.ascontiguousarray(vert_table, np.int32)
    return oracle_lib().gcor_set_block(host.size_x, host.size_z, host.blocks.ctypes.data, host.surface.ctypes.data, vert_table.ctypes.data,
                                       cull.ctypes.data, kill_zone, int(wx), int(wy), int(wz), int(block), flagged.ctypes.data)


def default_block_table():
    """(table: list[BlockInfo] of 24, kill_zone_id) as restated in gc_mesh_cpu.c from Code/Block.cpp:139-276."""
    arr = (BlockInfo * MAX_BLOCKS)()
    n = ctypes.c_int(0)
    kz = ctypes.c_int(0)
    oracle_lib().gcor_default_block_table(arr, ctypes.byref(n), ctypes.byref(kz))
    return arr, n.value, kz.value


def _check_batch(call, coords, desc, vertices, indices, threads):
    coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
    desc = np.ascontiguousarray(desc)
    assert desc.dtype.itemsize == 64 and len(desc) == len(coords)
    vertices, indices = np.ascontiguousarray(vertices), np.ascontiguousarray(indices)
    arena_quads = min(vertices.nbytes // 48, indices.nbytes // 12)
    result = np.zeros(len(coords), np.uint8)
    q = ctypes.c_int64(0)
    bad = call(coords.ctypes.data, len(coords), desc.ctypes.data, vertices.ctypes.data, indices.ctypes.data, arena_quads,
               int(threads) if threads and threads > 0 else (os.cpu_count() or 1), result.ctypes.data, ctypes.byref(q))
    return int(bad), result, int(q.value)


class OracleWorld:
    """The restated oracle over a HostWorld's arrays (borrowed, not copied)."""

    def __init__(self, host):
        self.host = host
        self._w = _OrWorld()
        oracle_lib().gcor_world_init(ctypes.byref(self._w), host.size_x, host.size_z, host.blocks.ctypes.data,
                                     host.sun.ctypes.data, host.light.ctypes.data, host.surface.ctypes.data)
        self._m = _OrMesh()

    def set_block_table(self, infos, kill_zone_id=20):
        for i, b in enumerate(infos):
            self._w.table[i] = b
        self._w.table_count = len(infos)
        self._w.kill_zone_id = kill_zone_id

    def build_chunk(self, cx, cy, cz) -> Mesh:
        m = self._m
        oracle_lib().gcor_build_chunk(ctypes.byref(self._w), int(cx), int(cy), int(cz), ctypes.byref(m))
        n = m.vert_count
        verts = np.frombuffer(m.vertices, np.uint8, n * 12).reshape(n, 12).copy()
        idx = [np.frombuffer(m.indices[t], np.uint16, m.index_count[t]).copy() for t in range(MESH_TYPES)]
        return Mesh(verts, idx, m.overflow, [int(a) for a in m.index_allocated])

    def build_chunk_uncapped(self, cx, cy, cz, max_quads: int = 196608) -> Mesh:
        """The traversal with the reference's cap lifted and 32-bit indices (the "stress" variant; not a reference behaviour)."""
        verts = np.zeros(max_quads * 48, np.uint8)
        idx = [np.zeros(max_quads * 6, np.uint32) for _ in range(MESH_TYPES)]
        ptrs = (ctypes.c_void_p * MESH_TYPES)(*[a.ctypes.data for a in idx])
        counts = np.zeros(MESH_TYPES, np.int64)
        q = oracle_lib().gcor_build_chunk_uncapped(ctypes.byref(self._w), int(cx), int(cy), int(cz), verts.ctypes.data, ptrs, max_quads, counts.ctypes.data)
        assert q >= 0
        return Mesh(verts[: q * 48].reshape(q * 4, 12).copy(), [idx[t][: counts[t]].copy() for t in range(MESH_TYPES)])

    def chunk_overflowed(self, cx, cy, cz, x, y, z, total_vertices) -> bool:
        return bool(oracle_lib().gcor_chunk_overflowed(ctypes.byref(self._w), cx, cy, cz, x, y, z, total_vertices))

    def check_batch(self, coords, desc, vertices, indices, threads: int = 0):
        """Exhaustive byte comparison of a mesher batch (64-byte GcChunkOut descriptors + the downloaded vertex / index
        arenas) with this oracle's build of every chunk -> (mismatching chunks, per-chunk result bytes, quads compared)."""
        return _check_batch(lambda *a: oracle_lib().gcor_check_batch(ctypes.byref(self._w), *a), coords, desc, vertices, indices, threads)

    def bench(self, coords: np.ndarray, threads: int):
        coords = np.ascontiguousarray(coords, np.int32)
        q = ctypes.c_int64(0)
        sec = oracle_lib().gcor_bench_build(ctypes.byref(self._w), coords.ctypes.data, len(coords), threads, ctypes.byref(q))
        return sec, q.value


class RefWorld:
    """The reference's own mesher (compiled verbatim) over a copy of a HostWorld (must be square)."""

    def __init__(self, host=None, size=None, seed=1337, radius=2**31 - 1, falloff=None):
        self.L = ref_lib()
        if host is not None:
            assert host.size_x == host.size_z, "the reference window is square (World::size)"
            size = host.size_x
        self.size = size
        self.h = ctypes.c_void_p(self.L.gcref_world_create(size, seed, radius, radius - 64 if falloff is None else falloff))
        if host is not None:
            self.upload(host)

    def close(self):
        if self.h:
            self.L.gcref_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, host, light=True):
        for cz in range(self.size):
            for cx in range(self.size):
                for cy in range(4):
                    s = host.slot(cx, cy, cz)
                    self.L.gcref_set_chunk(self.h, cx, cy, cz, host.blocks[s].ctypes.data,
                                           host.sun[s].ctypes.data if light else None,
                                           host.light[s].ctypes.data if light else None)
                self.L.gcref_set_surface(self.h, cx, cz, host.surface[host.column(cx, cz)].ctypes.data)

    def download(self, host):
        for cz in range(self.size):
            for cx in range(self.size):
                for cy in range(4):
                    s = host.slot(cx, cy, cz)
                    self.L.gcref_get_chunk(self.h, cx, cy, cz, host.blocks[s].ctypes.data, host.sun[s].ctypes.data, host.light[s].ctypes.data)
                self.L.gcref_get_surface(self.h, cx, cz, host.surface[host.column(cx, cz)].ctypes.data)

    def generate_flat(self):
        self.L.gcref_generate_flat(self.h)

    def generate(self, kind):
        """The reference's own noise-free generators over every column: "flat", "grid", "void" or "dungeon"."""
        assert self.L.gcref_generate(self.h, {"flat": 0, "grid": 3, "void": 4, "dungeon": 8}[kind]) == 0

    def generate_noise_terrain(self, kind, seed, own_random=True):
        """The reference's own GenerateForest / Islands / Snow / Desert / VolcanicTerrain over this project's noise fields
        (gamecraft_b200/worldgen's gcw_noise plugged into the FastNoiseSIMD stand-in).  Trees and cacti come from rand(): the
        producer's per-column stream with ``own_random``, else the C library's."""
        from gamecraft_b200 import worldgen

        wl = worldgen.lib()
        fn = ctypes.cast(wl.gcw_noise, ctypes.c_void_p)
        crng, nxt = (ctypes.cast(wl.gcw_column_rng, ctypes.c_void_p), ctypes.cast(wl.gcw_rng_next, ctypes.c_void_p)) if own_random else (None, None)
        self.L.gcref_set_noise(fn, crng, nxt)
        try:
            self.L.gcref_set_seed(self.h, int(seed))
            rc = self.L.gcref_generate_noise_terrain(self.h, {"forest": 1, "islands": 2, "snow": 5, "desert": 6, "volcanic": 7}[kind])
            assert rc == 0, rc
        finally:
            self.L.gcref_set_noise(None, None, None)

    def compute_surface(self):
        self.L.gcref_compute_surface_all(self.h)

    def compute_light(self):
        self.L.gcref_light_all(self.h)

    def block_table(self):
        out = np.zeros(24 * 12, np.int32)
        self.L.gcref_block_table(self.h, out.ctypes.data)
        return out.reshape(24, 12)

    def light_steps(self):
        out = np.zeros(24, np.int32)
        self.L.gcref_block_light_steps(self.h, out.ctypes.data)
        return out

    def rle_save_group(self, cx, cz):
        """SaveGroup on one column -> list of four uint16 records."""
        out = np.zeros(4 * (2 + 2 * 65536), np.uint16)
        sizes = np.zeros(4, np.int32)
        total = self.L.gcref_rle_save_group(self.h, cx, cz, out.ctypes.data, len(out), sizes.ctypes.data)
        assert total >= 0
        recs, p = [], 0
        for n in sizes:
            recs.append(out[p:p + n].copy()); p += int(n)
        return recs

    def set_origin(self, ox, oz):
        """group->pos = window position + (ox, oz) for every column (only the region functions read it)."""
        self.L.gcref_world_set_origin(self.h, ox, oz)

    def rle_store_record(self, cx, cz, record) -> int:
        """LoadRegionFile's filing of one record read from a region file: word 0 picks the region's list."""
        rec = np.ascontiguousarray(record, np.uint16)
        return int(self.L.gcref_rle_store_record(self.h, cx, cz, rec.ctypes.data, len(rec)))

    def rle_clear_region(self, cx, cz):
        self.L.gcref_rle_clear_region(self.h, cx, cz)

    def rle_load_group(self, cx, cz) -> bool:
        """LoadGroupFromDisk on one column (decodes what rle_save_group stored into the column's chunks)."""
        return bool(self.L.gcref_rle_load_group(self.h, cx, cz))

    def light_emitted(self):
        out = np.zeros(24, np.int32)
        self.L.gcref_block_light_emitted(self.h, out.ctypes.data)
        return out

    def build_chunk(self, cx, cy, cz) -> Mesh:
        verts = np.zeros(MAX_VERTICES * 12, np.uint8)
        idx = [np.zeros(MAX_INDICES, np.uint16) for _ in range(MESH_TYPES)]
        idxp = (ctypes.c_void_p * MESH_TYPES)(*[a.ctypes.data for a in idx])
        counts = np.zeros(11, np.int32)
        self.L.gcref_build_chunk(self.h, int(cx), int(cy), int(cz), verts.ctypes.data, idxp, counts.ctypes.data)
        n = int(counts[0])
        return Mesh(verts[: n * 12].reshape(n, 12).copy(), [idx[t][: counts[1 + t]].copy() for t in range(MESH_TYPES)],
                    False, [int(a) for a in counts[6:11]])

    def chunk_overflowed(self, cx, cy, cz, x, y, z, total_vertices) -> bool:
        return bool(self.L.gcref_chunk_overflowed(self.h, cx, cy, cz, x, y, z, total_vertices))

    def set_block(self, wx, wy, wz, block) -> bool:
        """The reference's own SetBlock (overflow pre-check, RecomputeLight, FlagChunkForUpdate); True if the stored block changed."""
        return bool(self.L.gcref_set_block(self.h, int(wx), int(wy), int(wz), int(block)))

    def set_chunk_built(self, cx, cy, cz, built=True, total_vertices=0):
        self.L.gcref_set_chunk_built(self.h, int(cx), int(cy), int(cz), int(bool(built)), int(total_vertices))

    def pending_updates(self, clear=True) -> np.ndarray:
        """chunk->pendingUpdate of every chunk, shaped (size_z, size_x, 4)."""
        out = np.zeros(self.size * self.size * 4, np.uint8)
        self.L.gcref_pending_updates(self.h, out.ctypes.data, int(clear))
        return out.reshape(self.size, self.size, 4)

    def check_batch(self, coords, desc, vertices, indices, threads: int = 0):
        """As OracleWorld.check_batch, against the reference's own BuildChunkAsync (chunks with flags != 0 count as mismatches)."""
        return _check_batch(lambda *a: self.L.gcref_check_batch(self.h, *a), coords, desc, vertices, indices, threads)

    def bench(self, coords: np.ndarray, threads: int):
        coords = np.ascontiguousarray(coords, np.int32)
        q = ctypes.c_int64(0)
        sec = self.L.gcref_bench_build(self.h, coords.ctypes.data, len(coords), threads, ctypes.byref(q))
        return sec, q.value

"""Deterministic synthetic-world producer (ctypes binding of ``libgcworld.so``, see ``gcworld.c``).

Input producer only: it restates the reference's terrain / surface / light-fill *rules*
(Code/Generation.cpp:83-836, Code/Dungeon.cpp:17-31; Code/Lighting.cpp:5-26,109-281) with its own seeded noise, because
the reference's FastNoiseSIMD + ``rand()`` worlds are not reproducible.  The mesher never links it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgcworld.so")

CHUNK_VOXELS = 65536
CHUNKS_PER_COLUMN = 4

FLAT, FOREST, ISLANDS, CHECKER, NOISE, STRESS, MIXED, EMPTY, SNOW, DESERT, VOLCANIC, DUNGEON = range(12)
KINDS = {"flat": FLAT, "forest": FOREST, "islands": ISLANDS, "checker": CHECKER, "noise": NOISE,
         "stress": STRESS, "mixed": MIXED, "empty": EMPTY, "snow": SNOW, "desert": DESERT, "volcanic": VOLCANIC,
         "dungeon": DUNGEON}
INT_MAX = 2**31 - 1


class _GcwWorld(ctypes.Structure):
    _fields_ = [("size_x", ctypes.c_int32), ("size_z", ctypes.c_int32),
                ("origin_cx", ctypes.c_int32), ("origin_cz", ctypes.c_int32),
                ("blocks", ctypes.c_void_p), ("sun", ctypes.c_void_p),
                ("light", ctypes.c_void_p), ("surface", ctypes.c_void_p)]


def build(force: bool = False) -> str:
    """Compile libgcworld.so in-tree (gcc, seconds)."""
    src = os.path.join(_HERE, "gcworld.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgcworld.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.gcw_generate.argtypes = [ctypes.POINTER(_GcwWorld), ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.gcw_compute_surface.argtypes = [ctypes.POINTER(_GcwWorld), ctypes.c_int]
        L.gcw_light.argtypes = [ctypes.POINTER(_GcwWorld), ctypes.c_void_p, ctypes.c_int]
        L.gcw_build.argtypes = [ctypes.POINTER(_GcwWorld), ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.gcw_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


class HostWorld:
    """A window of ``size_x x size_z`` columns in host memory, brick-major like the reference's chunks.

    ``slot(cx,cy,cz) = (cz*size_x + cx)*4 + cy``; ``blocks[slot, x + 32*(y + 64*z)]`` (Code/World.cpp:278-281);
    ``surface[cz*size_x + cx, z*32 + x]`` (Code/Lighting.cpp:7).  Arrays may be supplied (e.g. views of pinned
    torch tensors) or are allocated as numpy arrays.
    """

    def __init__(self, size_x: int, size_z: int, origin=(0, 0), arrays=None):
        self.size_x, self.size_z = int(size_x), int(size_z)
        self.origin = (int(origin[0]), int(origin[1]))
        ncol = self.size_x * self.size_z
        nslot = ncol * CHUNKS_PER_COLUMN
        if arrays is None:
            self.blocks = np.zeros((nslot, CHUNK_VOXELS), np.uint16)
            self.sun = np.zeros((nslot, CHUNK_VOXELS), np.uint8)
            self.light = np.zeros((nslot, CHUNK_VOXELS), np.uint8)
            self.surface = np.zeros((ncol, 1024), np.uint8)
        else:
            self.blocks, self.sun, self.light, self.surface = arrays
            assert self.blocks.shape == (nslot, CHUNK_VOXELS) and self.blocks.dtype == np.uint16
            assert self.sun.shape == (nslot, CHUNK_VOXELS) and self.sun.dtype == np.uint8
            assert self.light.shape == (nslot, CHUNK_VOXELS) and self.light.dtype == np.uint8
            assert self.surface.shape == (ncol, 1024) and self.surface.dtype == np.uint8
        for a in (self.blocks, self.sun, self.light, self.surface):
            assert a.flags["C_CONTIGUOUS"]

    # -- addressing -------------------------------------------------------------------------------
    def slot(self, cx: int, cy: int, cz: int) -> int:
        return (cz * self.size_x + cx) * CHUNKS_PER_COLUMN + cy

    def column(self, cx: int, cz: int) -> int:
        return cz * self.size_x + cx

    def interior_chunks(self, margin: int = 1) -> np.ndarray:
        """(n,3) int32 chunk coords (cx,cy,cz) of every chunk at least ``margin`` columns from the window
        edge, in column-major (z, x) then y order — the reference never meshes edge columns (WorldRender.cpp:235)."""
        out = [(cx, cy, cz)
               for cz in range(margin, self.size_z - margin)
               for cx in range(margin, self.size_x - margin)
               for cy in range(CHUNKS_PER_COLUMN)]
        return np.asarray(out, np.int32).reshape(-1, 3)

    def set_block(self, wx: int, wy: int, wz: int, block: int) -> None:
        cx, x = divmod(wx, 32)
        cz, z = divmod(wz, 32)
        cy, y = divmod(wy, 64)
        self.blocks[self.slot(cx, cy, cz), x + 32 * (y + 64 * z)] = block

    def get_block(self, wx: int, wy: int, wz: int) -> int:
        cx, x = divmod(wx, 32)
        cz, z = divmod(wz, 32)
        cy, y = divmod(wy, 64)
        return int(self.blocks[self.slot(cx, cy, cz), x + 32 * (y + 64 * z)])

    # -- production -------------------------------------------------------------------------------
    def _struct(self) -> _GcwWorld:
        return _GcwWorld(self.size_x, self.size_z, self.origin[0], self.origin[1],
                         self.blocks.ctypes.data, self.sun.ctypes.data, self.light.ctypes.data, self.surface.ctypes.data)

    def generate(self, kind, seed: int = 1337, radius: int = INT_MAX, falloff: int | None = None, threads: int | None = None):
        kind = KINDS[kind] if isinstance(kind, str) else int(kind)
        if falloff is None:
            falloff = radius - 64 if radius != INT_MAX else INT_MAX - 64  # World.cpp:911
        threads = threads or lib().gcw_max_threads()
        s = self._struct()
        lib().gcw_generate(ctypes.byref(s), kind, seed & 0xFFFFFFFF, radius, falloff, threads)
        return self

    def compute_surface(self, threads: int | None = None):
        s = self._struct()
        lib().gcw_compute_surface(ctypes.byref(s), threads or lib().gcw_max_threads())
        return self

    def compute_light(self, threads: int | None = None, clear: bool = True):
        if clear:
            self.sun[:] = 0
            self.light[:] = 0
        s = self._struct()
        lib().gcw_light(ctypes.byref(s), None, threads or lib().gcw_max_threads())
        return self

    def build(self, kind, seed: int = 1337, radius: int = INT_MAX, falloff: int | None = None, threads: int | None = None):
        """generate -> ComputeSurface -> light flood fill."""
        self.generate(kind, seed, radius, falloff, threads)
        self.compute_surface(threads)
        self.compute_light(threads, clear=False)
        return self


def max_threads() -> int:
    return lib().gcw_max_threads()

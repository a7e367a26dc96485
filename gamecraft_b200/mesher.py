"""Host-side mirror of the reference's mesh-build call sites over ``libgcmesh.so`` (ctypes; no torch types).

Reference call sites mirrored (Code/WorldRender.cpp):
  * ``BuildChunk(state, world, chunk)``  (:69-78)  -> :func:`build_chunk`
  * ``RebuildChunks(state, world)``      (:80-91)  -> :func:`rebuild_chunks`  (one batched GPU submission)
  * ``chunk->meshData`` (Code/Mesh.h:46-52)         -> :class:`ChunkMesh`
The CUDA library is mandatory: importing works anywhere, but creating a :class:`Context` without a usable
B200 raises :class:`GcMeshError` — there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("GCMESH_LIB") or os.path.join(CSRC, "libgcmesh.so")   # GCMESH_LIB: development A/B builds

CHUNK_VOXELS = 65536
MESH_TYPES = 5
MAX_QUADS = 10000
MESH_OPAQUE, MESH_CUTOUT, MESH_TRANSPARENT, MESH_FLUID, MESH_MAGMA = range(5)
CHUNK_OVERFLOW = 1
CHUNK_NO_SPACE = 2
CHUNK_HALO_MISSING = 4
WORLD_SURFACE_BOUNDS_BLOCKS = 1      # GC_WORLD_SURFACE_BOUNDS_BLOCKS
WORLD_COPY_RUN = 2                   # GC_WORLD_COPY_RUN
WORLD_SURFACE_BOUNDS_MESH = 3        # GC_WORLD_SURFACE_BOUNDS_MESH
WORLD_EDIT_REPLAY = 4                # GC_WORLD_EDIT_REPLAY
BATCH_UNCAPPED = 1                   # GC_BATCH_UNCAPPED

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
SOURCES = ["gcmesh_kernel.cu", "gclight_kernel.cu", "gcrle_kernel.cu", "gcedit_kernel.cu", "gchalo_kernel.cu", "gcgen_kernel.cu", "gcfill_kernel.cu", "gcmesh_api.cu"]


GEN_KINDS = {"flat": 0, "forest": 1, "islands": 2, "grid": 3, "void": 4, "snow": 5, "desert": 6, "volcanic": 7, "dungeon": 8}   # GC_GEN_*


class GcMeshError(RuntimeError):
    """A C-ABI call failed; ``status`` is its negative GcStatus."""

    def __init__(self, message: str, status: int = 0):
        super().__init__(message)
        self.status = status


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libgcmesh.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import glob

    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(_HERE, "..", "include", "gcmesh.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", LIB_PATH, "-lcudart"]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


class BlockInfo(ctypes.Structure):
    """GcBlockInfo: the fields of the reference's BlockData the mesher reads (Code/Block.h:85-100)."""
    _fields_ = [("textures", ctypes.c_uint8 * 6), ("mesh_type", ctypes.c_uint8), ("cull", ctypes.c_uint8),
                ("alpha", ctypes.c_uint8), ("invisible", ctypes.c_uint8), ("opaque", ctypes.c_uint8),
                ("reserved", ctypes.c_uint8)]


class ChunkOut(ctypes.Structure):
    """GcChunkOut."""
    _fields_ = [("quad_base", ctypes.c_uint32), ("vert_count", ctypes.c_uint32),
                ("index_count", ctypes.c_uint32 * MESH_TYPES), ("index_offset", ctypes.c_uint32 * MESH_TYPES),
                ("quads", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 2)]


CHUNK_OUT_DTYPE = np.dtype([("quad_base", "<u4"), ("vert_count", "<u4"), ("index_count", "<u4", (5,)),
                            ("index_offset", "<u4", (5,)), ("quads", "<u4"), ("flags", "<u4"), ("reserved", "<u4", (2,))])
assert CHUNK_OUT_DTYPE.itemsize == ctypes.sizeof(ChunkOut) == 64
HALO_HANDLE_BYTES = 320                                                                                                # GC_HALO_HANDLE_BYTES
DEVICE_MESH_DTYPE = np.dtype([("chunk", "<i4"), ("reserved", "<i4"), ("vertices", "<u8"), ("indices", "<u8", (5,))])           # GcDeviceMesh
BLOCK_EDIT_DTYPE = np.dtype([("wx", "<i4"), ("wy", "<i4"), ("wz", "<i4"), ("block", "<u2"), ("reserved", "<u2")])   # GcBlockEdit
EDIT_APPLIED, EDIT_UNCHANGED, EDIT_IGNORED, EDIT_NOT_BUILT, EDIT_OVERFLOW = range(5)                                   # GC_EDIT_*
RLE_CHUNK_DTYPE = np.dtype([("cx", "<i4"), ("cy", "<i4"), ("cz", "<i4"), ("first_word", "<u4"), ("words", "<u4")])   # GcRleChunk

EXPORTS = [
    "gcmesh_version", "gcmesh_last_error", "gcmesh_create", "gcmesh_destroy", "gcmesh_set_stream",
    "gcmesh_set_block_table", "gcmesh_default_block_table", "gcmesh_synchronize",
    "gcmesh_world_create", "gcmesh_world_destroy", "gcmesh_world_upload_chunk", "gcmesh_world_upload_surface",
    "gcmesh_world_upload_all", "gcmesh_world_upload_rows", "gcmesh_world_upload_sparse", "gcmesh_world_set_option", "gcmesh_world_set_origin", "gcmesh_world_validate_blocks", "gcmesh_query", "gcmesh_world_upload_launches", "gcmesh_world_upload_bytes", "gcmesh_world_device_ptrs", "gcmesh_world_halo_bytes",
    "gcmesh_world_pack_halo", "gcmesh_world_unpack_halo",
    "gcmesh_world_halo_export", "gcmesh_world_halo_connect_ipc", "gcmesh_world_halo_connect_local", "gcmesh_world_halo_push", "gcmesh_world_halo_status",
    "gcmesh_default_light_rules", "gcmesh_set_light_rules", "gcmesh_world_compute_surface", "gcmesh_world_light",
    "gcmesh_world_light_stats", "gcmesh_world_download",
    "gcmesh_world_upload_rle", "gcmesh_world_rle_refused", "gcmesh_world_encode_rle",
    "gcmesh_world_set_blocks", "gcmesh_world_vertex_counts", "gcmesh_world_generate",
    "gcmesh_batch_create", "gcmesh_batch_create_ex", "gcmesh_batch_destroy", "gcmesh_submit", "gcmesh_submit_exchange", "gcmesh_batch_set_timing", "gcmesh_mesh_host", "gcmesh_wait", "gcmesh_batch_results",
    "gcmesh_batch_download", "gcmesh_batch_fill_meshdata", "gcmesh_batch_fill_device", "gcmesh_batch_device_ptrs", "gcmesh_batch_elapsed_ms",
    "gcmesh_batch_launches", "gcmesh_kernel_info", "gcmesh_batch_profile",
]

_lib = None


def lib() -> ctypes.CDLL:
    """Load libgcmesh.so, rebuilding it first when a source or header is newer (where nvcc exists: the GPU box runs the
    library that travelled with the snapshot).  Fails loudly when it cannot be loaded."""
    global _lib
    if _lib is None:
        import shutil

        if not os.environ.get("GCMESH_LIB") and (not os.path.exists(LIB_PATH) or shutil.which("nvcc")):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.gcmesh_version.restype = ctypes.c_char_p
        L.gcmesh_last_error.restype = ctypes.c_char_p
        L.gcmesh_create.argtypes = [ci, vp, ctypes.POINTER(vp)]
        L.gcmesh_destroy.argtypes = [vp]
        L.gcmesh_set_stream.argtypes = [vp, vp]
        L.gcmesh_set_block_table.argtypes = [vp, ctypes.POINTER(BlockInfo), ci, ci]
        L.gcmesh_default_block_table.argtypes = [ctypes.POINTER(BlockInfo), ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.gcmesh_synchronize.argtypes = [vp]
        L.gcmesh_world_create.argtypes = [vp, ci, ci, ctypes.POINTER(vp)]
        L.gcmesh_world_destroy.argtypes = [vp]
        L.gcmesh_world_upload_chunk.argtypes = [vp, ci, ci, ci, vp, vp, vp]
        L.gcmesh_world_upload_surface.argtypes = [vp, ci, ci, vp]
        L.gcmesh_world_upload_all.argtypes = [vp, vp, vp, vp, vp]
        L.gcmesh_world_upload_rows.argtypes = [vp, ci, ci, vp, vp, vp, vp]
        L.gcmesh_world_upload_sparse.argtypes = [vp, vp, vp, vp, vp, vp, ci]
        L.gcmesh_world_set_option.argtypes = [vp, ci, ci]
        L.gcmesh_world_set_origin.argtypes = [vp, ci, ci]
        L.gcmesh_query.argtypes = [vp, ctypes.POINTER(ci)]
        L.gcmesh_world_validate_blocks.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), vp, ctypes.POINTER(ci)]
        L.gcmesh_world_upload_launches.argtypes = [vp]
        L.gcmesh_world_upload_bytes.argtypes = [vp]
        L.gcmesh_world_upload_bytes.restype = ctypes.c_uint64
        L.gcmesh_world_device_ptrs.argtypes = [vp] + [ctypes.POINTER(vp)] * 4
        L.gcmesh_world_halo_bytes.argtypes = [vp]
        L.gcmesh_world_halo_bytes.restype = ctypes.c_size_t
        L.gcmesh_world_pack_halo.argtypes = [vp, ci, ci, vp]
        L.gcmesh_world_unpack_halo.argtypes = [vp, ci, ci, vp]
        L.gcmesh_world_halo_export.argtypes = [vp, vp]
        L.gcmesh_world_halo_connect_ipc.argtypes = [vp, ci, vp, ci]
        L.gcmesh_world_halo_connect_local.argtypes = [vp, ci, vp, ci]
        L.gcmesh_world_halo_push.argtypes = [vp, ci, ci]
        L.gcmesh_world_halo_status.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ctypes.c_uint32)]
        L.gcmesh_default_light_rules.argtypes = [vp, vp]
        L.gcmesh_set_light_rules.argtypes = [vp, vp, vp]
        L.gcmesh_world_compute_surface.argtypes = [vp]
        L.gcmesh_world_light.argtypes = [vp]
        L.gcmesh_world_light_stats.argtypes = [vp, vp, ctypes.POINTER(ci)]
        L.gcmesh_world_download.argtypes = [vp, vp, vp, vp, vp]
        L.gcmesh_world_upload_rle.argtypes = [vp, vp, ci, vp, ctypes.c_size_t]
        L.gcmesh_world_rle_refused.argtypes = [vp, ctypes.POINTER(ci)]
        L.gcmesh_world_encode_rle.argtypes = [vp, vp, ci, vp, ctypes.c_size_t, vp]
        L.gcmesh_world_set_blocks.argtypes = [vp, vp, ci, vp, vp, ci, ctypes.POINTER(ci)]
        L.gcmesh_world_vertex_counts.argtypes = [vp, vp]
        L.gcmesh_world_generate.argtypes = [vp, ci, ctypes.c_uint32, ci, ci, ci, ci]
        L.gcmesh_batch_create.argtypes = [vp, ci, ctypes.c_uint64, ctypes.POINTER(vp)]
        L.gcmesh_batch_create_ex.argtypes = [vp, ci, ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(vp)]
        L.gcmesh_batch_destroy.argtypes = [vp]
        L.gcmesh_submit.argtypes = [vp, vp, vp, ci]
        L.gcmesh_submit_exchange.argtypes = [vp, vp, vp, ci, ci, ci]
        L.gcmesh_batch_set_timing.argtypes = [vp, ci]
        L.gcmesh_wait.argtypes = [vp]
        L.gcmesh_mesh_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, vp, vp, ctypes.c_uint64, ci]
        L.gcmesh_batch_results.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ci), ctypes.POINTER(ctypes.c_uint64)]
        L.gcmesh_batch_download.argtypes = [vp, vp, vp]
        L.gcmesh_batch_fill_meshdata.argtypes = [vp, ci, vp]
        L.gcmesh_batch_fill_device.argtypes = [vp, vp, ci]
        L.gcmesh_batch_device_ptrs.argtypes = [vp] + [ctypes.POINTER(vp)] * 3
        L.gcmesh_batch_elapsed_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
        L.gcmesh_batch_launches.argtypes = [vp]
        L.gcmesh_kernel_info.argtypes = [ctypes.POINTER(ci)] * 3
        L.gcmesh_batch_profile.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
        _lib = L
    return _lib


def _check(status: int, what: str) -> None:
    if status != 0:
        raise GcMeshError("%s failed (%d): %s" % (what, status, lib().gcmesh_last_error().decode()), status)


def _ptr(a) -> ctypes.c_void_p:
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return ctypes.c_void_p(a.ctypes.data)
    return ctypes.c_void_p(int(a))


def default_light_rules():
    step, emitted = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
    _check(lib().gcmesh_default_light_rules(_ptr(step), _ptr(emitted)), "gcmesh_default_light_rules")
    return step, emitted


def default_block_table():
    """The reference's 24 block types (Code/Block.cpp:139-276) as (list[BlockInfo], kill_zone_id)."""
    arr = (BlockInfo * 256)()
    n, kz = ctypes.c_int(0), ctypes.c_int(0)
    _check(lib().gcmesh_default_block_table(arr, ctypes.byref(n), ctypes.byref(kz)), "gcmesh_default_block_table")
    return [arr[i] for i in range(n.value)], kz.value


class Context:
    """One GPU + one stream (``GameState``'s role for this path).  ``stream`` is a raw ``cudaStream_t`` value."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.h = ctypes.c_void_p()
        _check(lib().gcmesh_create(int(device), ctypes.c_void_p(stream or 0), ctypes.byref(self.h)), "gcmesh_create")
        self.device = device
        self._children = weakref.WeakSet()   # worlds and batches: destroyed before the context, whatever order the GC picks

    def set_block_table(self, infos, kill_zone_id: int = 20) -> None:
        arr = (BlockInfo * len(infos))(*infos)
        _check(lib().gcmesh_set_block_table(self.h, arr, len(infos), kill_zone_id), "gcmesh_set_block_table")

    def set_stream(self, stream: int) -> None:
        _check(lib().gcmesh_set_stream(self.h, ctypes.c_void_p(stream)), "gcmesh_set_stream")

    def set_light_rules(self, step, emitted) -> None:
        """lightStep (1..15, 255 = opaque) and lightEmitted (0..15) per block id, 256 entries each (Code/Block.h:94-95)."""
        step, emitted = np.ascontiguousarray(step, np.uint8), np.ascontiguousarray(emitted, np.uint8)
        assert step.shape == (256,) and emitted.shape == (256,)
        _check(lib().gcmesh_set_light_rules(self.h, _ptr(step), _ptr(emitted)), "gcmesh_set_light_rules")

    def synchronize(self) -> None:
        _check(lib().gcmesh_synchronize(self.h), "gcmesh_synchronize")

    def kernel_info(self):
        r, s, t = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        _check(lib().gcmesh_kernel_info(ctypes.byref(r), ctypes.byref(s), ctypes.byref(t)), "gcmesh_kernel_info")
        return {"registers": r.value, "smem_bytes": s.value, "threads": t.value}

    def close(self) -> None:
        if self.h:
            for child in list(self._children):
                child.close()
            lib().gcmesh_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class World:
    """Device mirror of the reference's window of columns (``World::groups``, Code/World.h:131)."""

    def __init__(self, ctx: Context, size_x: int, size_z: int):
        self.ctx, self.size_x, self.size_z = ctx, int(size_x), int(size_z)
        self.h = ctypes.c_void_p()
        _check(lib().gcmesh_world_create(ctx.h, self.size_x, self.size_z, ctypes.byref(self.h)), "gcmesh_world_create")
        ctx._children.add(self)

    def upload(self, host) -> "World":
        """All columns of a HostWorld-shaped object (``blocks``/``sun``/``light``/``surface`` arrays)."""
        assert host.size_x == self.size_x and host.size_z == self.size_z
        _check(lib().gcmesh_world_upload_all(self.h, _ptr(host.blocks), _ptr(host.sun), _ptr(host.light), _ptr(host.surface)),
               "gcmesh_world_upload_all")
        return self

    def upload_sparse(self, host, nonzero_hint=None, threads: int = 0) -> "World":
        """Like :meth:`upload`, moving only the brick arrays that are not all zero (host scan, or the caller's hint)."""
        assert host.size_x == self.size_x and host.size_z == self.size_z
        _check(lib().gcmesh_world_upload_sparse(self.h, _ptr(host.blocks), _ptr(host.sun), _ptr(host.light), _ptr(host.surface),
                                                _ptr(nonzero_hint), threads), "gcmesh_world_upload_sparse")
        return self

    def upload_launches(self) -> int:
        return int(lib().gcmesh_world_upload_launches(self.h))

    def set_option(self, option: int, value: int = 1) -> "World":
        _check(lib().gcmesh_world_set_option(self.h, int(option), int(value)), "gcmesh_world_set_option")
        return self

    def set_origin(self, origin_cx: int, origin_cz: int) -> "World":
        """World-space column of window column (0, 0) (ChunkGroup::pos): region records carry it in their first word."""
        _check(lib().gcmesh_world_set_origin(self.h, int(origin_cx), int(origin_cz)), "gcmesh_world_set_origin")
        return self

    def validate_blocks(self):
        """(number of block ids >= the block table's size, (cx, cy, cz) of the first one's chunk, its voxel index) — the
        reference's ``assert(block < BLOCK_COUNT)`` (Code/World.cpp:294) as a scan of the device window."""
        n = ctypes.c_uint64(0)
        c = np.zeros(3, np.int32)
        v = ctypes.c_int(-1)
        _check(lib().gcmesh_world_validate_blocks(self.h, ctypes.byref(n), _ptr(c), ctypes.byref(v)), "gcmesh_world_validate_blocks")
        return int(n.value), (tuple(int(x) for x in c) if n.value else None), (int(v.value) if n.value else None)

    def upload_bytes(self) -> int:
        """Host -> device bytes the last sparse upload / mesh_host call moved."""
        return int(lib().gcmesh_world_upload_bytes(self.h))

    def upload_rows(self, host, cz0: int, cz1: int) -> "World":
        _check(lib().gcmesh_world_upload_rows(self.h, cz0, cz1, _ptr(host.blocks), _ptr(host.sun), _ptr(host.light), _ptr(host.surface)),
               "gcmesh_world_upload_rows")
        return self

    def upload_chunk(self, cx, cy, cz, blocks=None, sunlight=None, block_light=None) -> None:
        _check(lib().gcmesh_world_upload_chunk(self.h, cx, cy, cz, _ptr(blocks), _ptr(sunlight), _ptr(block_light)), "gcmesh_world_upload_chunk")

    def upload_surface(self, cx, cz, surface) -> None:
        _check(lib().gcmesh_world_upload_surface(self.h, cx, cz, _ptr(surface)), "gcmesh_world_upload_surface")

    def device_ptrs(self):
        p = [ctypes.c_void_p() for _ in range(4)]
        _check(lib().gcmesh_world_device_ptrs(self.h, *[ctypes.byref(x) for x in p]), "gcmesh_world_device_ptrs")
        return tuple(x.value for x in p)

    def halo_bytes(self) -> int:
        return int(lib().gcmesh_world_halo_bytes(self.h))

    # -- column pre-processing on the device (ComputeSurface / SetLightNodes, Code/Lighting.cpp) --
    def compute_surface(self) -> "World":
        _check(lib().gcmesh_world_compute_surface(self.h), "gcmesh_world_compute_surface")
        return self

    def light(self) -> "World":
        """Clear and fill the sunlight / blockLight arenas from the block ids and the surface map."""
        _check(lib().gcmesh_world_light(self.h), "gcmesh_world_light")
        return self

    def light_stats(self):
        st, n = np.zeros(4, np.uint32), ctypes.c_int(0)
        _check(lib().gcmesh_world_light_stats(self.h, _ptr(st), ctypes.byref(n)), "gcmesh_world_light_stats")
        return {"sun_candidates": int(st[0]), "brick_visits": int(st[1]), "sweeps": int(st[2]), "seeds": int(st[3]), "launches": n.value}

    # -- region records (the reference's saved-chunk format, Code/WorldIO.cpp:167-307) --
    def upload_rle(self, chunks: np.ndarray, data: np.ndarray) -> "World":
        """``chunks``: structured array (RLE_CHUNK_DTYPE) locating each record in the uint16 stream ``data``."""
        chunks = np.ascontiguousarray(chunks, RLE_CHUNK_DTYPE)
        self._rle_keep = (chunks, np.ascontiguousarray(data, np.uint16))          # asynchronous: keep the stream alive
        _check(lib().gcmesh_world_upload_rle(self.h, _ptr(chunks), len(chunks), _ptr(self._rle_keep[1]), len(self._rle_keep[1])),
               "gcmesh_world_upload_rle")
        return self

    def rle_refused(self) -> int:
        n = ctypes.c_int(0)
        _check(lib().gcmesh_world_rle_refused(self.h, ctypes.byref(n)), "gcmesh_world_rle_refused")
        return n.value

    def encode_rle(self, coords, capacity_words: int | None = None):
        """Device bricks -> (chunks locating array, uint16 record stream)."""
        coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
        cap = capacity_words or len(coords) * (2 + 2 * CHUNK_VOXELS)
        data = np.zeros(cap, np.uint16)
        out = np.zeros(len(coords), RLE_CHUNK_DTYPE)
        _check(lib().gcmesh_world_encode_rle(self.h, _ptr(coords), len(coords), _ptr(data), cap, _ptr(out)), "gcmesh_world_encode_rle")
        used = int(out["first_word"][-1] + out["words"][-1]) if len(out) else 0
        return out, data[:used]

    def generate(self, kind, seed: int = 1337, origin=(0, 0), radius: int = 2**31 - 1, falloff: int | None = None) -> "World":
        """Terrain of every column of the window, generated on the device: the reference's nine biomes ("flat", "forest",
        "islands", "grid", "void", "snow", "desert", "volcanic", "dungeon"; Code/Environment.cpp:83-143)."""
        k = GEN_KINDS[kind] if isinstance(kind, str) else int(kind)
        if falloff is None:
            falloff = radius - 64                                                   # World.cpp:911
        _check(lib().gcmesh_world_generate(self.h, k, seed & 0xFFFFFFFF, int(origin[0]), int(origin[1]), int(radius), int(falloff)),
               "gcmesh_world_generate")
        return self

    def set_blocks(self, edits, rebuild_capacity: int | None = None):
        """SetBlock for each (wx, wy, wz, block) in order on the device window -> (status per edit, (n, 3) chunks to rebuild)."""
        rows = [tuple(int(v) for v in e) + (0,) for e in edits]
        ed = np.array(rows, BLOCK_EDIT_DTYPE) if rows else np.zeros(0, BLOCK_EDIT_DTYPE)
        status = np.full(len(ed), -1, np.int32)
        cap = self.size_x * self.size_z * 4 if rebuild_capacity is None else int(rebuild_capacity)
        rebuild = np.zeros((max(cap, 1), 3), np.int32)
        n = ctypes.c_int(0)
        _check(lib().gcmesh_world_set_blocks(self.h, _ptr(ed) if len(ed) else None, len(ed), _ptr(status) if len(ed) else None,
                                             _ptr(rebuild) if cap else None, cap, ctypes.byref(n)), "gcmesh_world_set_blocks")
        return status, rebuild[: n.value].copy()

    def vertex_counts(self) -> np.ndarray:
        """chunk->totalVertices per brick, shaped (size_z, size_x, 4); -1 = never meshed."""
        out = np.zeros(self.size_x * self.size_z * 4, np.int32)
        _check(lib().gcmesh_world_vertex_counts(self.h, _ptr(out)), "gcmesh_world_vertex_counts")
        return out.reshape(self.size_z, self.size_x, 4)

    def download(self, host, blocks: bool = False, sun: bool = True, light: bool = True, surface: bool = True) -> None:
        """Device -> host copy into a HostWorld-shaped object's arrays."""
        _check(lib().gcmesh_world_download(self.h, _ptr(host.blocks) if blocks else None, _ptr(host.sun) if sun else None,
                                           _ptr(host.light) if light else None, _ptr(host.surface) if surface else None), "gcmesh_world_download")

    def pack_halo(self, cz: int, z: int, device_buf: int) -> None:
        _check(lib().gcmesh_world_pack_halo(self.h, cz, z, ctypes.c_void_p(device_buf)), "gcmesh_world_pack_halo")

    def unpack_halo(self, cz: int, z: int, device_buf: int) -> None:
        _check(lib().gcmesh_world_unpack_halo(self.h, cz, z, ctypes.c_void_p(device_buf)), "gcmesh_world_unpack_halo")

    # -- peer-memory halo exchange (side 0: the rank below, 1: the rank above) --
    def halo_export(self) -> bytes:
        buf = ctypes.create_string_buffer(HALO_HANDLE_BYTES)
        _check(lib().gcmesh_world_halo_export(self.h, buf), "gcmesh_world_halo_export")
        return buf.raw

    def halo_connect_ipc(self, side: int, handles: bytes, peer_row: int) -> None:
        assert len(handles) == HALO_HANDLE_BYTES
        _check(lib().gcmesh_world_halo_connect_ipc(self.h, side, ctypes.c_char_p(handles), peer_row), "gcmesh_world_halo_connect_ipc")

    def halo_connect_local(self, side: int, peer: "World", peer_row: int) -> None:
        _check(lib().gcmesh_world_halo_connect_local(self.h, side, peer.h, peer_row), "gcmesh_world_halo_connect_local")

    def halo_push(self, send_row_down: int, send_row_up: int) -> None:
        _check(lib().gcmesh_world_halo_push(self.h, send_row_down, send_row_up), "gcmesh_world_halo_push")

    def halo_status(self):
        """(timed_out, exchanges so far); synchronises the context stream."""
        t, n = ctypes.c_int(0), ctypes.c_uint32(0)
        _check(lib().gcmesh_world_halo_status(self.h, ctypes.byref(t), ctypes.byref(n)), "gcmesh_world_halo_status")
        return int(t.value), int(n.value)

    def close(self) -> None:
        if self.h:
            lib().gcmesh_world_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ChunkMesh:
    """What BuildChunkAsync leaves in ``chunk->meshData`` (Code/Mesh.h:46-52): (n,12) u8 vertices + 5 u16 index arrays."""

    def __init__(self, vertices, indices, flags=0, quads=0):
        self.vertices, self.indices, self.flags, self.quads = vertices, indices, int(flags), int(quads)

    @property
    def vert_count(self) -> int:
        return len(self.vertices)

    @property
    def overflow(self) -> bool:
        return bool(self.flags & CHUNK_OVERFLOW)


class Batch:
    """Output arenas + descriptors of one batched submission (the reference's ``vector<Chunk*>`` job)."""

    def __init__(self, ctx: Context, max_chunks: int, max_quads: int, uncapped: bool = False):
        """``uncapped``: GC_BATCH_UNCAPPED — no 10 000-quad cap per chunk, uint32 indices (the "stress" form; not a reference behaviour)."""
        self.ctx, self.max_chunks, self.max_quads, self.uncapped = ctx, int(max_chunks), int(max_quads), bool(uncapped)
        self.index_dtype = np.uint32 if uncapped else np.uint16
        self.h = ctypes.c_void_p()
        _check(lib().gcmesh_batch_create_ex(ctx.h, self.max_chunks, self.max_quads, BATCH_UNCAPPED if uncapped else 0, ctypes.byref(self.h)), "gcmesh_batch_create_ex")
        ctx._children.add(self)
        self._coords = None

    def submit(self, world: World, coords) -> "Batch":
        self._coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
        _check(lib().gcmesh_submit(world.h, self.h, _ptr(self._coords), len(self._coords)), "gcmesh_submit")
        return self

    def submit_exchange(self, world: World, coords, send_row_down: int, send_row_up: int) -> "Batch":
        """Slab halo exchange + mesh build in one kernel (gcmesh_submit_exchange)."""
        self._coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
        _check(lib().gcmesh_submit_exchange(world.h, self.h, _ptr(self._coords), len(self._coords), int(send_row_down), int(send_row_up)),
               "gcmesh_submit_exchange")
        return self

    def set_timing(self, on: bool) -> "Batch":
        """Record (default) or leave out the two timing events around the kernel (:meth:`elapsed_ms` needs them)."""
        _check(lib().gcmesh_batch_set_timing(self.h, 1 if on else 0), "gcmesh_batch_set_timing")
        return self

    def mesh_host(self, world: World, host, coords, vertices, indices, nonzero_hint=None, threads: int = 0) -> "Batch":
        """Host arrays in, host meshes out (gcmesh_mesh_host): ``vertices`` (q*4,12) u8 and ``indices`` (q*6,) u16 receive the
        arena prefix; descriptors via :meth:`results`."""
        self._coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
        cap = min(len(vertices) // 4, len(indices) // 6)
        _check(lib().gcmesh_mesh_host(world.h, self.h, _ptr(host.blocks), _ptr(host.sun), _ptr(host.light), _ptr(host.surface),
                                      _ptr(nonzero_hint), _ptr(self._coords), len(self._coords), _ptr(vertices), _ptr(indices), cap, threads),
               "gcmesh_mesh_host")
        return self

    def wait(self) -> "Batch":
        _check(lib().gcmesh_wait(self.h), "gcmesh_wait")
        return self

    def done(self) -> bool:
        """gcmesh_query: has the last submission finished?  (Does not block.)"""
        d = ctypes.c_int(0)
        _check(lib().gcmesh_query(self.h, ctypes.byref(d)), "gcmesh_query")
        return bool(d.value)

    def results(self):
        """(descriptors as a structured numpy array view, total_quads)."""
        p, n, q = ctypes.c_void_p(), ctypes.c_int(0), ctypes.c_uint64(0)
        _check(lib().gcmesh_batch_results(self.h, ctypes.byref(p), ctypes.byref(n), ctypes.byref(q)), "gcmesh_batch_results")
        if n.value == 0:
            return np.zeros(0, CHUNK_OUT_DTYPE), 0
        buf = (ctypes.c_uint8 * (n.value * 64)).from_address(p.value)
        return np.frombuffer(buf, CHUNK_OUT_DTYPE, n.value), int(q.value)

    def download(self, vertices=None, indices=None):
        """Copy the used arena prefix to host arrays: (total_quads*4, 12) u8 and (total_quads*6,) u16."""
        _, q = self.results()
        if vertices is None:
            vertices = np.empty((q * 4, 12), np.uint8)
        if indices is None:
            indices = np.empty(q * 6, self.index_dtype)
        _check(lib().gcmesh_batch_download(self.h, _ptr(vertices), _ptr(indices)), "gcmesh_batch_download")
        return vertices, indices

    def meshes(self):
        """One :class:`ChunkMesh` per submitted chunk, in submit order."""
        desc, q = self.results()
        verts, idx = self.download()
        out = []
        for d in desc:
            qb, nv = int(d["quad_base"]), int(d["vert_count"])
            v = verts[qb * 4: qb * 4 + nv]
            ii = [idx[qb * 6 + int(d["index_offset"][t]): qb * 6 + int(d["index_offset"][t]) + int(d["index_count"][t])] for t in range(MESH_TYPES)]
            out.append(ChunkMesh(v, ii, d["flags"], d["quads"]))
        return out

    def fill_device(self, targets) -> "Batch":
        """FillMeshData onto device buffers: ``targets`` = [(chunk index, vertices device pointer or 0, five index-stream
        device pointers or 0), ...] (e.g. graphics buffers mapped into CUDA).  Asynchronous on the context stream."""
        if isinstance(targets, np.ndarray) and targets.dtype == DEVICE_MESH_DTYPE:
            arr = np.ascontiguousarray(targets)                                     # prepared once, reused every frame
        else:
            arr = np.zeros(len(targets), DEVICE_MESH_DTYPE)
            for k, (chunk, v, idx) in enumerate(targets):
                arr[k]["chunk"] = chunk
                arr[k]["vertices"] = v or 0
                arr[k]["indices"] = [p or 0 for p in idx]
        _check(lib().gcmesh_batch_fill_device(self.h, _ptr(arr) if len(arr) else None, len(arr)), "gcmesh_batch_fill_device")
        return self

    def device_ptrs(self):
        p = [ctypes.c_void_p() for _ in range(3)]
        _check(lib().gcmesh_batch_device_ptrs(self.h, *[ctypes.byref(x) for x in p]), "gcmesh_batch_device_ptrs")
        return tuple(x.value for x in p)

    def elapsed_ms(self) -> float:
        ms = ctypes.c_float(0)
        _check(lib().gcmesh_batch_elapsed_ms(self.h, ctypes.byref(ms)), "gcmesh_batch_elapsed_ms")
        return float(ms.value)

    def profile(self):
        out = (ctypes.c_uint64 * 512)()
        _check(lib().gcmesh_batch_profile(self.h, out), "gcmesh_batch_profile")
        return [int(v) for v in out]

    def launches(self) -> int:
        return int(lib().gcmesh_batch_launches(self.h))

    def close(self) -> None:
        if self.h:
            lib().gcmesh_batch_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rebuild_chunks(world: World, coords, batch: Batch | None = None):
    """``RebuildChunks`` (Code/WorldRender.cpp:80-91): mesh a list of chunks in ONE submission; returns ChunkMesh list."""
    coords = np.ascontiguousarray(coords, np.int32).reshape(-1, 3)
    own = batch is None
    if own:
        batch = Batch(world.ctx, max(1, len(coords)), max(1, len(coords)) * MAX_QUADS)
    try:
        return batch.submit(world, coords).wait().meshes()
    finally:
        if own:
            batch.close()


def build_chunk(world: World, cx: int, cy: int, cz: int) -> ChunkMesh:
    """``BuildChunk`` (Code/WorldRender.cpp:69-78) for one chunk."""
    return rebuild_chunks(world, [(cx, cy, cz)])[0]

"""The C-ABI library loads and exports every symbol include/gcmesh.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from gamecraft_b200 import mesher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "gcmesh.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcmesh_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    mesher.build()
    L = ctypes.CDLL(mesher.LIB_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libgcmesh.so does not export %s" % n
    assert sorted(mesher.EXPORTS) == names, "mesher.EXPORTS and include/gcmesh.h disagree"


def test_struct_layouts_match_header():
    assert ctypes.sizeof(mesher.ChunkOut) == 64
    assert ctypes.sizeof(mesher.BlockInfo) == 12


def test_default_block_table_is_the_reference_table():
    from oracle import binding as ob

    infos, kz = mesher.default_block_table()
    tab, n, okz = ob.default_block_table()
    assert len(infos) == n == 24 and kz == okz == 20
    for a, b in zip(infos, tab):
        assert bytes(a) == bytes(b)


def test_no_gpu_fails_loudly():
    """There is no CPU path: without a usable device, context creation is an error, not a fallback."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mesher.GcMeshError) as e:
        mesher.Context(0)
    assert "no usable CUDA device" in str(e.value) or "-2" in str(e.value)


def test_argument_errors_without_gpu():
    L = mesher.lib()
    assert L.gcmesh_world_create(None, 4, 4, ctypes.byref(ctypes.c_void_p())) == -1
    assert L.gcmesh_wait(None) == -1
    assert b"null" in L.gcmesh_last_error()
    # the pre-processing and region-record entry points refuse null handles the same way (no device is touched)
    assert L.gcmesh_world_compute_surface(None) == -1
    assert L.gcmesh_world_light(None) == -1
    assert L.gcmesh_set_light_rules(None, None, None) == -1
    assert L.gcmesh_world_upload_rle(None, None, 0, None, 0) == -1
    assert L.gcmesh_world_encode_rle(None, None, 0, None, 0, None) == -1
    n = ctypes.c_int(7)
    assert L.gcmesh_world_set_blocks(None, None, 0, None, None, 0, ctypes.byref(n)) == -1 and n.value == 0
    assert L.gcmesh_world_vertex_counts(None, None) == -1
    assert ctypes.sizeof(ctypes.c_int32) * 4 == mesher.BLOCK_EDIT_DTYPE.itemsize          # GcBlockEdit: three int32 + two uint16
    step, emitted = mesher.default_light_rules()
    assert step[0] == 1 and step[3] == 255 and emitted[14] == 15 and emitted[17] == 10

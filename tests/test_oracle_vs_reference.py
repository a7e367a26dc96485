"""Plain-C restatement (oracle/gc_mesh_cpu.c) vs the reference's own sources compiled verbatim (oracle/_ref)."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util

pytestmark = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")


def test_struct_sizes_and_caps():
    L = ob.ref_lib()
    # VertexInfo 12, MeshData 480048, MeshIndexData 120004, Chunk 262248, ChunkGroup 1050040
    assert [L.gcref_sizeof(i) for i in range(5)] == [12, 480048, 120004, 262248, 1050040]
    assert [L.gcref_sizeof(i) for i in range(6, 11)] == [0, 3, 6, 10, 11]   # offsets pos, uv, color, alpha, reserved
    assert [L.gcref_sizeof(i) for i in (11, 12, 13)] == [40000, 60000, 24]


def test_block_table_matches_reference():
    r = ob.RefWorld(size=3)
    ref = r.block_table()
    tab, n, kz = ob.default_block_table()
    assert n == 24 and kz == 20
    for b in range(24):
        mine = list(tab[b].textures) + [tab[b].mesh_type, tab[b].cull, tab[b].alpha, tab[b].invisible, tab[b].opaque]
        assert mine == ref[b][:11].tolist(), "block %d" % b


@pytest.mark.parametrize("kind,size,seed", [("mixed", 5, 21), ("islands", 5, 8), ("forest", 5, 4), ("noise", 3, 5), ("checker", 3, 6)])
def test_port_equals_reference(kind, size, seed):
    kw = dict(radius=90, falloff=50) if kind in ("mixed", "islands") else {}
    w = util.HostWorld(size, size, origin=(-2, -2) if kw else (0, 0)).build(kind, seed, threads=1, **kw)
    o, r = ob.OracleWorld(w), ob.RefWorld(w)
    for cx, cy, cz in w.interior_chunks(1):
        a, b = o.build_chunk(cx, cy, cz), r.build_chunk(cx, cy, cz)
        util.assert_same_mesh(a, b, "%s %s" % (kind, (cx, cy, cz)))
        assert a.allocated == b.allocated


@pytest.mark.parametrize("kind,seed", [("mixed", 3), ("noise", 4), ("forest", 5), ("checker", 6)])
def test_light_fill_equals_reference(kind, seed):
    """gcworld's ComputeSurface + SetLightNodes restatement vs the reference's, same blocks, same column order."""
    kw = dict(radius=90, falloff=50) if kind == "mixed" else {}
    w = util.HostWorld(5, 5, origin=(-2, -2) if kw else (0, 0)).build(kind, seed, threads=1, **kw)
    z = util.HostWorld(5, 5)
    z.blocks[:] = w.blocks
    r = ob.RefWorld(z)
    r.compute_surface(); r.compute_light(); r.download(z)
    assert np.array_equal(z.surface, w.surface)
    assert np.array_equal(z.sun, w.sun)
    assert np.array_equal(z.light, w.light)


def test_chunk_overflowed_precheck():
    """ChunkOverflowed (WorldRender.cpp:412-439)."""
    w = util.cap_edge_world(pairs=0, singles=10)
    o, r = ob.OracleWorld(w), ob.RefWorld(w)
    for total in (0, 39975, 39976, 39977, 40000):
        for (x, y, z) in ((1, 1, 1), (2, 1, 1), (5, 1, 1)):
            assert o.chunk_overflowed(1, 1, 1, x, y, z, total) == r.chunk_overflowed(1, 1, 1, x, y, z, total)


def test_cap_edge_exactly_10000_quads():
    w = util.cap_edge_world(pairs=1, singles=1665)
    a, b = ob.OracleWorld(w).build_chunk(1, 1, 1), ob.RefWorld(w).build_chunk(1, 1, 1)
    assert a.vert_count == 40000 and not a.overflow
    util.assert_same_mesh(a, b)

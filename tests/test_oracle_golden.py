"""The oracle is pinned: golden vectors minted from the reference's own mesher compiled verbatim (oracle/_ref),
plus the known answers of SURVEY.md A.9."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util

GOLDEN = util.load_golden()


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_port_matches_golden(key):
    entry = GOLDEN[key]
    w = util.golden_world(entry)
    assert util.input_digest(w) == entry["inputs"], "world producer gave different inputs on this host"
    o = ob.OracleWorld(w)
    for name, want in entry["chunks"].items():
        cx, cy, cz = map(int, name.split(","))
        assert util.mesh_signature(o.build_chunk(cx, cy, cz)) == want, "%s chunk %s" % (key, name)


def test_flat_world_known_answers():
    """SURVEY.md A.9 (from the verbatim reference build): flat world, interior chunk with cy = 0."""
    w = util.HostWorld(5, 5).build("flat")
    m = ob.OracleWorld(w).build_chunk(2, 0, 2)
    assert m.vert_count == 8192 and [len(i) for i in m.indices] == [12288, 0, 0, 0, 0]
    assert m.allocated == [1, 0, 0, 0, 0]
    assert m.vertices[0].tolist() == [0, 0, 0, 0, 1, 6, 0, 0, 0, 0, 255, 0]
    assert m.vertices[1].tolist() == [0, 0, 1, 0, 0, 6, 0, 0, 0, 0, 255, 0]
    assert m.vertices[4096].tolist() == [1, 13, 0, 0, 1, 8, 0, 0, 0, 255, 255, 0]
    assert m.indices[0][:12].tolist() == [2, 1, 0, 3, 2, 0, 6, 5, 4, 7, 6, 4]
    assert ob.fnv1a64(m.vertices) == "268aab0869012b25"
    assert ob.fnv1a64(m.indices[0]) == "0b52aef0031ed325"


@pytest.mark.parametrize("name", ["flat", "mixed"])
def test_full_output_fixture(name):
    import os
    z = np.load(os.path.join(util.GOLDEN, "golden_%s_chunk.npz" % name))
    key = "flat5" if name == "flat" else "mixed7"
    w = util.golden_world(GOLDEN[key])
    cx, cy, cz = z["coord"].tolist()
    m = ob.OracleWorld(w).build_chunk(cx, cy, cz)
    assert np.array_equal(m.vertices, z["vertices"])
    for t in range(5):
        assert np.array_equal(m.indices[t], z["indices%d" % t])


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")
@pytest.mark.parametrize("key", ["flat5", "mixed7"])
def test_verbatim_reference_matches_golden(key):
    entry = GOLDEN[key]
    w = util.golden_world(entry)
    ref = ob.RefWorld(w)
    for name, want in entry["chunks"].items():
        cx, cy, cz = map(int, name.split(","))
        assert util.mesh_signature(ref.build_chunk(cx, cy, cz)) == want

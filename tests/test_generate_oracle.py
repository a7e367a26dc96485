"""The generators of the reference that need no noise library (GenerateFlatTerrain, GenerateGridTerrain, GenerateVoidTerrain,
Code/Generation.cpp:772-836, :371-410, :838-852) run from the reference's own sources (oracle/_ref): the host input
producer's flat rule is pinned against them here, the device generators in tests/test_gpu_generate.py."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util

needs_ref = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")


def reference_blocks(kind, size, radius, falloff):
    """Block ids of a size x size window at the world origin from the reference's own generator."""
    out = util.HostWorld(size, size)
    r = ob.RefWorld(size=size, radius=radius, falloff=falloff)
    r.generate(kind)
    r.download(out)
    r.close()
    return out.blocks


@needs_ref
@pytest.mark.parametrize("radius,falloff", [(2**31 - 1, 2**31 - 1 - 64), (100, 36), (70, 40), (33, 1)])
def test_host_flat_rule_equals_reference(radius, falloff):
    size = 5
    want = reference_blocks("flat", size, radius, falloff)
    got = util.HostWorld(size, size).generate("flat", 1, radius=radius, falloff=falloff)
    assert np.array_equal(got.blocks, want)
    if radius < 2**31 - 1:
        assert {0, 1, 2, 8} <= set(np.unique(want).tolist())      # air, grass, dirt and the sea around the island


@needs_ref
def test_reference_grid_and_void_shapes():
    """What the device generators are compared with: the mast's top crate lands at (16, 0, 17) (chunk-local y = 64 carried
    by BlockIndex), the void world is one grass plate."""
    grid = reference_blocks("grid", 3, 50, 0)
    assert set(np.unique(grid).tolist()) == {0, 7}
    for cy in range(4):
        chunk = grid[4 * 4 + cy].reshape(32, 64, 32)                 # column (1, 1): [z][y][x]
        assert chunk[17, 0, 16] == 7 and chunk[16, 62, 16] == 7 and chunk[16, 63, 16] == 0
    void = reference_blocks("void", 3, 50, 0)
    assert int((void != 0).sum()) == 1024 and (void[0].reshape(32, 64, 32)[:, 40, :] == 1).all()

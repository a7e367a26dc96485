"""The generators of the reference that need no noise library (GenerateFlatTerrain, GenerateGridTerrain, GenerateVoidTerrain,
Code/Generation.cpp:772-836, :371-410, :838-852; GenerateDungeon, Code/Dungeon.cpp:17-31) run from the reference's own
sources (oracle/_ref): the host input producer's flat and dungeon rules are pinned against them here, the device generators
in tests/test_gpu_generate.py.  The noise terrains' layering rules (snow, desert, volcanic) are checked for the properties the
reference's code implies."""
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


@needs_ref
def test_host_dungeon_rule_equals_reference():
    """GenerateDungeon: a stone floor and eleven-high walls along the border of every column's lowest chunk."""
    size = 4
    want = reference_blocks("dungeon", size, 50, 0)
    got = util.HostWorld(size, size).generate("dungeon", 1)
    assert np.array_equal(got.blocks, want)
    chunk = want[(1 * size + 2) * 4].reshape(32, 64, 32)             # [z][y][x]
    assert (chunk[:, 0, :] == 3).all() and (chunk[0, :11, :] == 3).all() and (chunk[:, :11, 31] == 3).all()
    assert not chunk[1:31, 1:, 1:31].any() and not chunk[:, 11:, :].any()
    assert not want.reshape(size * size, 4, 65536)[:, 1:].any()      # only the lowest chunk of a column is touched


def column_profile(w, cx, cz):
    """(32, 256, 32) [z][y][x] block ids of one column."""
    return np.concatenate([w.blocks[w.slot(cx, cy, cz)].reshape(32, 64, 32) for cy in range(4)], axis=1)


def test_snow_rules():
    """GenerateSnowTerrain (Generation.cpp:412-556): snow on top, dirt below, water under sea level 20, the sea-level layer is
    water or ice wherever the surface is not exactly there — also inside taller ground — and stone only four layers down."""
    w = util.HostWorld(6, 6, origin=(-3, -3)).generate("snow", 11)
    ids = set(np.unique(w.blocks).tolist())
    assert ids <= {0, 2, 3, 8, 12, 13} and {2, 8, 12} <= ids
    saw_ice = saw_buried = False
    for cz in range(6):
        for cx in range(6):
            col = column_profile(w, cx, cz)
            top = (col != 0).cumsum(axis=1).argmax(axis=1)             # highest non-air y per (z, x)
            snow_y = (col == 12).argmax(axis=1)
            assert ((col == 12).sum(axis=1) == 1).all()                # one snow block per cell: the surface
            h = snow_y
            level = col[:, 20, :]
            assert np.isin(level[h != 20], (8, 13)).all()              # y = 20 is water or ice unless it is the surface itself
            saw_buried |= bool((h > 20).any())
            saw_ice |= bool((level == 13).any())
            assert (top == np.maximum(h, 20)).all()
            below = np.arange(256).reshape(1, 256, 1) < h.reshape(32, 1, 32)
            assert np.isin(col[below & (np.arange(256).reshape(1, 256, 1) != 20)], (2, 3)).all()
    assert saw_ice and saw_buried


def test_desert_rules():
    """GenerateDesertTerrain (Generation.cpp:558-645): sand up to the surface, water up to sea level 10, no stone, at most two
    cacti per column of two to four blocks standing on the surface entry."""
    w = util.HostWorld(6, 6, origin=(2, -5)).generate("desert", 5)
    assert set(np.unique(w.blocks).tolist()) <= {0, 5, 8, 16} and 16 in np.unique(w.blocks)
    for cz in range(6):
        for cx in range(6):
            col = column_profile(w, cx, cz)
            cactus = (col == 16)
            stems = cactus.any(axis=1)
            assert stems.sum() <= 2
            for z, x in zip(*np.nonzero(stems)):
                ys = np.nonzero(cactus[z, :, x])[0]
                assert 2 <= len(ys) <= 4 and (np.diff(ys) == 1).all()
                assert col[z, ys[0] - 1, x] in (5, 8, 16) and (col[z, :ys[0], x] != 0).all()


def test_volcanic_rules():
    """GenerateVolcanicTerrain (Generation.cpp:647-770): obsidian with stone pockets, lava up to sea level 12, nothing else."""
    w = util.HostWorld(6, 6, origin=(-1, 7)).generate("volcanic", 9)
    ids = set(np.unique(w.blocks).tolist())
    assert ids <= {0, 3, 19, 21} and {19, 21} <= ids
    col = column_profile(w, 2, 2)
    solid = col != 0
    assert (solid[:, :13, :]).all()                                    # everything up to the lava level is filled
    assert not (col[:, 13:, :] == 21).any()


@needs_ref
@pytest.mark.parametrize("kind,seed,origin,radius,falloff", [
    ("forest", 1337, (-2, -3), 2**31 - 1, None), ("forest", 7, (-2, -2), 75, 30), ("islands", 4242, (-3, -2), 90, 40),
    ("islands", 8, (3, -4), 2**31 - 1, None), ("snow", 11, (-2, -2), 2**31 - 1, None), ("snow", 3, (-3, -2), 85, 30),
    ("desert", 5, (-2, -2), 2**31 - 1, None), ("desert", 6, (-3, -2), 75, 20), ("volcanic", 9, (-2, -3), 80, 25),
    ("volcanic", 2, (-2, -2), 2**31 - 1, None)])
def test_reference_generator_code_over_this_project_s_noise(kind, seed, origin, radius, falloff):
    """The five noise terrains, pinned to the reference's own CODE: GenerateForest / Islands / Snow / Desert / VolcanicTerrain run
    from the reference's sources (oracle/_ref) over this project's noise fields and random stream — gcw_noise plugged into the
    FastNoiseSIMD stand-in, the producer's per-column stream plugged in under rand() (RandRange = min + rand() % n, Random.h:32-35)
    — with the same seed, world position and island, and EVERY block equals the host producer's: heights, the biome blend, the
    island fall-off, sea levels, grass / dirt / snow / sand / obsidian, water / ice / lava, the snow biome's sea-level layer
    inside taller ground, stone pockets (including the reference's reading of its 3-D noise set with a stride of maxY where the
    set has maxY + 1 layers, Generation.cpp:16-18 against :157-161), trees and cacti with the reference's draw order.  What stays
    unpinned is what cannot be pinned here: the values FastNoiseSIMD and MSVC's rand() would have produced."""
    size = 5
    fo = (radius - 64) if falloff is None else falloff
    host = util.HostWorld(size, size, origin=origin).generate(kind, seed, radius=radius, falloff=fo)
    r = ob.RefWorld(size=size, seed=seed, radius=radius, falloff=fo)
    r.set_origin(*origin)
    r.generate_noise_terrain(kind, seed)
    got = util.HostWorld(size, size)
    r.download(got)
    r.close()
    diff = got.blocks != host.blocks
    assert not diff.any(), "%d cells differ from the reference's own %s generator" % (int(diff.sum()), kind)
    assert int((host.blocks != 0).sum()) > 100000                              # the comparison is not vacuous
    ids = set(np.unique(host.blocks).tolist())
    assert {"forest": {1, 2, 9, 10}, "islands": {1, 2, 8}, "snow": {2, 8, 12}, "desert": {5, 16}, "volcanic": {19, 21}}[kind] <= ids
    # with the C library's rand() instead of the producer's stream only the decorations differ
    r = ob.RefWorld(size=size, seed=seed, radius=radius, falloff=fo)
    r.set_origin(*origin)
    r.generate_noise_terrain(kind, seed, own_random=False)
    r.download(got)
    r.close()
    deco = np.isin(got.blocks, (9, 10, 16)) | np.isin(host.blocks, (9, 10, 16))
    assert not ((got.blocks != host.blocks) & ~deco).any()

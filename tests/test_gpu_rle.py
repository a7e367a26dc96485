"""Region records on the device (gcmesh_world_upload_rle / gcmesh_world_encode_rle, through the C ABI) against the
oracle codec (oracle/gc_rle_cpu.c, itself pinned to the reference's SaveGroup / LoadGroupFromDisk)."""
import numpy as np
import pytest

import gcutil as util
from oracle import binding as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gamecraft_b200 import mesher

    c = mesher.Context(0)
    yield c
    c.close()


def region_index(cx, cy, cz):
    return (cx & 7) + 8 * (cy + 4 * (cz & 7))


def host_records(w, coords):
    """the oracle's records of the listed chunks, concatenated, with the locating array"""
    from gamecraft_b200 import mesher

    ox, oz = w.origin
    recs = [ob.rle_encode_chunk(w.blocks[w.slot(*c)], region_index(c[0] + ox, c[1], c[2] + oz)) for c in coords.tolist()]
    loc = np.zeros(len(recs), mesher.RLE_CHUNK_DTYPE)
    p = 0
    for i, (c, r) in enumerate(zip(coords.tolist(), recs)):
        loc[i] = (c[0], c[1], c[2], p, len(r)); p += len(r)
    return loc, (np.concatenate(recs) if recs else np.zeros(0, np.uint16)), recs


def all_chunks(w):
    return np.asarray([(cx, cy, cz) for cz in range(w.size_z) for cx in range(w.size_x) for cy in range(4)], np.int32)


@pytest.mark.parametrize("kind,size,seed", [("forest", 4, 3), ("islands", 4, 9), ("mixed", 4, 12), ("noise", 3, 2), ("checker", 3, 1), ("empty", 3, 0)])
def test_decode_and_encode_equal_oracle(ctx, kind, size, seed):
    from gamecraft_b200 import mesher

    kw = dict(radius=90, falloff=50) if kind in ("islands", "mixed") else {}
    w = util.HostWorld(size, size, origin=(-(size // 2), -(size // 2)) if kw else (0, 0)).generate(kind, seed, **kw)
    coords = all_chunks(w)
    loc, data, recs = host_records(w, coords)
    world = mesher.World(ctx, size, size).set_origin(*w.origin)
    junk = util.HostWorld(size, size); junk.blocks[:] = 0x1234
    world.upload(junk)                                              # the decode has to overwrite every voxel
    world.upload_rle(loc, data)
    assert world.rle_refused() == 0
    got = util.HostWorld(size, size)
    world.download(got, blocks=True, sun=False, light=False, surface=False)
    assert np.array_equal(got.blocks, w.blocks)
    # and back: the device's records are the reference's
    loc2, data2 = world.encode_rle(coords)
    assert np.array_equal(loc2["words"], loc["words"]) and np.array_equal(loc2["first_word"], loc["first_word"])
    assert np.array_equal(data2, data)
    world.close()


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")
def test_records_of_a_window_off_the_region_grid_load_in_the_reference(ctx):
    """A window whose origin is no multiple of 8 columns (a slab window, a player far from the origin): the device's records
    carry RegionIndex of the chunks' WORLD positions, so the reference's own loader — LoadRegionFile's filing by word 0
    (WorldIO.cpp:48-62), then LoadGroupFromDisk (:167-221) — restores every column from them."""
    from gamecraft_b200 import mesher

    ox, oz = -3, 13
    w = util.HostWorld(3, 3, origin=(ox, oz)).generate("mixed", 8, radius=600, falloff=500)
    world = mesher.World(ctx, 3, 3).upload(w).set_origin(ox, oz)
    coords = all_chunks(w)
    loc, data = world.encode_rle(coords)
    for c, l in zip(coords.tolist(), loc):
        assert data[l["first_word"]] == region_index(c[0] + ox, c[1], c[2] + oz)
    ref = ob.RefWorld(util.HostWorld(3, 3))            # an empty reference window at the same world position
    ref.set_origin(ox, oz)
    for cz in range(3):
        for cx in range(3):
            ref.rle_clear_region(cx, cz)
    for c, l in zip(coords.tolist(), loc):
        rec = data[l["first_word"]:l["first_word"] + l["words"]]
        assert ref.rle_store_record(c[0], c[2], rec) == rec[0]
    for cz in range(3):
        for cx in range(3):
            assert ref.rle_load_group(cx, cz)
    got = util.HostWorld(3, 3)
    ref.download(got)
    assert np.array_equal(got.blocks, w.blocks)
    ref.close(); world.close()


def test_extreme_records(ctx):
    """one-voxel runs throughout (131 072 words, `items` stored as 0), a 65 535 + 1 split, a single wrapped run"""
    from gamecraft_b200 import mesher

    w = util.HostWorld(3, 3)
    w.blocks[w.slot(1, 0, 1)] = (np.arange(65536) & 1) + 1
    w.blocks[w.slot(1, 1, 1)] = 3; w.blocks[w.slot(1, 1, 1), 65535] = 9
    w.blocks[w.slot(1, 2, 1)] = 7
    w.blocks[w.slot(1, 3, 1), 40000:40003] = (5, 5, 6)
    coords = np.asarray([(1, cy, 1) for cy in range(4)], np.int32)
    loc, data, recs = host_records(w, coords)
    assert [len(r) for r in recs] == [2 + 2 * 65536, 6, 4, 10]
    world = mesher.World(ctx, 3, 3).upload_rle(loc, data)
    assert world.rle_refused() == 0
    got = util.HostWorld(3, 3)
    world.download(got, blocks=True, sun=False, light=False, surface=False)
    assert np.array_equal(got.blocks, w.blocks)
    loc2, data2 = world.encode_rle(coords)
    assert np.array_equal(data2, data)
    world.close()


def test_corrupt_records_are_refused(ctx):
    from gamecraft_b200 import mesher

    w = util.HostWorld(3, 3).generate("forest", 5)
    coords = np.asarray([(1, 0, 1), (1, 1, 1), (2, 0, 2)], np.int32)
    loc, data, recs = host_records(w, coords)
    bad = data.copy()
    bad[loc["first_word"][1] + 2] += 1                    # second record: one voxel too many
    world = mesher.World(ctx, 3, 3).upload_rle(loc, bad)
    assert world.rle_refused() == 1
    got = util.HostWorld(3, 3)
    world.download(got, blocks=True, sun=False, light=False, surface=False)
    assert np.array_equal(got.blocks[w.slot(1, 0, 1)], w.blocks[w.slot(1, 0, 1)])
    assert not got.blocks[w.slot(1, 1, 1)].any()         # refused: left as AIR
    assert np.array_equal(got.blocks[w.slot(2, 0, 2)], w.blocks[w.slot(2, 0, 2)])
    with pytest.raises(mesher.GcMeshError):
        loc3 = loc.copy(); loc3["first_word"][0] = 1      # odd start
        world.upload_rle(loc3, data)
    with pytest.raises(mesher.GcMeshError):
        loc4 = loc.copy(); loc4["words"][2] += 2           # runs past the stream
        world.upload_rle(loc4, data)
    world.close()


def test_records_to_meshes(ctx):
    """The load path on the device: region records -> block ids -> surface + light -> meshes, equal to the oracle's meshes
    of the host-prepared world."""
    from gamecraft_b200 import mesher

    w = util.HostWorld(6, 6, origin=(-3, -3)).build("islands", 4, radius=100, falloff=60)
    loc, data, _ = host_records(w, all_chunks(w))
    assert data.nbytes * 20 < w.blocks.nbytes            # the point of the format
    world = mesher.World(ctx, 6, 6).upload_rle(loc, data)
    world.compute_surface().light()
    coords = w.interior_chunks(1)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 4000)
    o = ob.OracleWorld(w)
    for c, m in zip(coords.tolist(), batch.submit(world, coords).wait().meshes()):
        util.assert_same_mesh(m, o.build_chunk(*c), str(c))
    batch.close(); world.close()

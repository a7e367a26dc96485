"""Region chunk format oracle (oracle/gc_rle_cpu.c) vs the reference's own SaveGroup / LoadGroupFromDisk (oracle/_ref)."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util

needs_ref = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")


def region_index(cx, cy, cz):
    return (cx & 7) + 8 * (cy + 4 * (cz & 7))          # RegionIndex, WorldIO.cpp:5-8


@needs_ref
@pytest.mark.parametrize("kind,seed", [("forest", 3), ("islands", 9), ("mixed", 12), ("noise", 2), ("checker", 1), ("empty", 0)])
def test_encode_equals_reference(kind, seed):
    kw = dict(radius=90, falloff=50) if kind in ("islands", "mixed") else {}
    w = util.HostWorld(3, 3, origin=(-1, -1) if kw else (0, 0)).generate(kind, seed, **kw)
    r = ob.RefWorld(w)
    for cz in range(3):
        for cx in range(3):
            recs = r.rle_save_group(cx, cz)
            for cy in range(4):
                mine = ob.rle_encode_chunk(w.blocks[w.slot(cx, cy, cz)], region_index(cx, cy, cz))
                assert np.array_equal(mine, recs[cy]), (kind, cx, cy, cz)
    r.close()


@needs_ref
def test_decode_equals_reference():
    w = util.HostWorld(3, 3, origin=(-1, -1)).generate("mixed", 5, radius=90, falloff=50)
    r = ob.RefWorld(w)
    recs = r.rle_save_group(1, 1)
    blank = np.zeros(65536, np.uint16)
    for cy in range(4):
        r.L.gcref_set_chunk(r.h, 1, cy, 1, blank.ctypes.data, None, None)
    assert r.rle_load_group(1, 1)
    got = util.HostWorld(3, 3)
    r.download(got)
    for cy in range(4):
        s = w.slot(1, cy, 1)
        assert np.array_equal(got.blocks[s], w.blocks[s])
        st, mine = ob.rle_decode_chunk(recs[cy])
        assert st == 0 and np.array_equal(mine, w.blocks[s])
    r.close()


@needs_ref
def test_offset_word_is_the_region_index_of_the_world_position():
    """SaveGroup files a chunk under RegionIndex(ChunkP & REGION_MASK) of its WORLD position (WorldIO.cpp:259-260): with a
    window origin that is no multiple of 8 the word differs from the window-local one, and LoadRegionFile's filing by that
    word (WorldIO.cpp:48-62) + LoadGroupFromDisk only finds records that carry the world-position word."""
    ox, oz = -3, 13
    w = util.HostWorld(3, 3, origin=(ox, oz)).generate("noise", 6)
    r = ob.RefWorld(w)
    r.set_origin(ox, oz)
    for cz in range(3):
        for cx in range(3):
            recs = r.rle_save_group(cx, cz)
            for cy in range(4):
                assert recs[cy][0] == region_index(cx + ox, cy, cz + oz) != region_index(cx, cy, cz)
                assert np.array_equal(ob.rle_encode_chunk(w.blocks[w.slot(cx, cy, cz)], region_index(cx + ox, cy, cz + oz)), recs[cy])
    # records with the window-local word land in other lists: the loader does not find the column
    recs = [ob.rle_encode_chunk(w.blocks[w.slot(1, cy, 1)], region_index(1, cy, 1)) for cy in range(4)]
    r.rle_clear_region(1, 1)
    for rec in recs:
        assert r.rle_store_record(1, 1, rec) == rec[0]
    assert not r.rle_load_group(1, 1)
    r.close()


def test_known_answers_and_round_trips():
    uniform = np.full(65536, 3, np.uint16)
    rec = ob.rle_encode_chunk(uniform, 77)
    assert rec.tolist() == [77, 2, 0, 3]                    # 65 536 wraps to 0 (WorldIO.cpp:200-205)
    st, out = ob.rle_decode_chunk(rec)
    assert st == 0 and np.array_equal(out, uniform)
    two = uniform.copy(); two[65535] = 9
    assert ob.rle_encode_chunk(two, 1).tolist() == [1, 4, 65535, 3, 1, 9]
    alt = (np.arange(65536) & 1).astype(np.uint16)         # every voxel its own run: items = 131 072 stored as uint16 0
    rec = ob.rle_encode_chunk(alt, 5)
    assert len(rec) == 2 + 2 * 65536 and rec[1] == 0 and rec[2] == 1
    st, out = ob.rle_decode_chunk(rec)
    assert st == 0 and np.array_equal(out, alt)
    rng = np.random.default_rng(4)
    for _ in range(5):
        runs = rng.integers(1, 4000, 200)
        blocks = np.repeat(rng.integers(0, 24, 200).astype(np.uint16), runs)[:65536]
        blocks = np.concatenate([blocks, np.zeros(65536 - len(blocks), np.uint16)])
        st, out = ob.rle_decode_chunk(ob.rle_encode_chunk(blocks, 0))
        assert st == 0 and np.array_equal(out, blocks)
    # corrupt records are refused, not overrun
    assert ob.rle_decode_chunk(np.array([0, 2, 100, 3], np.uint16))[0] == -1
    assert ob.rle_decode_chunk(np.array([0, 4, 65535, 3, 2, 9], np.uint16))[0] == -1

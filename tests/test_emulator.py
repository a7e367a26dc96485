"""The kernel's per-lane core (gamecraft_b200/csrc/mesh_core.cuh), compiled for the host and walked sequentially in the
kernel's phase order, against the oracle.  Bit-level logic is proven here before any GPU time is spent."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util

GOLDEN = util.load_golden()


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_emulator_matches_golden(key):
    entry = GOLDEN[key]
    w = util.golden_world(entry)
    for name, want in entry["chunks"].items():
        cx, cy, cz = map(int, name.split(","))
        assert util.mesh_signature(util.emu_build_chunk(w, cx, cy, cz)) == want, "%s chunk %s" % (key, name)


def test_emulator_empty_chunk():
    w = util.empty_world()
    m = util.emu_build_chunk(w, 1, 2, 1)
    assert m.vert_count == 0 and all(len(i) == 0 for i in m.indices)


def test_emulator_cap_edges():
    w = util.cap_edge_world(pairs=1, singles=1665)            # exactly 10 000 quads
    util.assert_same_mesh(util.emu_build_chunk(w, 1, 1, 1), ob.OracleWorld(w).build_chunk(1, 1, 1))
    w = util.cap_edge_world(pairs=2, singles=1664)            # 10 004 quads: over the cap
    m = util.emu_build_chunk(w, 1, 1, 1)
    assert m.overflow and m.quads_total == 10004 and m.vert_count == 0


def test_emulator_custom_block_table():
    infos, kz = util.custom_table()
    w = util.HostWorld(3, 3)
    rng = np.random.default_rng(5)
    w.blocks[:] = rng.integers(0, len(infos), w.blocks.shape, dtype=np.uint16) * (rng.random(w.blocks.shape) < 0.02)
    w.sun[:] = rng.integers(0, 16, w.sun.shape, dtype=np.uint8)
    w.light[:] = rng.integers(0, 16, w.light.shape, dtype=np.uint8)
    w.compute_surface(threads=1)
    o = ob.OracleWorld(w)
    o.set_block_table(infos, kz)
    for cy in range(4):
        want = o.build_chunk(1, cy, 1)
        got = util.emu_build_chunk(w, 1, cy, 1, table=infos, kill_zone=kz)
        if want.overflow:
            assert got.overflow
        else:
            util.assert_same_mesh(got, want, "cy=%d" % cy)


@pytest.mark.parametrize("seed,fill", [(1, 0.03), (2, 0.01), (3, 0.05)])
def test_emulator_arbitrary_inputs(seed, fill):
    """Random ids, random light nibbles, a surface map unrelated to the blocks: every chunk of the centre column."""
    w = util.arbitrary_world(seed, fill)
    o = ob.OracleWorld(w)
    for cy in range(4):
        want = o.build_chunk(1, cy, 1)
        got = util.emu_build_chunk(w, 1, cy, 1)
        if want.overflow:
            assert got.overflow
        else:
            util.assert_same_mesh(got, want, "seed %d cy %d" % (seed, cy))

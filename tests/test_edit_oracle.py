"""Block-edit oracle (oracle/gc_edit_cpu.c + the fill of gc_light_cpu.c on the edited blocks) vs the reference's own
SetBlock -> ChunkOverflowed -> RecomputeLight -> FlagChunkForUpdate compiled verbatim (oracle/_ref)."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util

needs_ref = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")

LANTERN, STONE, WATER = 14, 3, 8
APPLIED, UNCHANGED, IGNORED, NOT_BUILT, OVERFLOW = range(5)


def start(kind, seed, size=5, **kw):
    """blocks -> (host world lit by the reference's own SetLightNodes, the reference world holding the same state)."""
    w = util.HostWorld(size, size).generate(kind, seed, **kw)
    w.blocks[(w.blocks == 17) | (w.blocks == 21)] = LANTERN    # one emission level: the reference's seeding is a max then
    z = util.HostWorld(size, size)
    z.blocks[:] = w.blocks
    r = ob.RefWorld(z)
    r.compute_surface(); r.compute_light()
    for cz in range(size):
        for cx in range(size):
            for cy in range(4):
                r.set_chunk_built(cx, cy, cz, True, 0)
    r.download(z)
    r.pending_updates(clear=True)
    return z, r


def random_edit(rng, z, size):
    wx, wz = int(rng.integers(32, 32 * (size - 1))), int(rng.integers(32, 32 * (size - 1)))
    s = int(z.surface[(wz >> 5) * size + (wx >> 5)][(wz & 31) * 32 + (wx & 31)])
    mode = int(rng.integers(0, 5))
    if mode == 0: return wx, max(s - 1, 0), wz, 0                              # dig the top block
    if mode == 1: return wx, min(s, 254), wz, STONE                            # build on top
    if mode == 2: return wx, int(rng.integers(0, max(s, 1))), wz, 0            # dig somewhere below
    if mode == 3: return wx, min(s + int(rng.integers(0, 6)), 254), wz, LANTERN  # a lantern in the air
    return wx, int(rng.integers(0, max(s, 1))), wz, LANTERN                    # a lantern in the ground


def effective_sun(w):
    """Sunlight as GetSunlight reads it (Lighting.cpp:69-79)."""
    out = np.where(w.sun == 0, 1, w.sun).astype(np.uint8)
    for col in range(w.size_x * w.size_z):
        surf = w.surface[col].reshape(32, 1, 32)                                # [z][.][x]
        for cy in range(4):
            ys = (np.arange(64) + 64 * cy).reshape(1, 64, 1)
            e = out[col * 4 + cy].reshape(32, 64, 32)
            e[np.broadcast_to(ys >= surf, e.shape)] = 15
    return out


@needs_ref
@pytest.mark.parametrize("kind,seed", [("forest", 1), ("noise", 2), ("checker", 3)])
def test_decisions_block_surface_and_flags_match_reference(kind, seed):
    size = 5
    z, r = start(kind, seed, size)
    mine = util.HostWorld(size, size)
    mine.blocks[:] = z.blocks; mine.surface[:] = z.surface
    verts = np.zeros(size * size * 4, np.int32)
    rng = np.random.default_rng(seed)
    seen = set()
    for e in range(40):
        wx, wy, wz, b = random_edit(rng, z, size)
        if e % 9 == 4:
            wy = 300 if e % 2 else -3                                           # outside the world: ignored
        slot = ((wz >> 5) * size + (wx >> 5)) * 4 + (min(max(wy, 0), 255) >> 6)
        cx, cy, cz = wx >> 5, slot & 3, wz >> 5
        if e % 7 == 3:
            verts[slot] = -1; r.set_chunk_built(cx, cy, cz, False, 0)           # never built: refused
        elif e % 7 == 5:
            verts[slot] = 39990; r.set_chunk_built(cx, cy, cz, True, 39990)     # 12 or more new vertices overflow
        before = z.blocks.copy()
        flagged = np.zeros(size * size * 4, np.uint8)
        status = ob.set_block(mine, verts, wx, wy, wz, b, flagged)
        changed = r.set_block(wx, wy, wz, b)
        r.download(z)
        pending = r.pending_updates(clear=True).reshape(-1)
        assert np.array_equal(mine.blocks, z.blocks), "edit %d %r status %d" % (e, (wx, wy, wz, b), status)
        assert changed == bool((before != z.blocks).any())
        old = int(before[slot][(wx & 31) + 32 * ((wy & 63) + 64 * (wz & 31))]) if 0 <= wy < 256 else -1
        assert changed == (status == APPLIED or (status == OVERFLOW and old != 0))   # an overflowing block leaves AIR behind
        seen.add(status)
        if status == APPLIED:
            col = (wz >> 5) * size + (wx >> 5)
            i = (wz & 31) * 32 + (wx & 31)
            if mine.surface[col][i] != 0:                                        # empty column: the reference keeps a stale entry
                assert mine.surface[col][i] == z.surface[col][i]
            mine.surface[col][i] = z.surface[col][i]
            # FlagChunkForUpdate's 27 chunks are all pending in the reference (which adds the chunks its light queues touched)
            assert (pending[flagged != 0] == 1).all()
            assert flagged.sum() >= 1
        else:
            assert not flagged.any() and not pending.any()
        if verts[slot] != 0:
            verts[slot] = 0; r.set_chunk_built(cx, cy, cz, True, 0)
    r.close()
    assert seen >= {APPLIED, IGNORED, NOT_BUILT, OVERFLOW}


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_reference_relight_reaches_the_fill_on_terrain_edits(seed):
    """On forest terrain the reference's incremental RecomputeLight lands exactly on the from-scratch fill of the edited
    blocks (the oracle of the device relight), edit after edit."""
    size = 5
    z, r = start("forest", seed, size)
    rng = np.random.default_rng(seed)
    for e in range(30):
        r.set_block(*random_edit(rng, z, size))
        r.download(z)
    c = util.HostWorld(size, size)
    c.blocks[:] = z.blocks
    c.surface[:] = ob.compute_surface(c)
    c.sun[:], c.light[:] = ob.light_fill(c, c.surface)
    assert np.array_equal(c.surface, z.surface)
    assert np.array_equal(effective_sun(c), effective_sun(z))
    assert np.array_equal(c.light, z.light)
    r.close()


@needs_ref
def test_reference_relight_can_stop_short_of_the_fill():
    """Documented deviation: after edits under water or a surface rise over caves the reference's removal
    queues can leave sunlight a step below its own from-scratch fill — never above it, and block light is unaffected."""
    size = 5
    z, r = start("islands", 8, size, radius=90, falloff=50)
    rng = np.random.default_rng(8)
    for e in range(40):
        r.set_block(*random_edit(rng, z, size))
        r.download(z)
    c = util.HostWorld(size, size)
    c.blocks[:] = z.blocks
    c.surface[:] = ob.compute_surface(c)
    c.sun[:], c.light[:] = ob.light_fill(c, c.surface)
    assert np.array_equal(c.surface, z.surface)
    ec, ez = effective_sun(c), effective_sun(z)
    assert not (ez > ec).any()
    assert int((ez < ec).sum()) < 100
    assert np.array_equal(c.light, z.light)
    r.close()

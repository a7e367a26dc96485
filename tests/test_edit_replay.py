"""The relight replay (gamecraft_b200/csrc/relight_core.cuh: the code the device runs under GC_WORLD_EDIT_REPLAY, built for the
host in tests/emu) against the reference's own SetBlock -> RecomputeLight -> FlagChunkForUpdate compiled verbatim (oracle/_ref):
block ids, surface map, both raw light arrays and the pendingUpdate set after every edit — also on the worlds where the
reference's removal queues stop short of a from-scratch fill (islands) and where weak and strong emitters mix."""
import ctypes
import os

import numpy as np
import pytest

import gcutil as util
from oracle import binding as ob
from test_edit_oracle import APPLIED, random_edit, start

needs_ref = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")


def replay_edit(emu, h, verts, dirty, step, emitted, wx, wy, wz, block):
    """SetBlock with the replayed relight on host arrays: the decisions, the store and FlagChunkForUpdate's 27 chunks come from the
    edit oracle (pinned on its own); the surface entry is left to the replay, which compares the old one with the new one."""
    flagged = np.zeros_like(dirty)
    col, i = (wz >> 5) * h.size_x + (wx >> 5), (wz & 31) * 32 + (wx & 31)
    keep = int(h.surface[col][i]) if 0 <= wy < 256 else 0
    st = ob.set_block(h, verts, wx, wy, wz, block, flagged)
    if st == APPLIED:
        h.surface[col][i] = keep
        dirty |= flagged
        vp = ctypes.c_void_p
        emu.gcemu_recompute_light(h.size_x, h.size_z, vp(h.blocks.ctypes.data), vp(h.sun.ctypes.data), vp(h.light.ctypes.data),
                                  vp(h.surface.ctypes.data), vp(step.ctypes.data), vp(emitted.ctypes.data), vp(verts.ctypes.data),
                                  vp(dirty.ctypes.data), wx, wy, wz)
    return st


@needs_ref
@pytest.mark.parametrize("kind,seed,kw,n", [("forest", 1, {}, 40), ("islands", 8, dict(radius=90, falloff=50), 60),
                                            ("mixed", 3, dict(radius=90, falloff=50), 60), ("noise", 2, {}, 40)])
def test_replay_equals_the_reference_s_setblock(kind, seed, kw, n):
    emu = ctypes.CDLL(os.path.join(util.HERE, "emu", "libgcemu.so"))
    step, emitted = ob.default_light_rules()
    size = 5
    z, r = start(kind, seed, size, **kw)
    mine = util.HostWorld(size, size)
    mine.blocks[:] = z.blocks; mine.sun[:] = z.sun; mine.light[:] = z.light; mine.surface[:] = z.surface
    verts = np.zeros(size * size * 4, np.int32)
    rng = np.random.default_rng(seed)
    applied = 0
    for e in range(n):
        edit = random_edit(rng, z, size)
        if e % 10 == 7:                                                        # an emitter taken away again: removal + re-scatter of block light
            edit = edit[:3] + (0,)
        dirty = np.zeros(size * size * 4, np.uint8)
        st = replay_edit(emu, mine, verts, dirty, step, emitted, *edit)
        changed = r.set_block(*edit)
        r.download(z)
        pending = r.pending_updates(clear=True).reshape(-1)
        assert changed == (st == APPLIED)
        applied += int(changed)
        where = "%s edit %d %r" % (kind, e, edit)
        assert np.array_equal(mine.blocks, z.blocks), where
        assert np.array_equal(mine.surface, z.surface), where
        assert np.array_equal(mine.sun, z.sun), "%s: %d sunlight bytes differ" % (where, int((mine.sun != z.sun).sum()))
        assert np.array_equal(mine.light, z.light), "%s: %d blockLight bytes differ" % (where, int((mine.light != z.light).sum()))
        assert np.array_equal(dirty != 0, pending != 0), "%s: pendingUpdate sets differ" % where
    assert applied >= n // 2
    r.close()


@needs_ref
def test_replay_keeps_the_stale_surface_entry_of_an_emptied_column():
    """ComputeSurfaceAt never writes when the column holds no block at all (Lighting.cpp:9-16): the entry keeps its old value —
    in the reference, and in the replay."""
    emu = ctypes.CDLL(os.path.join(util.HERE, "emu", "libgcemu.so"))
    step, emitted = ob.default_light_rules()
    size = 3
    w = util.HostWorld(size, size)
    w.set_block(40, 70, 40, 3)                                                 # one stone block in an otherwise empty window
    r = ob.RefWorld(w)
    r.compute_surface(); r.compute_light()
    for cz in range(size):
        for cx in range(size):
            for cy in range(4):
                r.set_chunk_built(cx, cy, cz, True, 0)
    z = util.HostWorld(size, size)
    r.download(z)
    r.pending_updates(clear=True)
    mine = util.HostWorld(size, size)
    mine.blocks[:] = z.blocks; mine.sun[:] = z.sun; mine.light[:] = z.light; mine.surface[:] = z.surface
    verts = np.zeros(size * size * 4, np.int32)
    dirty = np.zeros(size * size * 4, np.uint8)
    assert replay_edit(emu, mine, verts, dirty, step, emitted, 40, 70, 40, 0) == APPLIED
    assert r.set_block(40, 70, 40, 0)
    r.download(z)
    col, i = 1 * size + 1, (40 & 31) * 32 + (40 & 31)
    assert z.surface[col][i] == 71 == mine.surface[col][i]                     # stale: the column is empty now
    assert np.array_equal(mine.sun, z.sun) and np.array_equal(mine.light, z.light) and np.array_equal(mine.surface, z.surface)
    r.close()

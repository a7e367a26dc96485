"""Column pre-processing oracle (oracle/gc_light_cpu.c: surface + light fill as a fixpoint) vs the reference's own
ComputeSurface / SetLightNodes compiled verbatim (oracle/_ref), on the same blocks."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util

needs_ref = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")

LANTERN, MAGMA, LAVA = 14, 17, 21


def reference_fill(blocks_world):
    z = util.HostWorld(blocks_world.size_x, blocks_world.size_z)
    z.blocks[:] = blocks_world.blocks
    r = ob.RefWorld(z)
    r.compute_surface(); r.compute_light(); r.download(z)
    r.close()
    return z


@needs_ref
def test_light_rules_match_reference():
    r = ob.RefWorld(size=3)
    steps, emitted = r.light_steps(), r.light_emitted()
    r.close()
    step, emit = ob.default_light_rules()
    for b in range(24):
        want = 255 if steps[b] == 2**31 - 1 else steps[b]
        assert step[b] == want and emit[b] == emitted[b], "block %d" % b
    assert (step[24:] == 255).all() and (emit[24:] == 0).all()


@needs_ref
@pytest.mark.parametrize("kind,seed", [("forest", 5), ("islands", 8), ("noise", 4), ("checker", 6), ("flat", 1)])
def test_fixpoint_equals_reference(kind, seed):
    """Worlds without light emitters (or none in reach of a column border): both arrays must be identical."""
    kw = dict(radius=90, falloff=50) if kind == "islands" else {}
    w = util.HostWorld(5, 5, origin=(-2, -2) if kw else (0, 0)).generate(kind, seed, **kw)
    # keep only the strongest emitter kind: with one emission level the reference's assignment (Lighting.cpp:240) is a max
    w.blocks[(w.blocks == MAGMA) | (w.blocks == LAVA)] = LANTERN
    ref = reference_fill(w)
    surface = ob.compute_surface(w)
    assert np.array_equal(surface, ref.surface)
    sun, light = ob.light_fill(w, surface)
    assert np.array_equal(sun, ref.sun)
    assert np.array_equal(light, ref.light)


@needs_ref
def test_sunlight_is_exact_with_mixed_emitters():
    """With weak and strong emitters the reference's block light depends on the column order; sunlight
    never does, and block light may only differ where the canonical fixpoint is brighter."""
    w = util.HostWorld(5, 5, origin=(-2, -2)).generate("mixed", 3, radius=90, falloff=50)
    ref = reference_fill(w)
    surface = ob.compute_surface(w)
    sun, light = ob.light_fill(w, surface)
    assert np.array_equal(surface, ref.surface)
    assert np.array_equal(sun, ref.sun)
    assert (light >= ref.light).all()
    assert light.max() == 15 and sun.max() >= 2


def test_lantern_in_a_cave():
    """Known answer: a lantern in a sealed 3x3x3 air pocket inside stone; light falls by one per step, sunlight stays out."""
    w = util.HostWorld(3, 3)
    w.blocks[:] = 0
    for cz in range(3):
        for cx in range(3):
            w.blocks[w.slot(cx, 0, cz), :] = 3                       # y 0..63 stone
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                w.set_block(48 + dx, 20 + dy, 48 + dz, 0)
    w.set_block(48, 20, 48, LANTERN)
    surface = ob.compute_surface(w)
    assert (surface == 64).all()
    sun, light = ob.light_fill(w, surface)
    s = w.slot(1, 0, 1)
    at = lambda a, x, y, z: int(a[s, (x - 32) + 32 * (y + 64 * (z - 32))])
    assert at(light, 48, 20, 48) == 15 and at(light, 49, 20, 48) == 14 and at(light, 49, 21, 48) == 13 and at(light, 49, 21, 49) == 12
    assert at(light, 50, 20, 48) == 0                                   # stone is opaque: nothing stored
    assert int(sun.max()) == 0                                          # sealed: every stored sunlight value stays 0
    assert int((light > 0).sum()) == 27


def _straight_down_start(w, surface):
    """numpy model of what the device scan (gc_light_scan_kernel) stores before the relaxation: each candidate starts at the
    light that reaches it straight from above — every cell hands its value less its own lightStep to the cell below it, a sky
    cell of a pre-processed column reads 15, an opaque cell passes nothing."""
    step, _ = ob.default_light_rules()
    step = step.astype(np.int32)
    nc = w.size_x * w.size_z
    ids = (w.blocks.reshape(nc, 4, 32, 64, 32) & 255).transpose(0, 1, 3, 2, 4).reshape(nc, 256, 32, 32)
    top = surface.reshape(nc, 32, 32).astype(np.int32)
    cx, cz = np.arange(nc) % w.size_x, np.arange(nc) // w.size_x
    inner = ((cx >= 1) & (cx < w.size_x - 1) & (cz >= 1) & (cz < w.size_z - 1))[:, None, None]
    offer = np.zeros((nc, 32, 32), np.int32)
    out = np.zeros((nc, 256, 32, 32), np.uint8)
    for y in range(255, -1, -1):
        st = step[ids[:, y]]
        opaque = st == 255
        val = np.where(offer > 1, offer, 0)
        out[:, y] = np.where((y < top) & ~opaque, val, 0)
        offer = np.where(opaque, 0, np.where(y >= top, np.where(inner, 15 - st, 0), val - st))
    return out.reshape(nc, 4, 64, 32, 32).transpose(0, 1, 3, 2, 4).reshape(nc * 4, -1)


@pytest.mark.parametrize("kind,seed", [("forest", 5), ("islands", 8), ("noise", 4), ("checker", 6), ("mixed", 3)])
def test_straight_down_start_is_a_lower_bound_of_the_fill(kind, seed):
    """The device fill starts its sunlight relaxation from the straight-down values instead of zero.  A max-relaxation from any
    state at or below the fixpoint ends at the fixpoint, so all this needs is: never above the fill's value, anywhere."""
    kw = dict(radius=90, falloff=50) if kind in ("islands", "mixed") else {}
    w = util.HostWorld(5, 5, origin=(-2, -2) if kw else (0, 0)).generate(kind, seed, **kw)
    surface = ob.compute_surface(w)
    sun, _ = ob.light_fill(w, surface)
    start = _straight_down_start(w, surface)
    assert (start <= sun).all()
    if kind == "islands":   # open water: most of the fill is the straight-down value
        lit = sun > 0
        assert ((start == sun) & lit).sum() > 0.8 * lit.sum()

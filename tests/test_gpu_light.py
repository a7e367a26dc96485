"""Device surface map + light flood fill (gcmesh_world_compute_surface / gcmesh_world_light, through the C ABI) against
the fixpoint oracle (oracle/gc_light_cpu.c) and, for sunlight, against the input producer's queue-based fill."""
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


def device_fill(ctx, host):
    """blocks only -> (surface, sun, light) computed on the device."""
    from gamecraft_b200 import mesher

    blank = util.HostWorld(host.size_x, host.size_z)
    blank.blocks[:] = host.blocks
    blank.sun[:] = 7; blank.light[:] = 9; blank.surface[:] = 200        # stale values the fill has to overwrite
    world = mesher.World(ctx, host.size_x, host.size_z).upload(blank)
    world.compute_surface().light()
    stats = world.light_stats()
    out = util.HostWorld(host.size_x, host.size_z)
    world.download(out)
    world.close()
    return out, stats


@pytest.mark.parametrize("kind,size,seed", [("forest", 6, 5), ("islands", 6, 8), ("mixed", 5, 3), ("noise", 3, 4), ("checker", 3, 6), ("flat", 3, 1), ("empty", 3, 1)])
def test_device_fill_equals_oracle(ctx, kind, size, seed):
    kw = dict(radius=90, falloff=50) if kind in ("islands", "mixed") else {}
    w = util.HostWorld(size, size, origin=(-(size // 2), -(size // 2)) if kw else (0, 0)).generate(kind, seed, **kw)
    got, stats = device_fill(ctx, w)
    surface = ob.compute_surface(w)
    sun, light = ob.light_fill(w, surface)
    assert np.array_equal(got.surface, surface)
    assert np.array_equal(got.sun, sun), "sunlight: %d cells differ" % int((got.sun != sun).sum())
    assert np.array_equal(got.light, light), "blockLight: %d cells differ" % int((got.light != light).sum())
    if kind in ("forest", "islands", "flat", "empty"):
        assert stats["seeds"] == 0 and not light.any()
    if kind in ("mixed", "noise"):
        assert stats["seeds"] > 0 and stats["brick_visits"] > 0 and light.max() == 15
    assert stats["launches"] == 2 and stats["sweeps"] <= 14                    # one scan + one persistent relaxation, whatever the world
    if kind == "empty":
        assert stats["sweeps"] == 0 and stats["brick_visits"] == 0


def test_device_sunlight_equals_queue_fill(ctx):
    """The producer's queue-based restatement (gcworld.c, tested equal to the reference's SetLightNodes) on a forest world."""
    w = util.HostWorld(8, 8).build("forest", 2)
    got, _ = device_fill(ctx, w)
    assert np.array_equal(got.surface, w.surface)
    assert np.array_equal(got.sun, w.sun)
    assert np.array_equal(got.light, w.light)


def test_meshes_from_device_filled_light(ctx):
    """blocks -> device surface + light -> meshes: identical to meshing the host-filled world."""
    from gamecraft_b200 import mesher

    w = util.HostWorld(6, 6, origin=(-3, -3)).build("islands", 11, radius=100, falloff=60)
    coords = w.interior_chunks(1)
    blank = util.HostWorld(6, 6)
    blank.blocks[:] = w.blocks
    world = mesher.World(ctx, 6, 6).upload(blank)
    world.compute_surface().light()
    batch = mesher.Batch(ctx, len(coords), len(coords) * 4000)
    batch.submit(world, coords).wait()
    meshes = batch.meshes()
    o = ob.OracleWorld(w)
    for (cx, cy, cz), m in zip(coords, meshes):
        util.assert_same_mesh(m, o.build_chunk(cx, cy, cz), "%s" % ((cx, cy, cz),))
    batch.close(); world.close()


def test_custom_light_rules(ctx):
    """A table where stone passes light (step 3) and grass glows (emits 6)."""
    from gamecraft_b200 import mesher

    step, emitted = mesher.default_light_rules()
    ostep, oemit = ob.default_light_rules()
    assert np.array_equal(step, ostep) and np.array_equal(emitted, oemit)
    step[3] = 3; emitted[1] = 6; step[1] = 1
    w = util.HostWorld(4, 4).generate("forest", 9)
    c2 = mesher.Context(0)
    c2.set_light_rules(step, emitted)
    got, stats = device_fill(c2, w)
    c2.close()
    surface = ob.compute_surface(w)
    sun, light = ob.light_fill(w, surface, step, emitted)
    assert stats["seeds"] > 0
    assert np.array_equal(got.sun, sun) and np.array_equal(got.light, light)


def test_non_square_window(ctx):
    """size_x != size_z: column strides of the mesher and of the light fill (the reference's window is always square)."""
    from gamecraft_b200 import mesher

    w = util.HostWorld(7, 4, origin=(-3, -2)).generate("mixed", 23, radius=80, falloff=40)
    surface = ob.compute_surface(w)
    sun, light = ob.light_fill(w, surface)
    got, _ = device_fill(ctx, w)
    assert np.array_equal(got.surface, surface) and np.array_equal(got.sun, sun) and np.array_equal(got.light, light)
    w.surface[:] = surface; w.sun[:] = sun; w.light[:] = light
    coords = w.interior_chunks(1)
    assert len(coords) == 5 * 2 * 4
    world = mesher.World(ctx, 7, 4).upload(w)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 6000)
    batch.submit(world, coords).wait()
    o = ob.OracleWorld(w)
    for (cx, cy, cz), m in zip(coords, batch.meshes()):
        util.assert_same_mesh(m, o.build_chunk(cx, cy, cz), "%s" % ((cx, cy, cz),))
    batch.close(); world.close()


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")
@pytest.mark.parametrize("kind,seed,kw", [("forest", 4, {}), ("islands", 9, dict(radius=90, falloff=50)), ("mixed", 3, dict(radius=90, falloff=50))])
def test_device_fill_against_the_reference_s_own_fill(ctx, kind, seed, kw):
    """The documented f1 divergence, bounded on the GPU path itself: the device fill next to ComputeSurface + SetLightNodes of the
    reference compiled verbatim (columns in z-major order).  Surface and sunlight are identical on every world; block light is
    identical where all emitters shine alike (forest, islands: none) and, where weak and strong emitters mix, never below the
    reference's — the reference re-assigns a seeded emitter's own value below light it had already received (Lighting.cpp:240),
    the fixpoint keeps the brighter value."""
    size = 5
    w = util.HostWorld(size, size, origin=(-(size // 2), -(size // 2)) if kw else (0, 0)).generate(kind, seed, **kw)
    got, _ = device_fill(ctx, w)
    z = util.HostWorld(size, size)
    z.blocks[:] = w.blocks
    r = ob.RefWorld(z)
    r.compute_surface(); r.compute_light()
    r.download(z)
    r.close()
    assert np.array_equal(got.surface, z.surface)
    assert np.array_equal(got.sun, z.sun), "sunlight: %d cells differ from the reference" % int((got.sun != z.sun).sum())
    if kind == "mixed":
        assert not (got.light < z.light).any()
        assert int((got.light != z.light).sum()) < got.light.size // 1000
    else:
        assert np.array_equal(got.light, z.light)

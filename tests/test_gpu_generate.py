"""Terrain generated on the device (gcmesh_world_generate, through the C ABI) against the host input producer
(gamecraft_b200/worldgen/gcworld.c: the reference's layering / tree / island fall-off rules over this project's own seeded
noise — the reference's FastNoiseSIMD is not vendored, so parity with the reference's *noise* is unpinned by construction):
block ids identical, and the device-only pipeline generate -> surface -> light -> mesh equal to the oracle's meshes."""
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


CASES = [
    ("forest", 6, 1337, (-3, -3), {}),
    ("forest", 5, 7, (40, -17), {}),
    ("islands", 7, 4242, (-3, -4), dict(radius=90, falloff=50)),          # the island's edge and fall-off ring cross the window
    ("islands", 5, 8, (9, 9), dict(radius=512, falloff=448)),
    ("flat", 4, 1, (0, 0), {}),
    ("flat", 5, 1, (-2, -3), dict(radius=70, falloff=30)),
    ("snow", 6, 11, (-3, -3), {}),                                          # ice over the blend band, water inside taller ground
    ("snow", 5, 3, (-2, -3), dict(radius=80, falloff=30)),
    ("desert", 6, 5, (2, -5), {}),                                          # cacti, also on the sea floor
    ("desert", 5, 6, (-3, -2), dict(radius=75, falloff=20)),
    ("volcanic", 6, 9, (-1, 7), {}),
    ("volcanic", 5, 10, (-2, -2), dict(radius=70, falloff=25)),
    ("dungeon", 4, 1, (0, 0), {}),
]


@pytest.mark.parametrize("kind,size,seed,origin,kw", CASES)
def test_generated_blocks_equal_host_producer(ctx, kind, size, seed, origin, kw):
    from gamecraft_b200 import mesher

    host = util.HostWorld(size, size, origin=origin).generate(kind, seed, **kw)
    junk = util.HostWorld(size, size)
    junk.blocks[:] = 5; junk.sun[:] = 3; junk.light[:] = 9; junk.surface[:] = 77      # the call has to clear all of it
    world = mesher.World(ctx, size, size).upload(junk)
    world.generate(kind, seed, origin=origin, **kw)
    got = util.HostWorld(size, size)
    world.download(got, blocks=True)
    world.close()
    diff = int((got.blocks != host.blocks).sum())
    assert diff == 0, "%d of %d block ids differ" % (diff, host.blocks.size)
    assert not got.sun.any() and not got.light.any()
    assert np.array_equal(got.surface, ob.compute_surface(host)), "the generator's surface map is not ComputeSurface's"
    if kind in ("forest", "islands"):
        assert len(np.unique(host.blocks)) >= 5                                    # air, grass, dirt, stone, water and / or trees


needs_ref = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")


@needs_ref
@pytest.mark.parametrize("kind,radius,falloff", [("flat", 2**31 - 1, 2**31 - 1 - 64), ("flat", 100, 36), ("flat", 33, 1),
                                                 ("grid", 2**31 - 1, 0), ("grid", 90, 0), ("void", 50, 0), ("dungeon", 50, 0)])
def test_noise_free_generators_equal_the_reference_itself(ctx, kind, radius, falloff):
    """GenerateFlatTerrain / GenerateGridTerrain / GenerateVoidTerrain / GenerateDungeon run from the reference's own sources (oracle/_ref)."""
    from gamecraft_b200 import mesher
    from test_generate_oracle import reference_blocks

    size = 5
    want = reference_blocks(kind, size, radius, falloff)
    world = mesher.World(ctx, size, size)
    world.generate(kind, 1, origin=(0, 0), radius=radius, falloff=falloff)
    got = util.HostWorld(size, size)
    world.download(got, blocks=True)
    world.close()
    diff = int((got.blocks != want).sum())
    assert diff == 0, "%d block ids differ from the reference's own %s generator" % (diff, kind)
    assert np.array_equal(got.surface, ob.compute_surface(got))


def test_void_plate_follows_the_origin(ctx):
    from gamecraft_b200 import mesher

    world = mesher.World(ctx, 4, 4)
    got = util.HostWorld(4, 4)
    world.generate("void", 1, origin=(-2, -1))                                    # the world origin is window column (2, 1)
    world.download(got, blocks=True)
    assert int((got.blocks != 0).sum()) == 1024 and (got.blocks[(1 * 4 + 2) * 4].reshape(32, 64, 32)[:, 40, :] == 1).all()
    world.generate("void", 1, origin=(5, 5))                                      # origin outside the window: nothing
    world.download(got, blocks=True)
    assert not got.blocks.any()
    world.close()


def test_device_only_pipeline_meshes_like_the_oracle(ctx):
    from gamecraft_b200 import mesher

    size, seed, origin = 5, 21, (-2, 3)
    world = mesher.World(ctx, size, size)
    world.generate("forest", seed, origin=origin).light()                          # the generator leaves the surface map behind
    host = util.HostWorld(size, size, origin=origin).generate("forest", seed)
    host.surface[:] = ob.compute_surface(host)
    host.sun[:], host.light[:] = ob.light_fill(host, host.surface)
    coords = host.interior_chunks()
    orc = ob.OracleWorld(host)
    quads = 0
    for c, m in zip(coords.tolist(), mesher.rebuild_chunks(world, coords)):
        util.assert_same_mesh(m, orc.build_chunk(*c), "chunk %r" % (c,))
        quads += m.vert_count // 4
    assert quads > 10000
    # freshly generated chunks have never been built: SetBlock refuses them (World.cpp:562-565) until they are meshed again
    assert (world.vertex_counts()[1:-1, 1:-1] >= 0).all()
    world.generate("forest", seed, origin=origin)
    assert (world.vertex_counts() == -1).all()
    status, rebuild = world.set_blocks([(40, 5, 40, 0)])
    assert status.tolist() == [mesher.EDIT_NOT_BUILT] and len(rebuild) == 0
    world.close()


@pytest.mark.parametrize("kind,seed", [("snow", 4), ("desert", 12), ("volcanic", 2), ("dungeon", 1)])
def test_device_only_pipeline_on_the_other_biomes(ctx, kind, seed):
    """generate -> light -> mesh entirely on the device for the biomes with other block kinds (ice and snow, cactus, lava's
    light, the dungeon's enclosed rooms), against the oracle's fill and meshes of the host producer's blocks."""
    from gamecraft_b200 import mesher

    size, origin = 5, (-2, -2)
    world = mesher.World(ctx, size, size)
    world.generate(kind, seed, origin=origin).light()
    host = util.HostWorld(size, size, origin=origin).generate(kind, seed)
    host.surface[:] = ob.compute_surface(host)
    host.sun[:], host.light[:] = ob.light_fill(host, host.surface)
    got = util.HostWorld(size, size); world.download(got, blocks=True)
    assert np.array_equal(got.blocks, host.blocks) and np.array_equal(got.surface, host.surface)
    assert np.array_equal(got.sun, host.sun) and np.array_equal(got.light, host.light)
    coords = host.interior_chunks()
    orc = ob.OracleWorld(host)
    quads = 0
    for c, m in zip(coords.tolist(), mesher.rebuild_chunks(world, coords)):
        util.assert_same_mesh(m, orc.build_chunk(*c), "chunk %r" % (c,))
        quads += m.vert_count // 4
    assert quads > 5000
    world.close()


def test_light_after_generate_skips_the_clear_only_while_the_arenas_are_known_zero(ctx):
    """gcmesh_world_generate leaves both light arenas zero, so the fill that follows does not clear them again; any writer in
    between (a second fill, an upload) brings the clear back."""
    from gamecraft_b200 import mesher

    size, seed = 4, 3
    world = mesher.World(ctx, size, size)
    world.generate("islands", seed, origin=(-2, -2), radius=60, falloff=20).compute_surface().light()
    a = util.HostWorld(size, size); world.download(a, blocks=True)
    want_sun, want_light = ob.light_fill(a, ob.compute_surface(a))
    assert np.array_equal(a.sun, want_sun) and np.array_equal(a.light, want_light)
    world.light()                                                                  # arenas hold the first fill now
    b = util.HostWorld(size, size); world.download(b)
    assert np.array_equal(b.sun, want_sun) and np.array_equal(b.light, want_light)
    world.generate("islands", seed, origin=(-2, -2), radius=60, falloff=20)
    junk = np.full(65536, 13, np.uint8)
    mesher._check(mesher.lib().gcmesh_world_upload_chunk(world.h, 1, 0, 1, None, junk.ctypes.data, junk.ctypes.data), "upload_chunk")
    world.compute_surface().light()
    c = util.HostWorld(size, size); world.download(c)
    assert np.array_equal(c.sun, want_sun) and np.array_equal(c.light, want_light)
    world.close()


def test_bad_arguments(ctx):
    from gamecraft_b200 import mesher

    world = mesher.World(ctx, 3, 3)
    with pytest.raises(mesher.GcMeshError):
        world.generate(9, 1)
    with pytest.raises(mesher.GcMeshError):
        world.generate("flat", 1, origin=(2000, 0))                                # Square() of the reference would overflow
    with pytest.raises(mesher.GcMeshError):
        world.generate("islands", 1, radius=50, falloff=50)
    world.close()

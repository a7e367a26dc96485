"""Block edits on a device-resident window (gcmesh_world_set_blocks, through the C ABI) against the edit oracle
(oracle/gc_edit_cpu.c for the decisions, the from-scratch fill of oracle/gc_light_cpu.c for the light) and against the
mesher oracle for the chunks the call flags — and for the ones it does not."""
import numpy as np
import pytest

import gcutil as util
from oracle import binding as ob
from test_edit_oracle import random_edit, effective_sun, APPLIED, UNCHANGED, IGNORED, NOT_BUILT, OVERFLOW

pytestmark = pytest.mark.gpu

LANTERN, STONE, WATER, GLASS, MAGMA = 14, 3, 8, 22, 17


@pytest.fixture(scope="module")
def ctx():
    from gamecraft_b200 import mesher

    c = mesher.Context(0)
    yield c
    c.close()


def to_window(w, arr):
    """brick-major (slot, voxel) -> (Z, Y, X) over the whole window."""
    a = arr.reshape(w.size_z, w.size_x, 4, 32, 64, 32)                       # cz, cx, cy, z, y, x
    return a.transpose(0, 3, 2, 4, 1, 5).reshape(w.size_z * 32, 256, w.size_x * 32)


def chunks_seeing(w, changed):
    """(Z, Y, X) mask of changed cells -> per-slot flags of the chunks that have one inside their one-voxel halo."""
    d = changed.copy()
    for axis in range(3):
        lo = [slice(None)] * 3; hi = [slice(None)] * 3
        lo[axis] = slice(0, -1); hi[axis] = slice(1, None)
        e = d.copy()
        e[tuple(lo)] |= d[tuple(hi)]
        e[tuple(hi)] |= d[tuple(lo)]
        d = e
    per = d.reshape(w.size_z, 32, 4, 64, w.size_x, 32).any(axis=(1, 3, 5))   # cz, cy, cx
    return per.transpose(0, 2, 1).reshape(-1).astype(np.uint8)               # slot = (cz * size_x + cx) * 4 + cy


def interior_only(w, flags):
    f = flags.reshape(w.size_z, w.size_x, 4).copy()
    f[0] = 0; f[-1] = 0; f[:, 0] = 0; f[:, -1] = 0
    return [(cx, cy, cz) for cz in range(w.size_z) for cx in range(w.size_x) for cy in range(4) if f[cz, cx, cy]]


def device_world(ctx, host):
    """blocks only up, surface + light computed on the device, every interior chunk meshed once."""
    from gamecraft_b200 import mesher

    world = mesher.World(ctx, host.size_x, host.size_z).upload(host)
    world.compute_surface().light()
    meshes = mesher.rebuild_chunks(world, host.interior_chunks())
    return world, meshes


def snapshot(world, like):
    out = util.HostWorld(like.size_x, like.size_z)
    world.download(out, blocks=True)
    return out


def oracle_edit(h, verts, edit):
    """One edit on the host copy: status, FlagChunkForUpdate flags, and the from-scratch fill of the edited blocks."""
    flagged = np.zeros(h.size_x * h.size_z * 4, np.uint8)
    before_sun, before_light = to_window(h, effective_sun(h)), to_window(h, h.light).copy()
    status = ob.set_block(h, verts, *edit, flagged)
    if status == APPLIED:
        h.sun[:], h.light[:] = ob.light_fill(h, h.surface)
        changed = (to_window(h, effective_sun(h)) != before_sun) | (to_window(h, h.light) != before_light)
        flagged |= chunks_seeing(h, changed)
    return status, flagged


@pytest.mark.parametrize("kind,seed", [("forest", 1), ("noise", 2), ("mixed", 3), ("islands", 8)])
def test_edits_relight_and_flag_like_the_oracle(ctx, kind, seed):
    from gamecraft_b200 import mesher

    size = 5
    kw = dict(radius=90, falloff=50) if kind in ("islands", "mixed") else {}
    src = util.HostWorld(size, size).generate(kind, seed, **kw)
    blank = util.HostWorld(size, size)
    blank.blocks[:] = src.blocks
    world, meshes = device_world(ctx, blank)
    h = snapshot(world, blank)
    assert np.array_equal(h.surface, ob.compute_surface(h))
    coords = h.interior_chunks()
    verts = world.vertex_counts().reshape(-1)
    for (cx, cy, cz), m in zip(coords, meshes):
        assert verts[(cz * size + cx) * 4 + cy] == m.vert_count
    edge = verts.reshape(size, size, 4)
    assert (edge[0] == -1).all() and (edge[:, 0] == -1).all()                # never meshed
    current = {tuple(c): m for c, m in zip(coords.tolist(), meshes)}
    rng = np.random.default_rng(seed)
    extra = [WATER, GLASS, MAGMA, 21, 13]
    for step in range(8):
        edits = []
        for k in range(1 if step % 3 else 3):                                # single edits and short batches
            wx, wy, wz, b = random_edit(rng, h, size)
            if (step + k) % 4 == 1:
                b = extra[(step + k) % len(extra)]
            edits.append((wx, wy, wz, b))
        if step == 5:
            # the world's top layer: y = 255 is outside ComputeSurface's walk (Lighting.cpp:9), an emitter there is seeded only
            # when a neighbouring column's surface reaches 255 (ComputeMaxY, Lighting.cpp:28-67)
            edits = [(70, 255, 70, LANTERN), (71, 254, 70, STONE), (70, 255, 71, MAGMA), (70, 254, 70, GLASS)]
        want_status, want_flags = [], np.zeros(size * size * 4, np.uint8)
        for e in edits:
            st, fl = oracle_edit(h, verts, e)
            want_status.append(st); want_flags |= fl
        status, rebuild = world.set_blocks(edits)
        assert status.tolist() == want_status, "step %d %r" % (step, edits)
        got = snapshot(world, h)
        assert np.array_equal(got.blocks, h.blocks)
        assert np.array_equal(got.surface, h.surface)
        assert np.array_equal(got.sun, h.sun), "step %d: %d sunlight cells differ" % (step, int((got.sun != h.sun).sum()))
        assert np.array_equal(got.light, h.light), "step %d: %d blockLight cells differ" % (step, int((got.light != h.light).sum()))
        assert [tuple(c) for c in rebuild.tolist()] == sorted(interior_only(h, want_flags), key=lambda c: (c[2], c[0], c[1]))
        # the flagged chunks mesh to the oracle's bytes of the edited world ...
        if len(rebuild):
            orc = ob.OracleWorld(h)
            for c, m in zip(rebuild.tolist(), mesher.rebuild_chunks(world, rebuild)):
                util.assert_same_mesh(m, orc.build_chunk(*c), "step %d chunk %r" % (step, c))
                current[tuple(c)] = m
        # ... and every other chunk to the bytes it had before
        if step in (2, 7):
            orc = ob.OracleWorld(h)
            for c, m in current.items():
                util.assert_same_mesh(m, orc.build_chunk(*c), "step %d unflagged chunk %r" % (step, c))
        verts = world.vertex_counts().reshape(-1)
    world.close()


def test_edit_equals_a_fresh_fill_on_the_device(ctx):
    """Same blocks, two routes on the device: edits with the regional refill vs upload + gcmesh_world_light from scratch."""
    from gamecraft_b200 import mesher

    size = 6
    src = util.HostWorld(size, size).generate("mixed", 5, radius=90, falloff=50)
    blank = util.HostWorld(size, size)
    blank.blocks[:] = src.blocks
    world, _ = device_world(ctx, blank)
    h = snapshot(world, blank)
    rng = np.random.default_rng(11)
    edits = [random_edit(rng, h, size) for _ in range(12)]
    # window corners of the interior: the refill box hangs over the window edge there
    edits += [(32, 70, 32, LANTERN), (32 * (size - 1) - 1, 40, 32 * (size - 1) - 1, 0), (32, 254, 32 * (size - 1) - 1, STONE)]
    status, rebuild = world.set_blocks(edits)
    got = snapshot(world, h)
    fresh = util.HostWorld(size, size)
    fresh.blocks[:] = got.blocks
    w2 = mesher.World(ctx, size, size).upload(fresh)
    w2.compute_surface().light()
    want = snapshot(w2, h)
    w2.close(); world.close()
    assert (status == APPLIED).sum() >= 10
    assert np.array_equal(got.surface, want.surface)
    assert np.array_equal(got.sun, want.sun)
    assert np.array_equal(got.light, want.light)
    assert len(rebuild) > 0


def test_graph_replay_equals_plain_launches(ctx, monkeypatch):
    """The per-edit kernel sequence replayed as a CUDA graph (default) and launched kernel by kernel (GCMESH_EDIT_GRAPH=0)."""
    size = 5
    src = util.HostWorld(size, size).generate("noise", 9)
    blank = util.HostWorld(size, size)
    blank.blocks[:] = src.blocks
    rng = np.random.default_rng(3)
    results = []
    for graph in ("1", "0"):
        monkeypatch.setenv("GCMESH_EDIT_GRAPH", graph)
        world, _ = device_world(ctx, blank)
        h = snapshot(world, blank)
        if not results:
            edits = [random_edit(rng, h, size) for _ in range(10)]
        status, rebuild = world.set_blocks(edits[:1])
        status2, rebuild2 = world.set_blocks(edits[1:])
        results.append((status.tolist() + status2.tolist(), rebuild.tolist(), rebuild2.tolist(), snapshot(world, h)))
        world.close()
    (s0, r0, q0, a0), (s1, r1, q1, a1) = results
    assert s0 == s1 and r0 == r1 and q0 == q1
    for x, y in ((a0.blocks, a1.blocks), (a0.sun, a1.sun), (a0.light, a1.light), (a0.surface, a1.surface)):
        assert np.array_equal(x, y)


def test_more_edits_than_one_pass_holds(ctx):
    """4 200 edits in one call (the device copy of the list holds 4 096 per pass): statuses against the decision oracle,
    final arrays against a from-scratch fill on the device."""
    from gamecraft_b200 import mesher

    size = 4
    src = util.HostWorld(size, size).generate("forest", 4)
    blank = util.HostWorld(size, size)
    blank.blocks[:] = src.blocks
    world, _ = device_world(ctx, blank)
    h = snapshot(world, blank)
    verts = world.vertex_counts().reshape(-1)
    rng = np.random.default_rng(5)
    edits, want = [], []
    flagged = np.zeros(size * size * 4, np.uint8)
    for i in range(4200):
        wx, wz = int(rng.integers(32, 32 * (size - 1))), int(rng.integers(32, 32 * (size - 1)))
        wy = int(rng.integers(-2, 130))
        b = int(rng.choice([0, 0, STONE, LANTERN, WATER, GLASS]))
        edits.append((wx, wy, wz, b))
        want.append(ob.set_block(h, verts, wx, wy, wz, b, flagged))        # decisions need blocks and vertex counts only
    status, rebuild = world.set_blocks(edits)
    assert status.tolist() == want
    assert {APPLIED, UNCHANGED, IGNORED} <= set(want)
    got = snapshot(world, h)
    assert np.array_equal(got.blocks, h.blocks)
    assert np.array_equal(got.surface, h.surface)
    fresh = mesher.World(ctx, size, size).upload(blank)
    fresh_host = util.HostWorld(size, size)
    fresh_host.blocks[:] = h.blocks
    fresh.upload(fresh_host).compute_surface().light()
    ref = snapshot(fresh, h)
    fresh.close(); world.close()
    assert np.array_equal(got.sun, ref.sun)
    assert np.array_equal(got.light, ref.light)
    assert len(rebuild) > 0


def test_refusals(ctx):
    """Never-built chunk, unchanged block, y outside the world, overflow (AIR left behind), edge columns, list capacity."""
    from gamecraft_b200 import mesher

    # centre chunk (1,1,1) with 1666 isolated voxels = 9 996 quads = 39 984 vertices: 16 vertices short of the cap
    cells = [(x, y, z) for y in range(1, 63, 2) for z in range(1, 31, 2) for x in range(1, 29, 3)]
    host = util.HostWorld(3, 3)
    for i in range(1666):
        x, y, z = cells[i]
        host.set_block(32 + x, 64 + y, 32 + z, STONE)
    world = mesher.World(ctx, 3, 3).upload(host)
    world.compute_surface().light()
    status, rebuild = world.set_blocks([(40, 70, 40, STONE)])
    assert status.tolist() == [NOT_BUILT] and len(rebuild) == 0           # nothing meshed yet (World.cpp:562-565)
    mesher.rebuild_chunks(world, [(1, 1, 1)])
    assert world.vertex_counts()[1, 1, 1] == 1666 * 24
    x, y, z = cells[1700]
    free = (32 + x, 64 + y, 32 + z)
    status, rebuild = world.set_blocks([free + (STONE,), (free[0], 300, free[2], STONE), (free[0], -1, free[2], STONE)])
    # 39 984 + 24 > 40 000: refused, and the cell holds AIR
    assert status.tolist() == [OVERFLOW, IGNORED, IGNORED] and len(rebuild) == 0
    got = snapshot(world, host)
    assert np.array_equal(got.blocks, host.blocks)
    x, y, z = cells[0]
    status, rebuild = world.set_blocks([(32 + x, 64 + y, 32 + z, STONE)])
    assert status.tolist() == [UNCHANGED] and len(rebuild) == 0
    status, rebuild = world.set_blocks([(32 + x + 1, 64 + y, 32 + z, STONE)])   # touches one voxel: 5 faces, 20 vertices > 16 left
    assert status.tolist() == [OVERFLOW]
    status, rebuild = world.set_blocks([(32 + x, 64 + y, 32 + z, 0)])           # removing never overflows
    assert status.tolist() == [APPLIED] and [1, 1, 1] in rebuild.tolist()
    with pytest.raises(mesher.GcMeshError):
        world.set_blocks([(5, 70, 40, STONE)])                                # edge column (World.cpp:560)
    with pytest.raises(mesher.GcMeshError):
        world.set_blocks([(40, 70, 32 * 3 + 1, STONE)])                       # outside the window
    # a list too short for the flagged chunks: the count is reported with GC_ERR_CAPACITY
    mesher.rebuild_chunks(world, host.interior_chunks())
    with pytest.raises(mesher.GcMeshError):
        world.set_blocks([(48, 64, 48, STONE)], rebuild_capacity=1)           # y = 64: two chunks of the column see it
    world.close()


def test_edit_after_a_sparse_upload_is_refused_until_the_light_is_whole(ctx):
    """A sparse upload leaves out (and clears) light arrays no mesh depends on, so the window does not hold the fixpoint the
    relight builds on: the edit is refused with GC_ERR_STATE; gcmesh_world_light, or a full upload of both light arrays,
    makes the window editable again and the result equals the from-scratch fill."""
    from gamecraft_b200 import mesher

    host = util.HostWorld(4, 4).build("forest", 5)
    world = mesher.World(ctx, 4, 4).upload_sparse(host)
    mesher.rebuild_chunks(world, host.interior_chunks())
    with pytest.raises(mesher.GcMeshError) as e:
        world.set_blocks([(50, 200, 50, STONE)])
    assert e.value.status == -4 and "sparse" in str(e.value)                  # GC_ERR_STATE
    world.upload(host)                                                        # every light array replaced
    status, _ = world.set_blocks([(50, 200, 50, STONE)])
    assert status.tolist() == [APPLIED]
    world.upload_sparse(host)
    with pytest.raises(mesher.GcMeshError):
        world.set_blocks([(50, 200, 50, 0)])
    world.compute_surface().light()                                           # the device's own fill
    status, _ = world.set_blocks([(50, 200, 50, LANTERN)])
    assert status.tolist() == [APPLIED]
    got = snapshot(world, host)
    host.set_block(50, 200, 50, LANTERN)
    assert np.array_equal(got.blocks, host.blocks)
    surface = ob.compute_surface(host)
    sun, light = ob.light_fill(host, surface)
    dl = util.HostWorld(4, 4); world.download(dl, blocks=False, light=True)
    assert np.array_equal(dl.sun, sun) and np.array_equal(dl.light, light)
    world.close()


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")
@pytest.mark.parametrize("kind,seed,kw,exact", [("forest", 2, {}, True), ("islands", 8, dict(radius=90, falloff=50), False)])
def test_device_edits_against_the_reference_s_own_setblock(ctx, kind, seed, kw, exact):
    """The documented f3 divergence, bounded on the GPU path itself: the same edits through gcmesh_world_set_blocks and
    through the reference's SetBlock -> RecomputeLight compiled verbatim.  Blocks and surface always agree; on forest terrain
    sunlight (as GetSunlight reads it) and block light are identical; after islands edits (under water, surface rises over
    caves) the reference's removal queues may stop short: its sunlight is never above the device's, differs in fewer than
    100 cells of the window and by one level, block light is identical."""
    from gamecraft_b200 import mesher
    from test_edit_oracle import start

    size = 5
    z, r = start(kind, seed, size, **kw)
    world = mesher.World(ctx, size, size).upload(z)                            # the reference's own state: blocks, surface, light
    mesher.rebuild_chunks(world, z.interior_chunks())
    rng = np.random.default_rng(seed)
    applied = 0
    for e in range(40):
        edit = random_edit(rng, z, size)
        changed = r.set_block(*edit)
        status, _ = world.set_blocks([edit])
        assert (status[0] == APPLIED) == changed
        applied += int(changed)
        r.download(z)
    assert applied >= 20
    d = util.HostWorld(size, size)
    world.download(d, blocks=True)
    assert np.array_equal(d.blocks, z.blocks)
    assert np.array_equal(d.surface, z.surface)
    ed, ez = effective_sun(d), effective_sun(z)
    assert np.array_equal(d.light, z.light)
    if exact:
        assert np.array_equal(ed, ez)
    else:
        assert not (ez > ed).any()
        short = ez < ed
        assert int(short.sum()) < 100
        assert int((ed.astype(np.int16) - ez)[short].max(initial=0)) <= 1
    r.close(); world.close()


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libgcref.so not built (no reference tree)")
@pytest.mark.parametrize("kind,seed,kw", [("forest", 2, {}), ("islands", 8, dict(radius=90, falloff=50)), ("mixed", 3, dict(radius=90, falloff=50))])
def test_edit_replay_mode_equals_the_reference_to_the_byte(ctx, kind, seed, kw):
    """GC_WORLD_EDIT_REPLAY: gcmesh_world_set_blocks replays the reference's own RecomputeLight node for node on the device.
    Against the reference's SetBlock compiled verbatim, edit after edit (single edits and batches): block ids, the surface map,
    BOTH RAW LIGHT ARRAYS and the set of chunks flagged for rebuild (the reference's pendingUpdate set over the meshed chunks) are
    identical — on islands too, where the default relight (the from-scratch fill) differs from the reference by design, and with
    weak and strong emitters mixed.  The flagged chunks then mesh to the oracle's bytes of that state."""
    from gamecraft_b200 import mesher
    from test_edit_oracle import start

    size = 5
    z, r = start(kind, seed, size, **kw)
    world = mesher.World(ctx, size, size).upload(z).set_option(mesher.WORLD_EDIT_REPLAY, 1)
    mesher.rebuild_chunks(world, z.interior_chunks())
    rng = np.random.default_rng(seed)
    d = util.HostWorld(size, size)
    applied = 0
    for step in range(24):
        edits = [random_edit(rng, z, size) for _ in range(1 if step % 4 else 3)]
        if step % 6 == 5:
            edits = [e[:3] + (0,) for e in edits]                              # take blocks (emitters among them) away again
        changed = [r.set_block(*e) for e in edits]
        status, rebuild = world.set_blocks(edits)
        assert [s == APPLIED for s in status.tolist()] == changed
        applied += sum(changed)
        r.download(z)
        pending = r.pending_updates(clear=True)                                # (size_z, size_x, 4)
        world.download(d, blocks=True)
        where = "%s step %d %r" % (kind, step, edits)
        assert np.array_equal(d.blocks, z.blocks), where
        assert np.array_equal(d.surface, z.surface), where
        assert np.array_equal(d.sun, z.sun), "%s: %d sunlight bytes differ" % (where, int((d.sun != z.sun).sum()))
        assert np.array_equal(d.light, z.light), "%s: %d blockLight bytes differ" % (where, int((d.light != z.light).sum()))
        want = sorted((cx, cy, cz) for cz in range(1, size - 1) for cx in range(1, size - 1) for cy in range(4) if pending[cz, cx, cy])
        assert sorted(tuple(c) for c in rebuild.tolist()) == want, where
        if len(rebuild):
            orc = ob.OracleWorld(z)
            for c, m in zip(rebuild.tolist(), mesher.rebuild_chunks(world, rebuild)):
                util.assert_same_mesh(m, orc.build_chunk(*c), "%s chunk %r" % (where, c))
    assert applied >= 12
    r.close(); world.close()

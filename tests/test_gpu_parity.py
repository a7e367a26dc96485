"""Parity proper: the CUDA path, called through the C ABI (libgcmesh.so), against the oracle and the golden vectors."""
import ctypes

import numpy as np
import pytest

import gcutil as util
from oracle import binding as ob

pytestmark = pytest.mark.gpu

GOLDEN = util.load_golden()


@pytest.fixture(scope="module")
def ctx():
    from gamecraft_b200 import mesher

    c = mesher.Context(0)
    yield c
    c.close()


def gpu_meshes(ctx, host, coords, max_quads=None):
    from gamecraft_b200 import mesher

    world = mesher.World(ctx, host.size_x, host.size_z).upload(host)
    batch = mesher.Batch(ctx, max(1, len(coords)), max_quads or max(1, len(coords)) * mesher.MAX_QUADS)
    batch.submit(world, coords).wait()
    out = batch.meshes()
    desc, q = batch.results()
    desc = desc.copy()
    batch.close(); world.close()
    return out, desc, q


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_gpu_matches_golden_and_oracle(ctx, key):
    entry = GOLDEN[key]
    w = util.golden_world(entry)
    assert util.input_digest(w) == entry["inputs"]
    coords = w.interior_chunks(1)
    meshes, desc, q = gpu_meshes(ctx, w, coords)
    o = ob.OracleWorld(w)
    assert q == sum(v[0] for v in entry["chunks"].values()) // 4
    for (cx, cy, cz), m in zip(coords, meshes):
        assert util.mesh_signature(m) == entry["chunks"]["%d,%d,%d" % (cx, cy, cz)], "%s %s" % (key, (cx, cy, cz))
        util.assert_same_mesh(m, o.build_chunk(cx, cy, cz), "%s %s" % (key, (cx, cy, cz)))


@pytest.mark.parametrize("kind,size,seed", [("mixed", 6, 31), ("islands", 6, 12), ("noise", 4, 8), ("forest", 9, 2)])
def test_gpu_matches_oracle_other_seeds(ctx, kind, size, seed):
    kw = dict(radius=110, falloff=60) if kind in ("mixed", "islands") else {}
    w = util.HostWorld(size, size + 1, origin=(-3, -3) if kw else (0, 0)).build(kind, seed, **kw)   # non-square window
    coords = w.interior_chunks(1)
    meshes, desc, q = gpu_meshes(ctx, w, coords)
    o = ob.OracleWorld(w)
    for (cx, cy, cz), m in zip(coords, meshes):
        util.assert_same_mesh(m, o.build_chunk(cx, cy, cz), "%s %s" % (kind, (cx, cy, cz)))


def test_gpu_empty_and_ragged_batches(ctx):
    from gamecraft_b200 import mesher

    w = util.HostWorld(4, 4).build("mixed", 3, radius=100, falloff=60)
    world = mesher.World(ctx, 4, 4).upload(w)
    batch = mesher.Batch(ctx, 8, 80000)
    batch.submit(world, np.zeros((0, 3), np.int32)).wait()          # empty submission
    desc, q = batch.results()
    assert len(desc) == 0 and q == 0 and batch.launches() == 0
    o = ob.OracleWorld(w)
    coords = np.array([(1, 0, 1), (2, 3, 2), (1, 0, 1), (2, 0, 1), (1, 0, 1)], np.int32)   # duplicates, odd count
    for m, c in zip(batch.submit(world, coords).wait().meshes(), coords):
        util.assert_same_mesh(m, o.build_chunk(*c), str(c))
    m = mesher.build_chunk(world, 2, 1, 2)                             # the single-chunk call site
    util.assert_same_mesh(m, o.build_chunk(2, 1, 2))
    batch.close(); world.close()


def test_gpu_all_air_world(ctx):
    w = util.empty_world(3)
    meshes, desc, q = gpu_meshes(ctx, w, w.interior_chunks(1))
    assert q == 0 and all(m.vert_count == 0 for m in meshes) and not desc["flags"].any()


def test_gpu_cap_edges_and_overflow_flag(ctx):
    from gamecraft_b200 import mesher

    w = util.cap_edge_world(pairs=1, singles=1665)                     # exactly 10 000 quads
    meshes, desc, q = gpu_meshes(ctx, w, [(1, 1, 1)])
    assert desc["flags"][0] == 0 and meshes[0].vert_count == 40000
    util.assert_same_mesh(meshes[0], ob.OracleWorld(w).build_chunk(1, 1, 1))
    w = util.cap_edge_world(pairs=2, singles=1664)                     # 10 004 quads
    meshes, desc, q = gpu_meshes(ctx, w, [(1, 1, 1), (1, 0, 1)])
    assert desc["flags"][0] == mesher.CHUNK_OVERFLOW and desc["quads"][0] == 10004 and meshes[0].vert_count == 0
    assert desc["flags"][1] == 0
    w = util.HostWorld(3, 3).build("stress", 1)                         # 196 608 quads per chunk
    meshes, desc, q = gpu_meshes(ctx, w, w.interior_chunks(1), max_quads=1000)
    assert (desc["flags"] == mesher.CHUNK_OVERFLOW).all() and (desc["quads"] == 196608).all() and q == 0


def test_gpu_uncapped_batch_meshes_what_the_cap_refuses(ctx):
    """GC_BATCH_UNCAPPED (the stress variant — no reference behaviour, no parity claim against the reference,
    which asserts at its cap): a full 3-D checkerboard chunk — 196 608 quads — is meshed with 32-bit indices; checked against the
    oracle's traversal with the cap lifted, and structurally (every quad indexed once with SetIndices' pattern, Mesh.cpp:41-47)."""
    from gamecraft_b200 import mesher

    w = util.HostWorld(3, 3).build("stress", 1)
    coords = w.interior_chunks(1)
    world = mesher.World(ctx, 3, 3).upload(w)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 196608, uncapped=True)
    o = ob.OracleWorld(w)
    pat = np.array([2, 1, 0, 3, 2, 0], np.int64)
    for _ in range(2):
        batch.submit(world, coords).wait()
        desc, q = batch.results()
        assert q == 4 * 196608 and not desc["flags"].any() and (desc["quads"] == 196608).all() and (desc["vert_count"] == 786432).all()
        for c, m in zip(coords, batch.meshes()):
            want = o.build_chunk_uncapped(*c)
            assert m.indices[0].dtype == np.uint32
            util.assert_same_mesh(m, want, "uncapped %s" % (c,))
            seg = m.indices[0].astype(np.int64).reshape(-1, 6)
            assert np.array_equal(seg, seg[:, 2:3] + pat[None, :]) and np.array_equal(seg[:, 2], np.arange(196608) * 4)
    md = ctypes.create_string_buffer(480000 + 5 * 120000 + 24)                 # GcMeshData is the reference's capped, 16-bit MeshData
    assert mesher.lib().gcmesh_batch_fill_meshdata(batch.h, 0, md) == -4          # GC_ERR_STATE
    batch.close(); world.close()
    # below the cap an uncapped batch writes the capped form's bytes, with the indices widened; every stream is exercised
    w = util.HostWorld(5, 5, origin=(-2, -2)).build("mixed", 11, radius=100, falloff=60)
    coords = w.interior_chunks(1)
    world = mesher.World(ctx, 5, 5).upload(w)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 12000, uncapped=True)
    o = ob.OracleWorld(w)
    streams = np.zeros(5, np.int64)
    for c, m in zip(coords, batch.submit(world, coords).wait().meshes()):
        want = o.build_chunk_uncapped(*c)
        util.assert_same_mesh(m, want, "uncapped %s" % (c,))
        streams += [len(i) for i in m.indices]
    assert (streams > 0).all()
    # a world of many batches per chunk and several mesh types: noise at several times the cap
    w = util.HostWorld(3, 3)
    rng = np.random.default_rng(4)
    w.blocks[:] = rng.choice(np.array([0, 0, 0, 3, 8, 13, 22, 17, 10], np.uint16), w.blocks.shape)
    util.light_world(w)
    coords = w.interior_chunks(1)
    world2 = mesher.World(ctx, 3, 3).upload(w)
    batch2 = mesher.Batch(ctx, len(coords), len(coords) * 196608, uncapped=True)
    o = ob.OracleWorld(w)
    for c, m in zip(coords, batch2.submit(world2, coords).wait().meshes()):
        want = o.build_chunk_uncapped(*c)
        assert want.quads > 30000
        util.assert_same_mesh(m, want, "uncapped noise %s" % (c,))
    batch.close(); world.close(); batch2.close(); world2.close()


def test_gpu_arena_too_small_is_flagged_not_overrun(ctx):
    from gamecraft_b200 import mesher

    w = util.golden_world(GOLDEN["checker5"])
    coords = w.interior_chunks(1)
    meshes, desc, q = gpu_meshes(ctx, w, coords, max_quads=3 * 9216 + 5)
    ok = desc["flags"] == 0
    assert ok.sum() >= 3 and (desc["flags"][~ok] == mesher.CHUNK_NO_SPACE).all() and (~ok).sum() > 0
    o = ob.OracleWorld(w)
    for c, m, f in zip(coords, meshes, ok):
        if f:
            util.assert_same_mesh(m, o.build_chunk(*c))
        else:
            assert m.vert_count == 0


def test_gpu_custom_block_table(ctx):
    from gamecraft_b200 import mesher

    infos, kz = util.custom_table()
    w = util.HostWorld(3, 3)
    rng = np.random.default_rng(5)
    w.blocks[:] = rng.integers(0, len(infos), w.blocks.shape, dtype=np.uint16) * (rng.random(w.blocks.shape) < 0.02)
    w.sun[:] = rng.integers(0, 16, w.sun.shape, dtype=np.uint8)
    w.light[:] = rng.integers(0, 16, w.light.shape, dtype=np.uint8)
    w.compute_surface(threads=1)
    o = ob.OracleWorld(w)
    o.set_block_table(infos, kz)
    c2 = mesher.Context(0)
    c2.set_block_table([mesher.BlockInfo.from_buffer_copy(bytes(b)) for b in infos], kz)
    meshes, desc, q = gpu_meshes(c2, w, w.interior_chunks(1))
    for cy, m in enumerate(meshes):
        want = o.build_chunk(1, cy, 1)
        if want.overflow:
            assert desc["flags"][cy] & mesher.CHUNK_OVERFLOW
        else:
            util.assert_same_mesh(m, want, "cy=%d" % cy)
    c2.close()


def test_gpu_argument_errors(ctx):
    from gamecraft_b200 import mesher

    world = mesher.World(ctx, 4, 4)
    batch = mesher.Batch(ctx, 2, 100)
    with pytest.raises(mesher.GcMeshError):
        batch.submit(world, [(0, 0, 1)])           # edge column (WorldRender.cpp:235)
    with pytest.raises(mesher.GcMeshError):
        batch.submit(world, [(1, 4, 1)])           # chunk row outside 0..3
    with pytest.raises(mesher.GcMeshError):
        batch.submit(world, [(1, 0, 1)] * 3)       # more chunks than the batch holds
    with pytest.raises(mesher.GcMeshError):
        mesher.Batch(ctx, 2, 100).results()        # nothing submitted
    batch.close(); world.close()


def test_gpu_per_chunk_upload_and_fill_meshdata(ctx):
    from gamecraft_b200 import mesher

    w = util.HostWorld(3, 3).build("mixed", 9, radius=100, falloff=60)
    world = mesher.World(ctx, 3, 3)
    for cz in range(3):
        for cx in range(3):
            for cy in range(4):
                s = w.slot(cx, cy, cz)
                world.upload_chunk(cx, cy, cz, w.blocks[s], w.sun[s], w.light[s])
            world.upload_surface(cx, cz, w.surface[w.column(cx, cz)])
    batch = mesher.Batch(ctx, 4, 40000)
    batch.submit(world, [(1, cy, 1) for cy in range(4)]).wait()

    class MeshData(ctypes.Structure):
        _fields_ = [("vertices", ctypes.c_uint8 * 480000), ("indices", (ctypes.c_uint16 * 60000) * 5),
                    ("index_count", ctypes.c_int32 * 5), ("vert_count", ctypes.c_int32)]

    md = MeshData()
    o = ob.OracleWorld(w)
    for cy in range(4):
        assert mesher.lib().gcmesh_batch_fill_meshdata(batch.h, cy, ctypes.byref(md)) == 0
        want = o.build_chunk(1, cy, 1)
        assert md.vert_count == want.vert_count
        assert np.array_equal(np.frombuffer(md.vertices, np.uint8, md.vert_count * 12).reshape(-1, 12), want.vertices)
        for t in range(5):
            assert md.index_count[t] == len(want.indices[t])
            assert np.array_equal(np.frombuffer(md.indices[t], np.uint16, md.index_count[t]), want.indices[t])
    batch.close(); world.close()


def test_gpu_halo_pack_unpack_kernels_match_host_layout(ctx):
    import torch

    from gamecraft_b200 import mesher, slab

    w = util.HostWorld(5, 4).build("noise", 6)
    world = mesher.World(ctx, 5, 4).upload(w)
    buf = torch.zeros(world.halo_bytes(), dtype=torch.uint8, device="cuda")
    assert world.halo_bytes() == slab.halo_bytes(5)
    world.pack_halo(2, 31, buf.data_ptr())
    ctx.synchronize()
    assert np.array_equal(buf.cpu().numpy(), slab.pack_halo_host(w, 2, 31))
    world.unpack_halo(0, 0, buf.data_ptr())            # scatter into another row / plane, then read it back packed
    out = torch.zeros_like(buf)
    world.pack_halo(0, 0, out.data_ptr())
    ctx.synchronize()
    assert torch.equal(out, buf)
    world.close()


def test_gpu_full_size_properties(ctx):
    """BASELINE config 2 at full size (4096 forest chunks): size-independent properties + EVERY chunk byte-compared with the
    oracle (and with the reference's own mesher where oracle/_ref travelled)."""
    from gamecraft_b200 import mesher

    w = util.HostWorld(36, 36, origin=(-18, -18)).build("forest", 1337)
    coords = w.interior_chunks(2)
    assert len(coords) == 4096
    world = mesher.World(ctx, 36, 36).upload(w)
    batch = mesher.Batch(ctx, 4096, 4096 * 2500)
    batch.submit(world, coords).wait()
    desc, q = batch.results()
    desc = desc.copy()
    verts, idx = batch.download()
    assert not desc["flags"].any() and int(desc["quads"].sum()) == q and q > 0
    # segments tile the arena exactly once
    order = np.argsort(desc["quad_base"], kind="stable")
    nz = order[desc["quads"][order] > 0]
    ends = desc["quad_base"][nz] + desc["quads"][nz]
    assert desc["quad_base"][nz][0] == 0 and np.array_equal(desc["quad_base"][nz][1:], ends[:-1]) and ends[-1] == q
    # every chunk: the five index streams reference each quad exactly once with SetIndices' pattern (Mesh.cpp:41-47)
    pat = np.array([2, 1, 0, 3, 2, 0], np.int64)
    for d in desc:
        nq = int(d["quads"])
        if nq == 0:
            continue
        seg = idx[int(d["quad_base"]) * 6: int(d["quad_base"]) * 6 + nq * 6].astype(np.int64).reshape(nq, 6)
        o = seg[:, 2]
        assert np.array_equal(seg, o[:, None] + pat[None, :]) and np.array_equal(np.sort(o), np.arange(nq) * 4)
        assert int(d["index_count"].sum()) == nq * 6 and int(d["vert_count"]) == nq * 4
    # every chunk, every byte, against the oracle
    bad, result, quads = ob.OracleWorld(w).check_batch(coords, desc, verts, idx)
    assert bad == 0 and quads == q, "chunks differing from the oracle: %s" % coords[np.nonzero(result)[0][:8]].tolist()
    if ob.have_ref():
        r = ob.RefWorld(w)
        bad, result, quads = r.check_batch(coords, desc, verts, idx)
        r.close()
        assert bad == 0 and quads == q, "chunks differing from the reference's own mesher: %s" % coords[np.nonzero(result)[0][:8]].tolist()
    # determinism: a second submission (history-based work order, segments elsewhere in the arena) gives identical bytes
    batch.submit(world, coords).wait()
    desc2, _ = batch.results()
    verts2, idx2 = batch.download()
    assert ob.OracleWorld(w).check_batch(coords, desc2.copy(), verts2, idx2)[0] == 0
    batch.close(); world.close()


@pytest.mark.parametrize("kind,cols,seed,radius,min_quads", [("islands", 32, 4242, 512, 1), ("checker", 16, 99, None, 9216), ("noise", 16, 77, None, 5000)])
def test_gpu_full_size_other_workloads_every_chunk(ctx, kind, cols, seed, radius, min_quads):
    """The other bench workloads at bench size (BASELINE configs 3 and 4: islands4096, checker1024, noise1024): every chunk
    byte-compared with the oracle."""
    from gamecraft_b200 import mesher
    from gamecraft_b200.worldgen import INT_MAX

    size = cols + 4
    w = util.HostWorld(size, size, origin=(-(size // 2), -(size // 2))).build(kind, seed, radius=radius or INT_MAX)
    coords = w.interior_chunks(2)
    world = mesher.World(ctx, size, size).upload(w)
    batch = mesher.Batch(ctx, len(coords), len(coords) * mesher.MAX_QUADS)
    for _ in range(2):                                   # first submission: static work order; second: by quad history
        batch.submit(world, coords).wait()
        desc, q = batch.results()
        verts, idx = batch.download()
        assert not desc["flags"].any()
        bad, result, quads = ob.OracleWorld(w).check_batch(coords, desc.copy(), verts, idx)
        assert bad == 0 and quads == q, "%s: chunks differing from the oracle: %s" % (kind, coords[np.nonzero(result)[0][:8]].tolist())
        assert desc["quads"].max() >= min_quads
    if kind == "islands":
        assert (desc["index_count"][:, 3] > 0).any()     # the FLUID stream is exercised (water top faces)
    batch.close(); world.close()


def test_gpu_sparse_upload_equals_dense(ctx):
    """gcmesh_world_upload_sparse leaves the device world exactly as a full upload would (pageable and pinned host arrays,
    host scan and caller hint, and an array that turns zero between uploads)."""
    import torch

    from gamecraft_b200 import mesher

    w = util.HostWorld(5, 5, origin=(-2, -2)).build("mixed", 23, radius=100, falloff=60)
    coords = w.interior_chunks(1)
    o = ob.OracleWorld(w)
    want = [o.build_chunk(*c) for c in coords]
    batch = mesher.Batch(ctx, len(coords), len(coords) * mesher.MAX_QUADS)

    def check(world):
        for c, m, x in zip(coords, batch.submit(world, coords).wait().meshes(), want):
            util.assert_same_mesh(m, x, str(c))

    world = mesher.World(ctx, 5, 5).upload_sparse(w)                      # pageable numpy arrays: memcpy path
    check(world)
    assert world.upload_launches() == 0
    pins = [torch.from_numpy(a).pin_memory() for a in (w.blocks, w.sun, w.light, w.surface)]

    class P:
        size_x, size_z = 5, 5
        blocks, sun, light, surface = [t.numpy() for t in pins]

    world2 = mesher.World(ctx, 5, 5).set_option(mesher.WORLD_COPY_RUN, 1 << 20).upload_sparse(P)   # pinned arrays, every array through the GPU pull kernel
    assert world2.upload_launches() == 1
    check(world2)
    world4 = mesher.World(ctx, 5, 5).set_option(mesher.WORLD_COPY_RUN, 2).upload_sparse(P)         # runs of two or more columns as strided copies, the rest pulled
    assert 0 < world4.upload_bytes() <= world2.upload_bytes()               # same arrays, a shorter (or no) pull list
    check(world4)
    world4.close()
    hint = (P.blocks.any(axis=1) * 1 + P.sun.any(axis=1) * 2 + P.light.any(axis=1) * 4).astype(np.uint8)
    world3 = mesher.World(ctx, 5, 5).upload_sparse(P, nonzero_hint=hint)   # caller's hint instead of the scan
    check(world3)
    # a chunk is cleared on the host: the next sparse upload must clear it on the device too
    s = w.slot(2, 0, 2)
    P.blocks[s] = 0; P.sun[s] = 0; P.light[s] = 0
    w.blocks[s] = 0; w.sun[s] = 0; w.light[s] = 0
    want = [o.build_chunk(*c) for c in coords]
    world2.upload_sparse(P)
    check(world2)
    batch.close(); world.close(); world2.close(); world3.close()


def test_gpu_mesh_host_pipeline(ctx):
    """gcmesh_mesh_host (host arrays in, host meshes out, strips pipelined on three streams) against the oracle: pinned and
    pageable host arrays, scan and hint, caller-order descriptors, repeated calls."""
    import torch

    from gamecraft_b200 import mesher

    w = util.HostWorld(6, 11, origin=(-3, -5)).build("mixed", 41, radius=140, falloff=80)     # 11 rows: three strips, the last partial
    coords = w.interior_chunks(1)
    rng = np.random.default_rng(3)
    coords = coords[rng.permutation(len(coords))]                                            # caller order is not strip order
    o = ob.OracleWorld(w)
    want = [o.build_chunk(*c) for c in coords]
    total = sum(m.quads for m in want)
    pins = [torch.from_numpy(a).pin_memory() for a in (w.blocks, w.sun, w.light, w.surface)]

    class P:
        size_x, size_z = 6, 11
        blocks, sun, light, surface = [t.numpy() for t in pins]

    world = mesher.World(ctx, 6, 11)
    batch = mesher.Batch(ctx, len(coords), total + 10)
    hv = torch.zeros((total * 4, 12), dtype=torch.uint8).pin_memory().numpy()
    hi = torch.zeros(total * 6, dtype=torch.uint16).pin_memory().numpy()
    hint = (P.blocks.any(axis=1) * 1 + P.sun.any(axis=1) * 2 + P.light.any(axis=1) * 4).astype(np.uint8)
    for host, kw in ((P, {}), (P, {"nonzero_hint": hint}), (w, {}), (P, {"threads": 3})):
        hv[:] = 0; hi[:] = 0
        batch.mesh_host(world, host, coords, hv, hi, **kw)
        desc, q = batch.results()
        assert q == total and not desc["flags"].any()
        for d, x, c in zip(desc, want, coords):
            qb, nv = int(d["quad_base"]), int(d["vert_count"])
            got = mesher.ChunkMesh(hv[qb * 4: qb * 4 + nv], [hi[qb * 6 + int(d["index_offset"][t]): qb * 6 + int(d["index_offset"][t]) + int(d["index_count"][t])] for t in range(5)])
            util.assert_same_mesh(got, x, str(c))
    # the ordinary submit path still works on the same batch / world afterwards
    for m, x in zip(batch.submit(world, coords).wait().meshes(), want):
        util.assert_same_mesh(m, x)
    batch.close(); world.close()


def test_gpu_surface_bounds_blocks_option(ctx):
    """GC_WORLD_SURFACE_BOUNDS_BLOCKS: the host scan takes bricks at or above their column's highest surface entry for AIR
    without reading them — except layer y = 255, which ComputeSurface (Lighting.cpp:5-17) never looks at.  Meshes equal the
    oracle's, fewer bytes are scanned, and a block at y = 255 above an otherwise empty sky is still found."""
    import torch

    from gamecraft_b200 import mesher

    w = util.HostWorld(6, 7, origin=(-3, -3)).generate("mixed", 29, radius=110, falloff=60)
    for wx, wz, b in ((70, 75, 3), (71, 75, 14), (100, 130, 22)):            # stone, a lantern and glass on the world's top layer
        w.set_block(wx, 255, wz, b)
    util.light_world(w)
    assert int(w.surface[w.column(2, 2)][(75 & 31) * 32 + (70 & 31)]) < 255  # the surface map does not see them
    coords = w.interior_chunks(1)
    o = ob.OracleWorld(w)
    want = [o.build_chunk(*c) for c in coords]
    total = sum(m.quads for m in want)
    pins = [torch.from_numpy(a).pin_memory() for a in (w.blocks, w.sun, w.light, w.surface)]

    class P:
        size_x, size_z = 6, 7
        blocks, sun, light, surface = [t.numpy() for t in pins]

    world = mesher.World(ctx, 6, 7).set_option(mesher.WORLD_SURFACE_BOUNDS_BLOCKS, 1)
    batch = mesher.Batch(ctx, len(coords), total + 10)
    hv = torch.zeros((total * 4, 12), dtype=torch.uint8).pin_memory().numpy()
    hi = torch.zeros(total * 6, dtype=torch.uint16).pin_memory().numpy()
    for _ in range(2):
        batch.mesh_host(world, P, coords, hv, hi)
        desc, q = batch.results()
        assert q == total
        assert o.check_batch(coords, desc.copy(), hv, hi)[0] == 0
    world.upload_sparse(P)
    for m, x, c in zip(batch.submit(world, coords).wait().meshes(), want, coords):
        util.assert_same_mesh(m, x, str(c))
    with pytest.raises(mesher.GcMeshError):
        world.set_option(99, 1)
    batch.close(); world.close()

    # light of an all-AIR brick that starts above the highest surface entry of the 3 x 3 columns around it is out of every
    # quad's reach (the top row of bricks aside): with the option such arrays are neither scanned nor moved — junk there
    # must not change a byte (forest terrain: everything from y = 64 up)
    f = util.HostWorld(6, 6).build("forest", 31)
    fcoords = f.interior_chunks(1)
    fo = ob.OracleWorld(f)
    fwant = [fo.build_chunk(*c) for c in fcoords]
    smax = f.surface.reshape(6, 6, -1).max(axis=2).astype(int)                              # [cz][cx]
    far = 0
    for cz in range(6):
        for cx in range(6):
            top = smax[max(cz - 1, 0):cz + 2, max(cx - 1, 0):cx + 2].max()
            for cy in range(3):
                sl = f.slot(cx, cy, cz)
                if cy * 64 > top and not f.blocks[sl].any():
                    f.sun[sl] = 5; f.light[sl] = 9
                    far += 1
    assert far >= 36
    fworld = mesher.World(ctx, 6, 6).set_option(mesher.WORLD_SURFACE_BOUNDS_BLOCKS, 1).upload_sparse(f)
    moved = fworld.upload_bytes()
    for m, x, c in zip(mesher.rebuild_chunks(fworld, fcoords), fwant, fcoords):
        util.assert_same_mesh(m, x, str(c))
    fworld.close()
    fworld = mesher.World(ctx, 6, 6).upload_sparse(f)                                       # without the option the junk is moved
    assert fworld.upload_bytes() >= moved + 36 * 65536                                       # (blockLight of the 36 bricks of row cy = 1)
    fworld.close()


def test_gpu_work_order_does_not_change_results(ctx):
    """The library claims chunks heaviest-first (by cy, then by the quad counts of the previous finished submission of the same
    list).  Whatever the caller's order and whatever the internal order, every chunk's mesh is the oracle's."""
    from gamecraft_b200 import mesher

    w = util.HostWorld(6, 6, origin=(-3, -3)).build("mixed", 17, radius=100, falloff=60)
    coords = w.interior_chunks(1)
    o = ob.OracleWorld(w)
    want = {tuple(c): o.build_chunk(*c) for c in coords.tolist()}
    world = mesher.World(ctx, 6, 6).upload(w)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 4000)
    rng = np.random.default_rng(5)
    shuffled = coords[rng.permutation(len(coords))]
    for round_, cs in enumerate((coords, coords, coords, shuffled, shuffled, coords[::-1].copy(), coords[:7].copy(), coords)):
        batch.submit(world, cs).wait()            # 2nd / 3rd submission of a list: ordered by the previous result's quad counts
        desc, q = batch.results()
        assert q == sum(want[tuple(c)].quads for c in cs.tolist())
        for c, m in zip(cs.tolist(), batch.meshes()):
            util.assert_same_mesh(m, want[tuple(c)], "round %d chunk %s" % (round_, c))
    # back-to-back submissions without a wait in between reuse the order already on the device
    for _ in range(3):
        batch.submit(world, shuffled)
    batch.wait()
    for c, m in zip(shuffled.tolist(), batch.meshes()):
        util.assert_same_mesh(m, want[tuple(c)], "back to back %s" % (c,))
    batch.close(); world.close()


def test_gpu_light_that_is_never_read_is_not_transferred(ctx):
    """Stored sunlight at or above a column's surface is ignored (GetSunlight, Lighting.cpp:69-79): the sparse upload skips
    such bricks without scanning them.  Junk written there must not change a single byte of any mesh."""
    import torch

    from gamecraft_b200 import mesher

    w = util.HostWorld(6, 6).build("forest", 31)
    coords = w.interior_chunks(1)
    o = ob.OracleWorld(w)
    want = [o.build_chunk(*c) for c in coords]
    junk = 0
    for col in range(36):
        top = int(w.surface[col].max())
        for cy in range(4):
            if cy * 64 >= top:
                w.sun[col * 4 + cy] = 7          # never read: every cell of the brick lies at or above its column's surface
                junk += 1
    assert junk > 36
    # ... and light of either kind in a brick whose 3 x 3 x 3 brick neighbourhood holds no block at all is out of reach of
    # every quad (VertexLight looks one voxel around the block it lights): the sparse upload skips those arrays as well
    solid = w.blocks.reshape(6, 6, 4, -1).any(axis=3)                       # [cz][cx][cy]
    far = 0
    for cz in range(6):
        for cx in range(6):
            for cy in range(4):
                if not solid[max(cz - 1, 0):cz + 2, max(cx - 1, 0):cx + 2, max(cy - 1, 0):cy + 2].any():
                    w.light[(cz * 6 + cx) * 4 + cy] = 9
                    w.sun[(cz * 6 + cx) * 4 + cy] = 5
                    far += 1
    assert far >= 36
    o2 = ob.OracleWorld(w)
    for c, x in zip(coords, want):
        util.assert_same_mesh(o2.build_chunk(*c), x, "oracle %s" % (c,))
    pins = [torch.from_numpy(a).pin_memory() for a in (w.blocks, w.sun, w.light, w.surface)]

    class P:
        size_x, size_z = 6, 6
        blocks, sun, light, surface = [t.numpy() for t in pins]

    batch = mesher.Batch(ctx, len(coords), len(coords) * 4000)
    for world in (mesher.World(ctx, 6, 6).upload(P), mesher.World(ctx, 6, 6).upload_sparse(P)):
        for c, m, x in zip(coords, batch.submit(world, coords).wait().meshes(), want):
            util.assert_same_mesh(m, x, str(c))
        world.close()
    world = mesher.World(ctx, 6, 6)
    _, q = batch.results()
    hv = np.zeros((q * 4, 12), np.uint8); hi = np.zeros(q * 6, np.uint16)
    batch.mesh_host(world, P, coords, hv, hi)
    desc, q2 = batch.results()
    assert q2 == q
    for (c, d, x) in zip(coords, desc, want):
        v0 = int(d["quad_base"]) * 4
        assert np.array_equal(hv[v0:v0 + int(d["vert_count"])], x.vertices), str(c)
    batch.close(); world.close()


@pytest.mark.parametrize("seed,fill", [(1, 0.03), (2, 0.01), (3, 0.05), (4, 0.002)])
def test_gpu_arbitrary_inputs(ctx, seed, fill):
    """Random ids, random light nibbles, a surface map unrelated to the blocks (inputs no generator would produce)."""
    w = util.arbitrary_world(seed, fill, size=4)
    coords = w.interior_chunks(1)
    meshes, desc, _ = gpu_meshes(ctx, w, coords)
    o = ob.OracleWorld(w)
    for (cx, cy, cz), m, d in zip(coords, meshes, desc):
        want = o.build_chunk(cx, cy, cz)
        if want.overflow:
            assert d["flags"] & 1
        else:
            util.assert_same_mesh(m, want, "seed %d %s" % (seed, (cx, cy, cz)))


def test_gpu_fill_device_buffers(ctx):
    """FillMeshData onto device buffers (gcmesh_batch_fill_device): per chunk one vertex buffer and one element buffer per
    mesh type, as the reference creates them with glBufferData (Mesh.cpp:71-130) — here torch tensors stand in for
    graphics buffers mapped into CUDA.  After gcmesh_submit and after gcmesh_mesh_host (descriptors in the caller's order)."""
    import torch

    from gamecraft_b200 import mesher

    w = util.HostWorld(5, 5).build("mixed", 9, radius=90, falloff=50)
    coords = w.interior_chunks(1)
    o = ob.OracleWorld(w)
    want = [o.build_chunk(*c) for c in coords]
    world = mesher.World(ctx, 5, 5).upload(w)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 6000)

    def check():
        desc, q = batch.results()
        bufs, targets = [], []
        for i, d in enumerate(desc):
            v = torch.full((max(int(d["vert_count"]), 1) * 12 + 16,), 0xEE, dtype=torch.uint8, device="cuda")
            idx = [torch.full((max(int(n), 1) + 8,), 0xEEEE, dtype=torch.uint16, device="cuda") for n in d["index_count"]]
            bufs.append((v, idx))
            # odd chunks get a vertex buffer that is only 4-byte aligned, and no buffer for their opaque index stream
            vp = v.data_ptr() + (4 if i % 2 else 0)
            targets.append((i, vp, [0 if (i % 2 and t == 0) else x.data_ptr() for t, x in enumerate(idx)]))
        batch.fill_device(targets)
        ctx.synchronize()
        torch.cuda.synchronize()
        for i, (d, m) in enumerate(zip(desc, want)):
            v, idx = bufs[i]
            n = int(d["vert_count"]) * 12
            off = 4 if i % 2 else 0
            got = v.cpu().numpy()
            assert np.array_equal(got[off:off + n].reshape(-1, 12), np.asarray(m.vertices).reshape(-1, 12)), "chunk %d vertices" % i
            assert (got[off + n:] == 0xEE).all() and (got[:off] == 0xEE).all()                     # nothing written past the stream
            for t in range(5):
                k = int(d["index_count"][t])
                g = idx[t].cpu().numpy()
                if i % 2 and t == 0:
                    assert (g == 0xEEEE).all()
                else:
                    assert np.array_equal(g[:k], m.indices[t]) and (g[k:] == 0xEEEE).all(), "chunk %d stream %d" % (i, t)

    batch.submit(world, coords).wait()
    check()
    _, q = batch.results()
    pins = [torch.from_numpy(a).pin_memory() for a in (w.blocks, w.sun, w.light, w.surface)]

    class P:
        size_x, size_z = 5, 5
        blocks, sun, light, surface = [t.numpy() for t in pins]

    hv = np.zeros((q * 4, 12), np.uint8); hi = np.zeros(q * 6, np.uint16)
    shuffled = coords[np.random.default_rng(2).permutation(len(coords))]
    want = [o.build_chunk(*c) for c in shuffled]
    batch.mesh_host(world, P, shuffled, hv, hi)
    check()
    with pytest.raises(mesher.GcMeshError):
        batch.fill_device([(len(coords), 0, [0] * 5)])
    batch.close(); world.close()


def test_gpu_validate_blocks(ctx):
    """The reference asserts block < BLOCK_COUNT where a block is stored (World.cpp:294): the device scan counts the ids of a
    window that lie outside the block table and names the first."""
    from gamecraft_b200 import mesher

    w = util.HostWorld(3, 3).generate("mixed", 4)
    world = mesher.World(ctx, 3, 3).upload(w)
    assert world.validate_blocks() == (0, None, None)
    n_types = len(mesher.default_block_table()[0])                           # 24
    bad = [((2, 1, 0), 77, n_types), ((1, 3, 2), 65535, 300), ((1, 3, 2), 4, 0xFF00)]
    for (cx, cy, cz), v, value in bad:
        w.blocks[w.slot(cx, cy, cz)][v] = value
    world.upload(w)
    n, chunk, voxel = world.validate_blocks()
    assert n == 3 and chunk == (2, 1, 0) and voxel == 77                     # slot order: (cz * size_x + cx) * 4 + cy
    w.blocks[w.slot(2, 1, 0)][77] = n_types - 1
    world.upload(w)
    n, chunk, voxel = world.validate_blocks()
    assert n == 2 and chunk == (1, 3, 2) and voxel == 4
    world.close()


def test_gpu_surface_bounds_mesh_option(ctx):
    """GC_WORLD_SURFACE_BOUNDS_MESH: with the surface map being ComputeSurface of the block ids, the kernel answers bricks at or above
    their column's highest entry from the map — same bytes as without the option, on terrain, on a world with blocks on layer
    y = 255 (outside ComputeSurface's walk: the top brick is still looked at), with and without a work order from an earlier
    submission.  With a surface map that lies, the option changes the result (that is its contract); without it, it does not."""
    from gamecraft_b200 import mesher

    for kind, seed, kw in (("forest", 3, {}), ("mixed", 29, dict(radius=110, falloff=60)), ("islands", 8, dict(radius=90, falloff=50))):
        w = util.HostWorld(6, 6, origin=(-3, -3)).generate(kind, seed, **kw)
        for wx, wz, b in ((70, 75, 3), (71, 75, 14), (100, 130, 22)):        # stone, a lantern and glass on the world's top layer
            w.set_block(wx, 255, wz, b)
        util.light_world(w)
        coords = w.interior_chunks(1)
        o = ob.OracleWorld(w)
        want = [o.build_chunk(*c) for c in coords]
        air = sum(1 for m in want if m.vert_count == 0)
        assert air >= len(want) // 3 or kind == "mixed"                        # (the mixed world reaches the top of every column)
        world = mesher.World(ctx, 6, 6).upload(w).set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 1)
        batch = mesher.Batch(ctx, len(coords), sum(m.quads for m in want) + 10)
        for _ in range(3):                                                   # the second and third submission carry a work order
            for m, x, c in zip(batch.submit(world, coords).wait().meshes(), want, coords):
                util.assert_same_mesh(m, x, "%s %s" % (kind, c))
        assert (world.vertex_counts()[1:-1, 1:-1].reshape(-1) == [m.vert_count for m in want]).all()
        batch.close(); world.close()

    # a surface map that claims empty columns: bricks are skipped on its word (option on), or meshed regardless (option off)
    w = util.HostWorld(3, 3).build("forest", 5)
    lie = util.HostWorld(3, 3)
    lie.blocks[:] = w.blocks; lie.sun[:] = w.sun; lie.light[:] = w.light; lie.surface[:] = 0
    truth = ob.OracleWorld(lie).build_chunk(1, 0, 1)
    assert truth.vert_count > 0
    world = mesher.World(ctx, 3, 3).upload(lie)
    util.assert_same_mesh(mesher.rebuild_chunks(world, [(1, 0, 1)])[0], truth)
    world.set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 1)
    assert mesher.rebuild_chunks(world, [(1, 0, 1)])[0].vert_count == 0
    world.close()


@pytest.mark.parametrize("kind,size,seed,kw", [("forest", 9, 3, {}), ("mixed", 8, 29, dict(radius=140, falloff=80)), ("islands", 11, 8, dict(radius=150, falloff=90)),
                                               ("volcanic", 7, 2, {})])
def test_gpu_mesh_host_without_light_arrays(ctx, kind, size, seed, kw):
    """gcmesh_mesh_host with sunlight == block_light == NULL: only block ids cross the link, the light is filled on the device
    group of rows by group of rows (each group a window of its own with a two-row margin), and the meshes equal the oracle's
    meshes of the window lit by the oracle's fill of the WHOLE window — whatever the group boundaries cut through (tree shadows,
    water, emitters next to a boundary).  Called twice: the second call finds the scratch arenas used."""
    import torch

    from gamecraft_b200 import mesher

    w = util.HostWorld(size, size, origin=(-(size // 2), -(size // 2))).generate(kind, seed, **kw)
    if kind == "forest":
        for wz in (127, 128, 191, 192):                                       # lanterns either side of a group boundary (rows 5 | 6) and of its margin (3 | 4)
            w.set_block(100, 60, wz, 14)
            w.set_block(140, 58, wz, 17)
    w.surface[:] = ob.compute_surface(w)
    w.sun[:], w.light[:] = ob.light_fill(w, w.surface)
    coords = w.interior_chunks(1)
    o = ob.OracleWorld(w)
    pins = [torch.from_numpy(a).pin_memory() for a in (w.blocks, w.surface)]

    class P:
        size_x, size_z = size, size
        blocks, surface = [t.numpy() for t in pins]
        sun = None
        light = None

    world = mesher.World(ctx, size, size)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 3000)
    hv = torch.zeros((len(coords) * 3000 * 4, 12), dtype=torch.uint8).pin_memory().numpy()
    hi = torch.zeros(len(coords) * 3000 * 6, dtype=torch.uint16).pin_memory().numpy()
    for rep in range(2):
        if rep == 1:
            world.set_option(mesher.WORLD_SURFACE_BOUNDS_BLOCKS, 1).set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 1)
        batch.mesh_host(world, P, coords, hv, hi)
        desc, q = batch.results()
        bad, result, checked = o.check_batch(coords, desc.copy(), hv, hi)
        assert bad == 0, "%d of %d chunks differ (call %d)" % (bad, len(coords), rep)
        assert checked == q and q > 20000
        assert world.upload_bytes() < w.blocks.nbytes + w.surface.nbytes + 16 * len(coords)   # block ids only
        if kind != "mixed" and rep == 1:
            assert world.upload_bytes() < w.blocks.nbytes // 2                  # and not the air above the terrain
    with pytest.raises(mesher.GcMeshError):
        world.set_blocks([(40, 70, 40, 3)])                                     # the window's own light arenas do not hold the fill
    if kind == "forest":
        # pageable host arrays (plain numpy) and a caller's hint instead of the scan: the same meshes
        class Q:
            size_x, size_z = size, size
            blocks, surface = w.blocks, w.surface
            sun = None
            light = None

        hint = (w.blocks.any(axis=1) * 1).astype(np.uint8)
        pv, pi = np.zeros_like(hv), np.zeros_like(hi)
        for nz in (None, hint):
            batch.mesh_host(world, Q, coords, pv, pi, nonzero_hint=nz)
            desc, q = batch.results()
            assert o.check_batch(coords, desc.copy(), pv, pi)[0] == 0
    P.sun = w.sun
    with pytest.raises(mesher.GcMeshError):
        batch.mesh_host(world, P, coords, hv, hi)                               # one light array without the other
    batch.close(); world.close()


def test_gpu_query_polls_a_submission(ctx):
    """gcmesh_query: the non-blocking form of gcmesh_wait — refuses a batch nothing was submitted to, turns true once the
    submission has finished, and the results are then readable without a wait."""
    import time

    from gamecraft_b200 import mesher

    w = util.HostWorld(6, 6).build("forest", 9)
    world = mesher.World(ctx, 6, 6).upload(w)
    coords = w.interior_chunks(1)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 3000)
    with pytest.raises(mesher.GcMeshError):
        batch.done()
    batch.submit(world, coords)
    t0 = time.time()
    while not batch.done():
        assert time.time() - t0 < 10
    desc, q = batch.results()                                                  # no gcmesh_wait in between
    o = ob.OracleWorld(w)
    assert q == sum(o.build_chunk(*c).quads for c in coords)
    assert batch.done()
    batch.close(); world.close()

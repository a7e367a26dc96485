"""Slab halo exchange over peer memory (gcmesh_world_halo_*): every rank stores its boundary planes straight into its
neighbours' windows; the meshes of the slabs must equal the oracle's meshes of the one global world.

One GPU: three slab worlds of one process, each with its own context and stream (gcmesh_world_halo_connect_local).
Two or more GPUs: one process per GPU over CUDA IPC handles (gcmesh_world_halo_connect_ipc), as bench.py --gpus N runs it."""
import os
import socket
import sys

import numpy as np
import pytest

import gcutil as util
from oracle import binding as ob

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS_PER_RANK, SIZE_X, KIND, SEED = 2, 5, "mixed", 17


def build_window(rank, world_size):
    from gamecraft_b200 import slab
    from gamecraft_b200.worldgen import HostWorld

    first_row, size_z, first_owned, last_owned = slab.slab_window(rank, world_size, ROWS_PER_RANK)
    w = HostWorld(SIZE_X, size_z, origin=(-2, first_row - 2)).build(KIND, SEED, radius=120, falloff=70, threads=1)
    # scribble over the halo planes this rank must RECEIVE: the meshes are wrong unless the exchange restores them
    for peer, _, (rrow, rz) in slab.exchange_plan(rank, world_size, w.size_z):
        slab.unpack_halo_host(w, rrow, rz, np.zeros(slab.halo_bytes(w.size_x), np.uint8))
    return w, first_row, first_owned, last_owned


def global_oracle(world_size):
    from gamecraft_b200 import slab
    from gamecraft_b200.worldgen import HostWorld

    size_z = ROWS_PER_RANK * world_size + 2 * slab.MARGIN
    g = HostWorld(SIZE_X, size_z, origin=(-2, -slab.MARGIN - 2)).build(KIND, SEED, radius=120, falloff=70, threads=1)
    return g, ob.OracleWorld(g)


def owned_coords(w, first_owned, last_owned):
    return np.asarray([(cx, cy, cz) for cz in range(first_owned, last_owned + 1) for cx in range(1, w.size_x - 1) for cy in range(4)], np.int32)


@pytest.mark.parametrize("fused", [False, True])
def test_three_slabs_one_gpu_local_peers(fused):
    """fused = False: gcmesh_world_halo_push (exchange kernel) + gcmesh_submit; True: gcmesh_submit_exchange (one kernel)."""
    from gamecraft_b200 import mesher, slab

    n = 3
    ranks = []
    for r in range(n):
        w, first_row, first_owned, last_owned = build_window(r, n)
        ctx = mesher.Context(0)
        world = mesher.World(ctx, w.size_x, w.size_z).upload(w)
        coords = owned_coords(w, first_owned, last_owned)
        ranks.append(dict(host=w, first_row=first_row, fo=first_owned, lo=last_owned, ctx=ctx, world=world, coords=coords,
                          batch=mesher.Batch(ctx, len(coords), len(coords) * 4096)))
    for r in range(n):
        me = ranks[r]
        if r + 1 < n:
            me["world"].halo_connect_local(1, ranks[r + 1]["world"], ranks[r + 1]["fo"] - 1)
        if r > 0:
            me["world"].halo_connect_local(0, ranks[r - 1]["world"], ranks[r - 1]["lo"] + 1)
    with pytest.raises(mesher.GcMeshError):
        ranks[0]["world"].halo_connect_local(1, ranks[1]["world"], 1)         # already connected
    for c in ranks:
        c["ctx"].synchronize()
    for k in range(3):                                                           # exchange -> mesh, three times over
        for c in ranks:
            if fused:
                c["batch"].submit_exchange(c["world"], c["coords"], c["fo"], c["lo"])
            else:
                c["world"].halo_push(c["fo"], c["lo"])
                c["batch"].submit(c["world"], c["coords"])
    g, oracle = global_oracle(n)
    for c in ranks:
        meshes = c["batch"].wait().meshes()
        timed_out, serial = c["world"].halo_status()
        assert not timed_out and serial == 3, "timed out: 0x%x" % timed_out
        for (cx, cy, cz), m in zip(c["coords"].tolist(), meshes):
            util.assert_same_mesh(m, oracle.build_chunk(cx, cy, c["first_row"] + cz + slab.MARGIN), "chunk %r" % ((cx, cy, cz),))
    for c in ranks:
        c["batch"].close(); c["world"].close()
    for c in ranks:
        c["ctx"].close()


def test_missing_neighbour_times_out_instead_of_hanging():
    """Own process: the wait bound is read once per process (GCMESH_HALO_WAIT_MS)."""
    import subprocess

    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import test_gpu_halo_p2p as t; t.missing_neighbour()\n" % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, GCMESH_HALO_WAIT_MS="300")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "timed out as expected" in r.stdout


def missing_neighbour():
    from gamecraft_b200 import mesher

    w, _, fo, lo = build_window(0, 2)
    ctx_a, ctx_b = mesher.Context(0), mesher.Context(0)
    a = mesher.World(ctx_a, w.size_x, w.size_z).upload(w)
    b = mesher.World(ctx_b, w.size_x, w.size_z).upload(w)
    try:
        a.halo_connect_local(1, mesher.World(ctx_a, w.size_x, w.size_z), fo - 1)   # a neighbour on the same context / stream could never answer
        raise AssertionError("expected GcMeshError")
    except mesher.GcMeshError:
        pass
    a.halo_connect_local(1, b, fo - 1)
    a.halo_push(fo, lo)                                                         # b never calls halo_push
    timed_out, serial = a.halo_status()
    assert timed_out and serial == 1
    assert b.halo_status()[0] == timed_out                                      # ... and the neighbour's flag block says so too
    # the fused form: the boundary planes never arrive -> the chunks that needed them are flagged, gcmesh_wait fails
    coords = owned_coords(w, fo, lo)
    batch = mesher.Batch(ctx_a, len(coords), len(coords) * 4096)
    batch.submit_exchange(a, coords, fo, lo)
    try:
        batch.wait()
        raise AssertionError("expected GcMeshError")
    except mesher.GcMeshError as e:
        assert "never arrived" in str(e)
    desc, _ = batch.results()
    top = coords[:, 2] == lo
    assert (desc["flags"][top] == 4).all() and (desc["vert_count"][top] == 0).all()        # GC_CHUNK_HALO_MISSING
    assert (desc["flags"][~top] == 0).all()
    batch.close()
    try:
        b.halo_push(fo, lo)                                                     # no neighbour connected
        raise AssertionError("expected GcMeshError")
    except mesher.GcMeshError:
        pass
    a.close(); b.close(); ctx_a.close(); ctx_b.close()
    print("timed out as expected: 0x%x" % timed_out)


def _ipc_worker(rank, world_size, port, out_dir, fused):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import json

    import torch
    import torch.distributed as dist

    from gamecraft_b200 import mesher

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    w, first_row, fo, lo = build_window(rank, world_size)
    ctx = mesher.Context(rank)
    world = mesher.World(ctx, w.size_x, w.size_z).upload(w)
    ctx.synchronize()
    handles = [None] * world_size
    dist.all_gather_object(handles, world.halo_export())
    if rank + 1 < world_size:
        world.halo_connect_ipc(1, handles[rank + 1], fo - 1)
    if rank > 0:
        world.halo_connect_ipc(0, handles[rank - 1], lo + 1)
    dist.barrier()
    coords = owned_coords(w, fo, lo)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 4096)
    for k in range(3):
        if fused:
            batch.submit_exchange(world, coords, fo, lo)
        else:
            world.halo_push(fo, lo)
            batch.submit(world, coords)
    meshes = batch.wait().meshes()
    timed_out, serial = world.halo_status()
    sigs = {"timed_out": bool(timed_out), "serial": serial}
    for (cx, cy, cz), m in zip(coords.tolist(), meshes):
        sigs["%d,%d,%d" % (cx, cy, first_row + cz)] = [m.vert_count, ob.fnv1a64(m.vertices) if m.vert_count else "",
                                                       [ob.fnv1a64(i) if len(i) else "" for i in m.indices]]
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(sigs, f)
    dist.barrier()                                                              # nobody unmaps a window a neighbour still writes
    batch.close(); world.close(); ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("fused", [False, True])
def test_one_process_per_gpu_ipc_peers(tmp_path, fused):
    import json

    import torch
    import torch.multiprocessing as mp

    from gamecraft_b200 import slab

    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_ipc_worker, args=(n, port, str(tmp_path), fused), nprocs=n, join=True)
    g, oracle = global_oracle(n)
    seen = 0
    for r in range(n):
        got = json.load(open(tmp_path / ("rank%d.json" % r)))
        assert got.pop("timed_out") is False and got.pop("serial") == 3
        for key, sig in got.items():
            cx, cy, grow = map(int, key.split(","))
            m = oracle.build_chunk(cx, cy, grow + slab.MARGIN)
            want = [m.vert_count, ob.fnv1a64(m.vertices) if m.vert_count else "", [ob.fnv1a64(i) if len(i) else "" for i in m.indices]]
            assert sig == want, key
            seen += 1
    assert seen == n * ROWS_PER_RANK * (SIZE_X - 2) * 4

"""Multi-rank host logic on CPU: two ranks (gloo) each build their z-slab window, exchange the boundary voxel planes
with send/recv in the GPU kernels' buffer layout, and mesh their own rows; the result must equal meshing the one
global world.  (On the GPU box the same plan runs over NCCL with the pack/unpack kernels.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ROWS_PER_RANK = 2
SIZE_X = 5
KIND, SEED = "mixed", 17


def build_window(rank, world_size):
    from gamecraft_b200 import slab
    from gamecraft_b200.worldgen import HostWorld

    first_row, size_z, first_owned, last_owned = slab.slab_window(rank, world_size, ROWS_PER_RANK)
    w = HostWorld(SIZE_X, size_z, origin=(-2, first_row - 2)).build(KIND, SEED, radius=120, falloff=70, threads=1)
    return w, first_row, first_owned, last_owned


def global_world(world_size):
    from gamecraft_b200 import slab
    from gamecraft_b200.worldgen import HostWorld

    size_z = ROWS_PER_RANK * world_size + 2 * slab.MARGIN
    return HostWorld(SIZE_X, size_z, origin=(-2, -slab.MARGIN - 2)).build(KIND, SEED, radius=120, falloff=70, threads=1)


def worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch

    from gamecraft_b200 import slab
    from oracle import binding as ob

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    w, first_row, first_owned, last_owned = build_window(rank, world_size)

    # scribble over the halo planes this rank must RECEIVE, so the test fails unless the exchange restores them
    plan = slab.exchange_plan(rank, world_size, w.size_z)
    for peer, _, (rrow, rz) in plan:
        junk = np.full(slab.halo_bytes(w.size_x), 0, np.uint8)
        slab.unpack_halo_host(w, rrow, rz, junk)

    reqs, recvs = [], []
    for peer, (srow, sz), (rrow, rz) in plan:
        send = torch.from_numpy(slab.pack_halo_host(w, srow, sz))
        recv = torch.empty(slab.halo_bytes(w.size_x), dtype=torch.uint8)
        reqs += [dist.isend(send, peer), dist.irecv(recv, peer)]
        recvs.append((rrow, rz, recv))
    for r in reqs:
        r.wait()
    for rrow, rz, recv in recvs:
        slab.unpack_halo_host(w, rrow, rz, recv.numpy())

    o = ob.OracleWorld(w)
    sigs = {}
    for cz in range(first_owned, last_owned + 1):
        for cx in range(1, w.size_x - 1):
            for cy in range(4):
                m = o.build_chunk(cx, cy, cz)
                sigs["%d,%d,%d" % (cx, cy, first_row + cz)] = [m.vert_count, ob.fnv1a64(m.vertices) if m.vert_count else "",
                                                               [ob.fnv1a64(i) if len(i) else "" for i in m.indices]]
    import json
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(sigs, f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_exchange_matches_global_world(tmp_path):
    import json

    from gamecraft_b200 import slab
    from oracle import binding as ob

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)

    g = global_world(2)
    o = ob.OracleWorld(g)
    got = {}
    for r in range(2):
        got.update(json.load(open(tmp_path / ("rank%d.json" % r))))
    assert len(got) == 2 * ROWS_PER_RANK * (SIZE_X - 2) * 4
    for key, sig in got.items():
        cx, cy, grow = map(int, key.split(","))
        m = o.build_chunk(cx, cy, grow + slab.MARGIN)     # global window starts MARGIN rows before global row 0
        want = [m.vert_count, ob.fnv1a64(m.vertices) if m.vert_count else "", [ob.fnv1a64(i) if len(i) else "" for i in m.indices]]
        assert sig == want, key


def test_exchange_plan_shapes():
    from gamecraft_b200 import slab

    assert slab.exchange_plan(0, 1, 36) == []
    assert slab.exchange_plan(0, 2, 36) == [(1, (33, 31), (34, 0))]
    assert slab.exchange_plan(1, 2, 36) == [(0, (2, 0), (1, 31))]
    assert len(slab.exchange_plan(1, 4, 36)) == 2
    assert slab.halo_bytes(36) == 36 * 4 * 8192 + 36 * 32


def test_pack_unpack_roundtrip():
    from gamecraft_b200 import slab
    from gamecraft_b200.worldgen import HostWorld

    w = HostWorld(4, 4).build("noise", 3, threads=1)
    buf = slab.pack_halo_host(w, 2, 31)
    z = HostWorld(4, 4)
    slab.unpack_halo_host(z, 1, 31, buf)
    for cx in range(4):
        for cy in range(4):
            a, b = w.slot(cx, cy, 2), z.slot(cx, cy, 1)
            assert np.array_equal(w.blocks[a, 31 * 2048:], z.blocks[b, 31 * 2048:])
            assert np.array_equal(w.light[a, 31 * 2048:], z.light[b, 31 * 2048:])
            assert not z.blocks[b, : 31 * 2048].any()

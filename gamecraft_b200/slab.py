"""Slab decomposition of a world across GPUs (host logic) and the boundary-plane buffer layout.

The path shards by z-slabs of whole columns: rank r owns ``rows_per_rank`` rows of columns.  Every rank holds
its rows plus ``MARGIN`` rows each side (one halo ring the mesher reads, one more so the halo ring's light is
final — the reference's 9-column GROUP_PREPROCESSED gate, Code/WorldRender.cpp:255-277).  The only data a
rank needs from a neighbour at run time is the one-voxel boundary plane of the neighbour's first / last owned
row; ``exchange_plan`` names those planes and ``pack_halo_host`` / ``unpack_halo_host`` restate on the host
(numpy) the buffer layout of the GPU pack/unpack kernels (gcmesh_world_pack_halo), for CPU tests of the
multi-rank logic and for checking the kernels.
"""
from __future__ import annotations

import numpy as np

MARGIN = 2
PLANE_BYTES = 8192  # per (column, cy): 2048 u16 blocks + 2048 B sunlight + 2048 B blockLight


def slab_window(rank: int, world_size: int, rows_per_rank: int, margin: int = MARGIN):
    """(first global row of the rank's window, window rows, first/last owned LOCAL row)."""
    first_global = rank * rows_per_rank - margin
    size_z = rows_per_rank + 2 * margin
    return first_global, size_z, margin, margin + rows_per_rank - 1


def exchange_plan(rank: int, world_size: int, size_z: int, margin: int = MARGIN):
    """List of (peer, (send_row, send_z), (recv_row, recv_z)) in LOCAL rows of this rank's window.

    Towards rank+1: send the z = 31 plane of the last owned row, receive the z = 0 plane of the peer's first owned
    row into the row just past the owned range.  Towards rank-1: the mirror image."""
    first_owned, last_owned = margin, size_z - margin - 1
    plan = []
    if rank + 1 < world_size:
        plan.append((rank + 1, (last_owned, 31), (last_owned + 1, 0)))
    if rank > 0:
        plan.append((rank - 1, (first_owned, 0), (first_owned - 1, 31)))
    return plan


def halo_bytes(size_x: int) -> int:
    return size_x * 4 * PLANE_BYTES + size_x * 32


def pack_halo_host(host, cz: int, z: int) -> np.ndarray:
    """Row ``cz`` of columns, voxel plane ``z``: per column cx, per chunk cy -> blocks plane | sun plane | light plane;
    then per column the surface row z.  Same bytes as gcmesh_world_pack_halo."""
    sx = host.size_x
    out = np.empty(halo_bytes(sx), np.uint8)
    for cx in range(sx):
        for cy in range(4):
            s = host.slot(cx, cy, cz)
            o = (cx * 4 + cy) * PLANE_BYTES
            out[o:o + 4096] = host.blocks[s, z * 2048:(z + 1) * 2048].view(np.uint8)
            out[o + 4096:o + 6144] = host.sun[s, z * 2048:(z + 1) * 2048]
            out[o + 6144:o + 8192] = host.light[s, z * 2048:(z + 1) * 2048]
        so = sx * 4 * PLANE_BYTES + cx * 32
        out[so:so + 32] = host.surface[host.column(cx, cz), z * 32:(z + 1) * 32]
    return out


def unpack_halo_host(host, cz: int, z: int, buf: np.ndarray) -> None:
    sx = host.size_x
    for cx in range(sx):
        for cy in range(4):
            s = host.slot(cx, cy, cz)
            o = (cx * 4 + cy) * PLANE_BYTES
            host.blocks[s, z * 2048:(z + 1) * 2048] = buf[o:o + 4096].view(np.uint16)
            host.sun[s, z * 2048:(z + 1) * 2048] = buf[o + 4096:o + 6144]
            host.light[s, z * 2048:(z + 1) * 2048] = buf[o + 6144:o + 8192]
        so = sx * 4 * PLANE_BYTES + cx * 32
        host.surface[host.column(cx, cz), z * 32:(z + 1) * 32] = buf[so:so + 32]

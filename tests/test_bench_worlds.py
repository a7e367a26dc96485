"""bench.py's windows: the slabs of neighbouring ranks overlap in identical voxels (noise is sampled in world space), for
the weak-scaling workloads (fixed rows per rank) and for BASELINE config 5 (one world split over the ranks)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def rows(host, lo, hi):
    per_row = host.size_x * 4
    return host.blocks[lo * per_row:hi * per_row], host.surface[lo * host.size_x:hi * host.size_x]


def test_strong_scaling_slabs_tile_one_world():
    n = 32                                                    # 128 rows over 32 ranks: 4 owned rows + 2 margin rows each side
    assert bench.rows_per_rank("forest65536", n) == 4 and bench.rows_per_rank("forest4096", n) == 32
    a, ca, _ = bench.make_world("forest65536", 5, n)
    b, cb, _ = bench.make_world("forest65536", 6, n)
    assert (a.size_x, a.size_z) == (132, 8) and len(ca) == 128 * 4 * 4
    assert b.origin[1] - a.origin[1] == 4
    # rank 5's rows 4..7 (its last two owned rows and its upper margin) are rank 6's rows 0..3 (lower margin, first owned rows):
    # same block ids and surface map (light near a window's edge row is not final — that is what the halo exchange is for)
    for x, y in zip(rows(a, 4, 8), rows(b, 0, 4)):
        assert np.array_equal(x, y)
    # the owned rows of all ranks cover the 128 rows once: first owned global row = origin + 2
    first = [bench.make_world("forest65536", r, n)[0].origin[1] + 2 for r in (0, n - 1)]
    assert first[1] - first[0] == 4 * (n - 1)

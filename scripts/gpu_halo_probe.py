"""Diagnostic: the peer-memory halo exchange between slab worlds of ONE process (local peers), with the flag blocks printed.
    python scripts/gpu_halo_probe.py [ranks]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from gamecraft_b200 import mesher
from test_gpu_halo_p2p import build_window, owned_coords

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
submit = len(sys.argv) <= 2 or sys.argv[2] != "nosubmit"
ranks = []
for r in range(n):
    w, first_row, fo, lo = build_window(r, n)
    ctx = mesher.Context(0)
    world = mesher.World(ctx, w.size_x, w.size_z).upload(w)
    coords = owned_coords(w, fo, lo)
    ranks.append(dict(fo=fo, lo=lo, ctx=ctx, world=world, coords=coords, batch=mesher.Batch(ctx, len(coords), len(coords) * 4096)))
for r in range(n):
    if r + 1 < n:
        ranks[r]["world"].halo_connect_local(1, ranks[r + 1]["world"], ranks[r + 1]["fo"] - 1)
    if r > 0:
        ranks[r]["world"].halo_connect_local(0, ranks[r - 1]["world"], ranks[r - 1]["lo"] + 1)
for c in ranks:
    c["ctx"].synchronize()
for k in range(3):
    for c in ranks:
        c["world"].halo_push(c["fo"], c["lo"])
        if submit:
            c["batch"].submit(c["world"], c["coords"])
for i, c in enumerate(ranks):
    if submit:
        c["batch"].wait()
    print("rank", i, "status 0x%x serial %d" % c["world"].halo_status())

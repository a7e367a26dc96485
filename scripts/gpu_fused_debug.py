"""development: two slab worlds of one process, fused exchange, under GCMESH_HALO_DEBUG / GCMESH_HALO_TASKS settings (one subprocess each)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import test_gpu_halo_p2p as t
from gamecraft_b200 import mesher
n = 2
ranks = []
for r in range(n):
    w, first_row, fo, lo = t.build_window(r, n)
    ctx = mesher.Context(0)
    world = mesher.World(ctx, w.size_x, w.size_z).upload(w)
    coords = t.owned_coords(w, fo, lo)
    ranks.append(dict(w=w, fo=fo, lo=lo, ctx=ctx, world=world, coords=coords, batch=mesher.Batch(ctx, len(coords), len(coords) * 4096)))
ranks[0]["world"].halo_connect_local(1, ranks[1]["world"], ranks[1]["fo"] - 1)
ranks[1]["world"].halo_connect_local(0, ranks[0]["world"], ranks[0]["lo"] + 1)
for c in ranks: c["ctx"].synchronize()
for k in range(int(os.environ.get("ROUNDS", "1"))):
    for c in ranks:
        c["batch"].submit_exchange(c["world"], c["coords"], c["fo"], c["lo"])
for c in ranks:
    c["batch"].wait()
    d, q = c["batch"].results()
    print("ok quads", q, "flags", sorted(set(d["flags"].tolist())), "status", c["world"].halo_status())
''' % (ROOT, ROOT)
PLAIN = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import test_gpu_halo_p2p as t
from gamecraft_b200 import mesher
w, first_row, fo, lo = t.build_window(0, 1)
ctx = mesher.Context(0)
world = mesher.World(ctx, w.size_x, w.size_z).upload(w)
coords = t.owned_coords(w, fo, lo)
b = mesher.Batch(ctx, len(coords), len(coords) * 4096)
for k in range(3): b.submit(world, coords)
b.wait(); print("plain ok", b.results()[1])
''' % (ROOT, ROOT)
r = subprocess.run([sys.executable, "-c", PLAIN], env=dict(os.environ, GCMESH_FORCE_HALO_KERNEL="1"), capture_output=True, text=True, timeout=300)
print("forced halo kernel on a plain submission: rc", r.returncode, r.stdout.strip(), r.stderr.strip().splitlines()[-1][:200] if r.returncode else "")
for env in ({"GCMESH_HALO_DEBUG": "15"}, {"GCMESH_HALO_DEBUG": "7"}, {"GCMESH_HALO_DEBUG": "3"}, {"GCMESH_HALO_DEBUG": "1"}, {"GCMESH_HALO_DEBUG": "2"}, {"GCMESH_HALO_DEBUG": "6"}, {"GCMESH_HALO_TASKS": "1"}, {}, {"ROUNDS": "3"}):
    e = dict(os.environ, GCMESH_HALO_WAIT_MS="2000", **env)
    r = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True, timeout=300)
    print(env, "rc", r.returncode, "|", r.stdout.strip().replace("\n", " ; "), "|", r.stderr.strip().splitlines()[-1][:200] if r.returncode else "")

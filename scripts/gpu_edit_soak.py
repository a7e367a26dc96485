"""Soak: many random edits (dig / build / emitters / water / glass) on emitter- and water-rich worlds; after every batch the
device arrays must equal a from-scratch surface + light fill of the same blocks on a second device world.
    python scripts/gpu_edit_soak.py [edits per world]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld

n_edits = int(sys.argv[1]) if len(sys.argv) > 1 else 600
ctx = mesher.Context(0)
bad = 0
for kind, seed, kw in [("mixed", 3, dict(radius=90, falloff=50)), ("islands", 8, dict(radius=90, falloff=50)), ("noise", 5, {}), ("forest", 2, {})]:
    size = 5
    src = HostWorld(size, size).generate(kind, seed, **kw)
    blank = HostWorld(size, size); blank.blocks[:] = src.blocks
    world = mesher.World(ctx, size, size).upload(blank)
    world.compute_surface().light()
    mesher.rebuild_chunks(world, blank.interior_chunks())
    rng = np.random.default_rng(seed)
    palette = [0, 0, 0, 3, 3, 14, 17, 21, 8, 13, 22, 10]
    done = 0
    while done < n_edits:
        batch = []
        for _ in range(50):
            wx, wz = int(rng.integers(32, 32 * (size - 1))), int(rng.integers(32, 32 * (size - 1)))
            wy = int(rng.integers(0, 256)) if rng.random() < 0.2 else int(rng.integers(0, 70))
            batch.append((wx, wy, wz, int(rng.choice(palette))))
        status, rebuild = world.set_blocks(batch)
        done += len(batch)
        got = HostWorld(size, size); world.download(got, blocks=True)
        fresh = HostWorld(size, size); fresh.blocks[:] = got.blocks
        w2 = mesher.World(ctx, size, size).upload(fresh)
        w2.compute_surface().light()
        want = HostWorld(size, size); w2.download(want)
        w2.close()
        ds, dl, df = int((got.sun != want.sun).sum()), int((got.light != want.light).sum()), int((got.surface != want.surface).sum())
        if ds or dl or df:
            bad += 1
            print(kind, "after", done, "edits: sun", ds, "light", dl, "surface", df)
        if len(rebuild):
            mesher.rebuild_chunks(world, rebuild)
    print(kind, seed, done, "edits,", int((status == 0).sum()), "of the last 50 applied, mismatching batches so far:", bad)
    world.close()
print("SOAK", "FAILED" if bad else "ok")

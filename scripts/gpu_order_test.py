import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld
host = HostWorld(36, 36, origin=(-18, -18)).build("forest", 1337)
coords = host.interior_chunks(2)
ctx = mesher.Context(0)
world = mesher.World(ctx, 36, 36).upload(host)
batch = mesher.Batch(ctx, len(coords), 4_000_000)
def run(c, name):
    ms = []
    for _ in range(8):
        batch.submit(world, c).wait(); ms.append(batch.elapsed_ms())
    print("%-34s kernel ms: min %.4f  median %.4f" % (name, min(ms), sorted(ms)[len(ms) // 2]))
    return batch.results()[0]["quads"].copy()
q = run(coords, "caller order (z, x, cy)")
run(coords[np.argsort(-q.astype(np.int64), kind="stable")], "descending quads (LPT)")
cy = coords[:, 1]
run(coords[np.argsort(np.where(cy == 1, 0, np.where(cy == 0, 1, 2)), kind="stable")], "cy 1, then 0, then 2 and 3")
run(coords[np.argsort(cy, kind="stable")], "cy ascending")
rng = np.random.default_rng(1); run(coords[rng.permutation(len(coords))], "random")
batch.close(); world.close(); ctx.close()

"""In-kernel phase attribution (GCMESH_PROFILE=1): cycles per chunk per phase and wait times."""
import os, sys
os.environ["GCMESH_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld

kind = sys.argv[1] if len(sys.argv) > 1 else "forest"
host = HostWorld(36, 36, origin=(-18, -18)).build(kind, 1337)
coords = host.interior_chunks(2)
ctx = mesher.Context(0)
world = mesher.World(ctx, 36, 36).upload(host)
batch = mesher.Batch(ctx, len(coords), len(coords) * 2500)
for _ in range(3):
    batch.submit(world, coords).wait()
pr = batch.profile()
n = len(coords)
names = ["P0 next chunk", "P1 stage+convert ids", "P1x x-halo classes", "P2 faces+scan", "P3a quad list", "P3b light gather+vertices"]
tot = sum(pr[:6])
print("kernel ms %.3f  chunks %d  total cycles/chunk (thread 0 of each CTA) %.0f" % (batch.elapsed_ms(), n, tot / n))
for i, nm in enumerate(names):
    print("  %-26s %8.0f cycles/chunk  %5.1f%%" % (nm, pr[i] / n, 100.0 * pr[i] / tot))
print("  P1, warp 0 of each CTA: mbarrier wait %.0f cycles per slab, load + convert + refill %.0f cycles per slab (%.1f slabs per chunk)" % (
    pr[8] / max(pr[10], 1), pr[9] / max(pr[10], 1), pr[10] / n))

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
names = ["P0 fetch+surface", "P1 stage+convert", "P1x x-halo", "P2 faces+scan", "P3a list+indices", "P3b vertices+tail"]
tot = sum(pr[:6])
print("kernel ms %.3f  chunks %d  total cycles/chunk (thread0 view) %.0f" % (batch.elapsed_ms(), n, tot / n))
for i, nm in enumerate(names):
    print("  %-20s %8.0f cycles/chunk  %5.1f%%" % (nm, pr[i] / n, 100.0 * pr[i] / tot))
print("  producer wait(empty) %8.0f cycles/chunk ; class warp1 wait(full) %8.0f ; light warp wait(full) %8.0f" % (pr[8] / n, pr[9] / n, pr[10] / n))
print("  first-slab latency %6.0f cycles ; class warp0 per-slab processing %6.0f cycles (%d slabs/chunk) ; light warp per-slab %6.0f ; producer done issuing at %6.0f" % (
    pr[11] / n, pr[12] / max(pr[13], 1), pr[13] / n, pr[14] / max(pr[13], 1), pr[15] / n))
print("  CTA 0, last chunk: per slab  issue / ready / done  (cycles since P1 start)")
print("   " + " ".join("%d:%d/%d/%d" % (i, pr[64 + i], pr[128 + i], pr[192 + i]) for i in range(34)))

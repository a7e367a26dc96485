"""One surface + light fill of the bench world (for the ncu launch list of the pre-processing kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld
kind = sys.argv[1] if len(sys.argv) > 1 else "forest"
host = HostWorld(36, 36, origin=(-18, -18)).generate(kind, 1337, **(dict(radius=500, falloff=436) if kind == "islands" else {}))
ctx = mesher.Context(0)
world = mesher.World(ctx, 36, 36).upload(host)
for _ in range(2):
    world.compute_surface().light()
ctx.synchronize()
print(world.light_stats())

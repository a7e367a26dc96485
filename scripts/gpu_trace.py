import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld
host = HostWorld(36, 36, origin=(-18, -18)).build("forest", 1337)
coords = host.interior_chunks(2)
pin = [torch.from_numpy(a).pin_memory() for a in (host.blocks, host.sun, host.light, host.surface)]
class P:
    size_x, size_z = 36, 36
    blocks, sun, light, surface = [t.numpy() for t in pin]
if os.environ.get("NOLIGHT"):          # gcmesh_mesh_host(blocks, NULL, NULL, surface): the light is filled on the device
    P.sun = None; P.light = None
ctx = mesher.Context(0)
world = mesher.World(ctx, 36, 36)
if os.environ.get("HINT"):
    world.set_option(mesher.WORLD_SURFACE_BOUNDS_BLOCKS, 1).set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 1)
batch = mesher.Batch(ctx, len(coords), 4_000_000)
world.upload(host); ctx.synchronize(); batch.submit(world, coords).wait(); desc, q = batch.results()
hv = torch.empty((q * 4, 12), dtype=torch.uint8).pin_memory().numpy(); hi = torch.empty(q * 6, dtype=torch.uint16).pin_memory().numpy()
th = int(sys.argv[1]) if len(sys.argv) > 1 else 0
hint = None
if os.environ.get("CALLER_HINT"):
    hint = (P.blocks.any(axis=1) * 1 + P.sun.any(axis=1) * 2 + P.light.any(axis=1) * 4).astype(np.uint8)
for i in range(4):
    t0 = time.perf_counter(); batch.mesh_host(world, P, coords, hv, hi, threads=th, nonzero_hint=hint); print("call %.2f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)

import os, sys, time, subprocess
if len(sys.argv) == 1:
    for scan, pf in (("0", "0"), ("1", "0"), ("1", "1024"), ("1", "2048"), ("1", "4096"), ("1", "8192")):
        env = dict(os.environ, GCMESH_SCAN=scan, GCMESH_SCAN_PREFETCH=pf)
        print("SCAN=%s PREFETCH=%s" % (scan, pf), flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env)
    sys.exit(0)
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
ctx = mesher.Context(0)
world = mesher.World(ctx, 36, 36)
batch = mesher.Batch(ctx, len(coords), 4_000_000)
world.upload_sparse(P); ctx.synchronize(); batch.submit(world, coords).wait(); desc, q = batch.results()
hv = torch.empty((q * 4, 12), dtype=torch.uint8).pin_memory().numpy(); hi = torch.empty(q * 6, dtype=torch.uint16).pin_memory().numpy()
def t(f, n=8):
    f(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f(); ctx.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for th in (1, 8, 16):
    print("  upload_sparse threads %2d: %.2f ms" % (th, t(lambda: world.upload_sparse(P, threads=th))))
print("  mesh_host 16 threads: %.2f ms" % t(lambda: batch.mesh_host(world, P, coords, hv, hi, threads=16)))

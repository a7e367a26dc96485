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
ctx = mesher.Context(0)
world = mesher.World(ctx, 36, 36)
batch = mesher.Batch(ctx, len(coords), 4_000_000)
world.upload_sparse(P); ctx.synchronize(); batch.submit(world, coords).wait(); desc, q = batch.results()
hv = torch.empty((q * 4, 12), dtype=torch.uint8).pin_memory().numpy(); hi = torch.empty(q * 6, dtype=torch.uint16).pin_memory().numpy()
hint = (P.blocks.any(axis=1) * 1 + P.sun.any(axis=1) * 2 + P.light.any(axis=1) * 4).astype(np.uint8)
def t(f, n=5):
    f(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f(); ctx.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("upload_sparse (scan, %d threads) %.2f ms" % (os.cpu_count(), t(lambda: world.upload_sparse(P))))
for th in (1, 4, 8, 16):
    print("  scan threads %2d: %.2f ms" % (th, t(lambda: world.upload_sparse(P, threads=th))))
print("upload_sparse (hint)           %.2f ms" % t(lambda: world.upload_sparse(P, nonzero_hint=hint)))
print("upload dense                   %.2f ms" % t(lambda: world.upload(P)))
print("submit+wait                    %.2f ms" % t(lambda: batch.submit(world, coords).wait()))
print("download                       %.2f ms" % t(lambda: batch.download(hv, hi)))

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
coords_np = np.asarray(coords, dtype=np.int32)
def mh(hint_, threads=0):
    batch.mesh_host(world, P, coords, hv, hi, nonzero_hint=hint_, threads=threads)
print("mesh_host (scan)               %.2f ms" % t(lambda: mh(None)))
for th in (4, 8, 16, 32):
    print("  mesh_host scan threads %2d: %.2f ms" % (th, t(lambda: mh(None, th))))
print("mesh_host (hint)               %.2f ms" % t(lambda: mh(hint)))
# raw PCIe rates
a = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def cp(dst, src):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
for name, f in (("H2D", lambda: cp(d, a)), ("D2H", lambda: cp(a, d))):
    f(); t0 = time.perf_counter(); f(); f(); f(); dt = (time.perf_counter() - t0) / 3
    print("%s 256 MiB: %.2f ms = %.1f GB/s" % (name, dt * 1e3, (256 << 20) / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
a2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def both():
    with torch.cuda.stream(s1): d.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): a2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
both(); t0 = time.perf_counter(); both(); both(); both(); dt = (time.perf_counter() - t0) / 3
print("H2D+D2H concurrent 256 MiB each: %.2f ms = %.1f GB/s per direction" % (dt * 1e3, (256 << 20) / dt / 1e9))
# host zero scan alone: numpy view of how fast 16 threads can read the arrays
import ctypes, threading
print("host cores", os.cpu_count())

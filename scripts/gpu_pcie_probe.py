"""development: what the PCIe link gives on this box — copy-engine H2D / D2H (alone and together) against the library's pull kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
MB = 1 << 20
n = 256 * MB
hp = torch.empty(n, dtype=torch.uint8).pin_memory(); hp.fill_(1)
hq = torch.empty(n, dtype=torch.uint8).pin_memory()
d0 = torch.empty(n, dtype=torch.uint8, device="cuda"); d1 = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(f, k=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k
def h2d():
    with torch.cuda.stream(s1): d0.copy_(hp, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hq.copy_(d1, non_blocking=True)
def both(): h2d(); d2h()
print("copy engine H2D 256 MB: %.1f GB/s" % (n / timed(h2d) / 1e9))
print("copy engine D2H 256 MB: %.1f GB/s" % (n / timed(d2h) / 1e9))
t = timed(both); print("both at once: H2D %.1f + D2H %.1f GB/s" % (n / t / 1e9, n / t / 1e9))
def chunks(sz):
    def f():
        with torch.cuda.stream(s1):
            for o in range(0, n, sz): d0[o:o + sz].copy_(hp[o:o + sz], non_blocking=True)
    return f
for sz in (128 << 10, 1 * MB, 16 * MB):
    print("copy engine H2D in %d KB pieces: %.1f GB/s" % (sz >> 10, n / timed(chunks(sz), 2) / 1e9))
# the library's pull kernel (zero-copy reads by SMs): a world whose every array is non-zero, caller's hint -> no scan
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld
ctx = mesher.Context(0)
sx = sz_ = 16
pins = [torch.ones((sx * sz_ * 4, 65536), dtype=torch.uint16).pin_memory(), torch.ones((sx * sz_ * 4, 65536), dtype=torch.uint8).pin_memory(),
        torch.ones((sx * sz_ * 4, 65536), dtype=torch.uint8).pin_memory(), torch.ones((sx * sz_, 1024), dtype=torch.uint8).pin_memory()]
class P:
    size_x, size_z = sx, sz_
    blocks, sun, light, surface = [t.numpy() for t in pins]
world = mesher.World(ctx, sx, sz_)
hint = np.full(sx * sz_ * 4, 7, np.uint8)
total = sum(t.numel() * t.element_size() for t in pins[:3])
for ctas in (32, 64, 128, 296, 592):
    os.environ["GCMESH_PULL_CTAS"] = str(ctas)
def pull():
    world.upload_sparse(P, nonzero_hint=hint); ctx.synchronize()
print("pull kernel (%s CTAs) %d MB: %.1f GB/s" % (os.environ.get("GCMESH_PULL_CTAS_RUN", "64"), total >> 20, total / timed(pull, 3) / 1e9))
def pull_and_d2h():
    d2h(); world.upload_sparse(P, nonzero_hint=hint); ctx.synchronize()
t = timed(pull_and_d2h, 3); print("pull kernel with a D2H copy beside it: pull %.1f GB/s (D2H 256 MB in the same time: %.1f GB/s)" % (total / t / 1e9, n / t / 1e9))

"""Timing of the region-record decode / encode on a bench-sized world."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld
kind = sys.argv[1] if len(sys.argv) > 1 else "forest"
size = 36
host = HostWorld(size, size, origin=(-18, -18)).generate(kind, 1337, **(dict(radius=500, falloff=436) if kind == "islands" else {}))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = mesher.Context(0, stream.cuda_stream)
world = mesher.World(ctx, size, size).upload(host); ctx.synchronize()
every = np.asarray([(cx, cy, cz) for cz in range(size) for cx in range(size) for cy in range(4)], np.int32)
t0 = time.perf_counter(); loc, data = world.encode_rle(every); t1 = time.perf_counter()
print("%s: %d records, %.2f MB (blocks %.0f MB), runs %d; encode incl. read-back %.2f ms" % (kind, len(loc), data.nbytes / 1e6, host.blocks.nbytes / 1e6, (len(data) - 2 * len(loc)) // 2, (t1 - t0) * 1e3))
data = torch.from_numpy(data.copy()).pin_memory().numpy()
w2 = mesher.World(ctx, size, size)
def timed(f, n=5):
    f(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); ctx.synchronize()
    return e0.elapsed_time(e1) / n
print("upload_rle (H2D of the records + decode): %.3f ms" % timed(lambda: w2.upload_rle(loc, data)))
got = HostWorld(size, size); w2.download(got, blocks=True, sun=False, light=False, surface=False)
print("decoded == original:", np.array_equal(got.blocks, host.blocks), " refused:", w2.rle_refused())
w2.close(); world.close(); ctx.close()

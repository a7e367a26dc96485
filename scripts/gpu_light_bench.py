"""Timing of the device column pre-processing (surface + light fill) on a bench-sized world, next to the host fills."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld
kind = sys.argv[1] if len(sys.argv) > 1 else "forest"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 36
kw = dict(radius=500, falloff=436) if kind == "islands" else {}
t0 = time.perf_counter()
host = HostWorld(size, size, origin=(-(size // 2), -(size // 2))).generate(kind, 1337, **kw)
t1 = time.perf_counter(); host.compute_surface(); t2 = time.perf_counter(); host.compute_light(); t3 = time.perf_counter()
print("host: generate %.1f ms, surface %.1f ms, light fill %.1f ms (%d threads)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, os.cpu_count()))
t4 = time.perf_counter(); host.compute_light(threads=1); t5 = time.perf_counter()
print("host: light fill on one thread %.1f ms" % ((t5 - t4) * 1e3))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = mesher.Context(0, stream.cuda_stream)
blank = HostWorld(size, size); blank.blocks[:] = host.blocks
world = mesher.World(ctx, size, size).upload(blank); ctx.synchronize()
def timed(f, n=5):
    f(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record()
    for _ in range(n): f()
    e1.record(); ctx.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t) / n * 1e3
print("device surface: %.3f ms (events) %.3f ms (wall)" % timed(lambda: world.compute_surface()))
print("device light:   %.3f ms (events) %.3f ms (wall)" % timed(lambda: world.light()))
print(world.light_stats())
out = HostWorld(size, size); world.download(out)
print("equal to host fill: surface", np.array_equal(out.surface, host.surface), "sun", np.array_equal(out.sun, host.sun), "light", np.array_equal(out.light, host.light))

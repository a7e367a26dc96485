"""Soak of gcmesh_mesh_host without light arrays: many calls on several worlds, every chunk of every call against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import gcutil as util
from gamecraft_b200 import mesher
from oracle import binding as ob

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = mesher.Context(0)
bad_total = 0
for kind, size, seed, kw in (("forest", 12, 3, {}), ("islands", 9, 8, dict(radius=130, falloff=70)), ("mixed", 6, 29, dict(radius=100, falloff=60))):
    w = util.HostWorld(size, size, origin=(-(size // 2), -(size // 2))).generate(kind, seed, **kw)
    w.surface[:] = ob.compute_surface(w)
    w.sun[:], w.light[:] = ob.light_fill(w, w.surface)
    coords = w.interior_chunks(1)
    o = ob.OracleWorld(w)
    pins = [torch.from_numpy(a).pin_memory() for a in (w.blocks, w.sun, w.light, w.surface)]

    class P:
        size_x, size_z = size, size
        blocks, sun, light, surface = [t.numpy() for t in pins]

    class Q(P):
        sun = None
        light = None

    world = mesher.World(ctx, size, size).set_option(mesher.WORLD_SURFACE_BOUNDS_BLOCKS, 1).set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 1)
    cap = len(coords) * 4000
    batch = mesher.Batch(ctx, len(coords), cap)
    hv = torch.zeros((cap * 4, 12), dtype=torch.uint8).pin_memory().numpy()
    hi = torch.zeros(cap * 6, dtype=torch.uint16).pin_memory().numpy()
    t0 = time.time()
    bad = 0
    for k in range(calls):
        src = Q if k % 3 else P                      # both calls, interleaved
        hv[:64] = 0xAB
        batch.mesh_host(world, src, coords, hv, hi)
        desc, q = batch.results()
        b, _, checked = o.check_batch(coords, desc.copy(), hv, hi, threads=8)
        bad += b
        assert checked == q
    print("%s %dx%d: %d calls, %d chunks each, %d mismatching chunks, %.1f s" % (kind, size, size, calls, len(coords), bad, time.time() - t0), flush=True)
    bad_total += bad
    batch.close(); world.close()
# the light fill on its own, many times
host = util.HostWorld(10, 10).generate("mixed", 5, radius=150, falloff=90)
surface = ob.compute_surface(host); sun, light = ob.light_fill(host, surface)
world = mesher.World(ctx, 10, 10).upload(host)
out = util.HostWorld(10, 10)
bad = 0
for k in range(calls):
    world.compute_surface().light()
    world.download(out)
    bad += int(not (np.array_equal(out.sun, sun) and np.array_equal(out.light, light) and np.array_equal(out.surface, surface)))
print("light fill: %d calls, %d differing" % (calls, bad))
sys.exit(1 if (bad_total or bad) else 0)

"""A small pass through the round's new device code for compute-sanitizer (memcheck / racecheck): light fill (scan + persistent
relaxation), gcmesh_mesh_host without light arrays (group fills into shifted scratch arenas), the surface-bounded claim, the nine
generators, gcmesh_world_validate_blocks.  Results are compared with the oracle as well."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gcutil as util
from gamecraft_b200 import mesher
from oracle import binding as ob

ctx = mesher.Context(0)
size = 5
w = util.HostWorld(size, size, origin=(-2, -2)).generate("mixed", 29, radius=100, falloff=60)
w.surface[:] = ob.compute_surface(w)
w.sun[:], w.light[:] = ob.light_fill(w, w.surface)
world = mesher.World(ctx, size, size).upload(w)
world.compute_surface().light()
out = util.HostWorld(size, size); world.download(out)
assert np.array_equal(out.sun, w.sun) and np.array_equal(out.light, w.light) and np.array_equal(out.surface, w.surface)
print("light fill ok", world.light_stats())
assert world.validate_blocks()[0] == 0
coords = w.interior_chunks(1)
o = ob.OracleWorld(w)
batch = mesher.Batch(ctx, len(coords), len(coords) * 6000)
hv = np.zeros((len(coords) * 6000 * 4, 12), np.uint8); hi = np.zeros(len(coords) * 6000 * 6, np.uint16)

class Q:
    size_x, size_z = size, size
    blocks, surface = w.blocks, w.surface
    sun = None
    light = None

for opt in (0, 1):
    world.set_option(mesher.WORLD_SURFACE_BOUNDS_BLOCKS, opt).set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, opt)
    batch.mesh_host(world, Q, coords, hv, hi)
    desc, q = batch.results()
    assert o.check_batch(coords, desc.copy(), hv, hi)[0] == 0
print("mesh_host without light arrays ok")
f = util.HostWorld(4, 4).build("forest", 3)
fw = mesher.World(ctx, 4, 4).upload(f).set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 1)
fo = ob.OracleWorld(f)
for c, m in zip(f.interior_chunks(1).tolist(), mesher.rebuild_chunks(fw, f.interior_chunks(1))):
    util.assert_same_mesh(m, fo.build_chunk(*c), str(c))
print("surface-bounded submission ok")
for kind in ("snow", "desert", "volcanic", "dungeon"):
    g = mesher.World(ctx, 3, 3); g.generate(kind, 5, origin=(-1, -1)).light()
    got = util.HostWorld(3, 3); g.download(got, blocks=True)
    host = util.HostWorld(3, 3, origin=(-1, -1)).generate(kind, 5)
    assert np.array_equal(got.blocks, host.blocks), kind
    g.close()
print("generators ok")

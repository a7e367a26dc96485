"""Time of one edit under GC_WORLD_EDIT_REPLAY against the default relight, by kind of edit (forest window)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld

host = HostWorld(12, 12).build("forest", 1337)
ctx = mesher.Context(0)
for mode in (0, 1):
    world = mesher.World(ctx, 12, 12).upload(host).set_option(mesher.WORLD_EDIT_REPLAY, mode)
    mesher.rebuild_chunks(world, host.interior_chunks())
    wx, wz = 190, 200
    s = int(host.surface[(wz >> 5) * 12 + (wx >> 5)][(wz & 31) * 32 + (wx & 31)])
    cases = [("dig the top block", (wx, s - 1, wz, 0)), ("put it back", (wx, s - 1, wz, int(host.get_block(wx, s - 1, wz)))),
             ("lantern in the air", (wx, s + 3, wz, 14)), ("lantern removed", (wx, s + 3, wz, 0)),
             ("lantern in a pit", (wx + 5, s - 1, wz, 14)), ("pit lantern removed", (wx + 5, s - 1, wz, 0))]
    world.set_blocks([(60, 200, 60, 3)]); world.set_blocks([(60, 200, 60, 0)])        # scratch allocation, graph capture
    for name, e in cases:
        ctx.synchronize(); t0 = time.perf_counter()
        status, rebuild = world.set_blocks([e])
        print("%-8s %-22s %.3f ms, %d chunks flagged, status %d" % ("replay" if mode else "default", name, 1e3 * (time.perf_counter() - t0), len(rebuild), int(status[0])))
    world.close()

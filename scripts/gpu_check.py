"""GPU-side check used during development: parity of a few worlds vs the oracle + a quick timing.
usage: python scripts/gpu_check.py [kind size seed]..."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gamecraft_b200 import mesher
from gamecraft_b200.worldgen import HostWorld
from oracle import binding as ob


def check(kind, size, seed, time_it=True):
    kw = dict(radius=100, falloff=60) if kind in ("islands", "mixed") else {}
    host = HostWorld(size, size, origin=(-(size // 2), -(size // 2)) if kw else (0, 0)).build(kind, seed, **kw)
    ctx = mesher.Context(0)
    world = mesher.World(ctx, size, size).upload(host)
    coords = host.interior_chunks(1)
    batch = mesher.Batch(ctx, len(coords), len(coords) * 10000)
    t0 = time.time()
    batch.submit(world, coords).wait()
    meshes = batch.meshes()
    desc, q = batch.results()
    oracle = ob.OracleWorld(host)
    bad = 0
    for (cx, cy, cz), got, d in zip(coords, meshes, desc):
        want = oracle.build_chunk(cx, cy, cz)
        same = got.vert_count == want.vert_count and np.array_equal(got.vertices, want.vertices) and all(
            np.array_equal(a, b) for a, b in zip(got.indices, want.indices))
        if want.overflow:
            same = bool(d["flags"] & 1)
        if not same:
            bad += 1
            if bad <= 3:
                print("  MISMATCH", (cx, cy, cz), got.vert_count, want.vert_count, [len(i) for i in got.indices], [len(i) for i in want.indices], d["flags"])
                if got.vert_count == want.vert_count and got.vert_count:
                    rows = np.nonzero((got.vertices != want.vertices).any(1))[0]
                    print("   vertex rows differing:", len(rows), rows[:6])
                    for r in rows[:3]:
                        print("    got", got.vertices[r], "want", want.vertices[r])
    ms = []
    if time_it:
        for _ in range(5):
            batch.submit(world, coords).wait()
            ms.append(batch.elapsed_ms())
    print("%-8s size %2d seed %3d: chunks %5d quads %8d bad %d  kernel ms %s" % (kind, size, seed, len(coords), q, bad, ["%.3f" % m for m in ms]), flush=True)
    batch.close(); world.close(); ctx.close()
    return bad


if __name__ == "__main__":
    args = sys.argv[1:]
    jobs = [(args[i], int(args[i + 1]), int(args[i + 2])) for i in range(0, len(args), 3)] or [
        ("flat", 5, 1), ("mixed", 7, 7), ("forest", 7, 3), ("islands", 7, 5), ("checker", 5, 9), ("noise", 5, 2)]
    print(mesher.lib().gcmesh_version().decode(), "TMA" if os.environ.get("GCMESH_NO_TMA") != "1" else "NO_TMA")
    total_bad = sum(check(*j) for j in jobs)
    print("TOTAL BAD", total_bad)
    sys.exit(1 if total_bad else 0)

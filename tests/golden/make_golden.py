"""Mint golden vectors from the GOLD oracle: the reference's own mesher compiled verbatim (oracle/_ref).

    python tests/golden/make_golden.py          (only where /root/reference exists)

For each seeded world: FNV-1a-64 digests of the generated inputs (so a host whose float noise differs is detected,
not misreported as a mesher bug) and, per interior chunk, vertCount, the five index counts and digests of the
vertex bytes and of each index stream, all taken from BuildChunkAsync of the verbatim reference build.
Also writes two full-output fixtures (.npz) for byte-level debugging.  Committed outputs: golden_meshes.json,
golden_flat_chunk.npz, golden_mixed_chunk.npz.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gamecraft_b200.worldgen import HostWorld  # noqa: E402
from oracle import binding as ob  # noqa: E402

WORLDS = [
    # key, kind, size, seed, kwargs
    ("flat5", "flat", 5, 1, {}),
    ("mixed7", "mixed", 7, 7, {"radius": 100, "falloff": 60}),
    ("forest7", "forest", 7, 3, {}),
    ("islands7", "islands", 7, 5, {"radius": 100, "falloff": 60}),
    ("checker5", "checker", 5, 9, {}),
    ("noise5", "noise", 5, 2, {}),
]


def make_world(kind, size, seed, kw):
    origin = (-(size // 2), -(size // 2)) if kw else (0, 0)
    return HostWorld(size, size, origin=origin).build(kind, seed, threads=1, **kw)


def input_digest(w):
    return [ob.fnv1a64(w.blocks), ob.fnv1a64(w.sun), ob.fnv1a64(w.light), ob.fnv1a64(w.surface)]


def main():
    ob.build()
    assert ob.have_ref(), "needs oracle/_ref/libgcref.so (the reference tree must be present to build it)"
    out = {}
    for key, kind, size, seed, kw in WORLDS:
        w = make_world(kind, size, seed, kw)
        ref = ob.RefWorld(w)
        chunks = {}
        for cx, cy, cz in w.interior_chunks(1):
            m = ref.build_chunk(cx, cy, cz)
            chunks["%d,%d,%d" % (cx, cy, cz)] = [m.vert_count, [len(i) for i in m.indices], ob.fnv1a64(m.vertices) if m.vert_count else "",
                                                 [ob.fnv1a64(i) if len(i) else "" for i in m.indices]]
        out[key] = {"kind": kind, "size": size, "seed": seed, "kwargs": kw, "inputs": input_digest(w), "chunks": chunks}
        print(key, len(chunks), "chunks", sum(c[0] for c in chunks.values()) // 4, "quads")
        if key in ("flat5", "mixed7"):
            cx, cy, cz = (2, 0, 2) if key == "flat5" else (3, 0, 3)
            m = ref.build_chunk(cx, cy, cz)
            np.savez_compressed(os.path.join(HERE, "golden_%s_chunk.npz" % key[:-1]), coord=np.array([cx, cy, cz]), vertices=m.vertices,
                                **{"indices%d" % t: m.indices[t] for t in range(5)})
    with open(os.path.join(HERE, "golden_meshes.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()

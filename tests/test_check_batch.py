"""The exhaustive batch checker (oracle.check_batch: every chunk rebuilt on the CPU and compared byte for byte with a mesher
batch's descriptors + arenas) sees what it must see — exercised here on batches assembled from the kernel's own per-lane
code run on the host (tests/emu), without a GPU."""
import numpy as np
import pytest

import gcutil as util
from oracle import binding as ob

DESC = np.dtype([("quad_base", "<u4"), ("vert_count", "<u4"), ("index_count", "<u4", (5,)), ("index_offset", "<u4", (5,)),
                 ("quads", "<u4"), ("flags", "<u4"), ("reserved", "<u4", (2,))])


def assemble(w, coords, build):
    """A batch in the library's layout: segments in a shuffled order (the kernel reserves them in completion order)."""
    meshes = [build(*c) for c in coords]
    order = np.random.default_rng(1).permutation(len(coords))
    desc = np.zeros(len(coords), DESC)
    total = sum(m.vert_count // 4 for m in meshes)
    verts, idx = np.zeros((total * 4 + 4, 12), np.uint8), np.zeros(total * 6 + 6, np.uint16)
    base = 0
    for i in order:
        m, d = meshes[i], desc[i]
        q = m.vert_count // 4
        d["quads"] = getattr(m, "quads_total", q)
        if m.overflow:
            d["flags"] = 1
            continue
        d["quad_base"], d["vert_count"] = base, m.vert_count
        verts[base * 4: base * 4 + m.vert_count] = m.vertices
        off = 0
        for t in range(5):
            n = len(m.indices[t])
            d["index_offset"][t], d["index_count"][t] = off, n
            idx[base * 6 + off: base * 6 + off + n] = m.indices[t]
            off += n
        base += q
    return desc, verts, idx


@pytest.mark.parametrize("kind,seed", [("mixed", 4), ("noise", 2)])
def test_checker_accepts_the_emulated_kernel_and_sees_every_kind_of_damage(kind, seed):
    kw = dict(radius=100, falloff=60) if kind == "mixed" else {}
    w = util.HostWorld(4, 4, origin=(-2, -2)).build(kind, seed, **kw)
    coords = w.interior_chunks(1)
    desc, verts, idx = assemble(w, coords, lambda cx, cy, cz: util.emu_build_chunk(w, cx, cy, cz))
    o = ob.OracleWorld(w)
    bad, result, quads = o.check_batch(coords, desc, verts, idx, threads=3)
    assert bad == 0 and not result.any() and quads == int(desc["quads"].sum()) and quads > 0
    if ob.have_ref():
        r = ob.RefWorld(w)
        assert r.check_batch(coords, desc, verts, idx, threads=2)[0] == 0
        r.close()
    victim = int(np.argmax(desc["quads"]))
    qb = int(desc["quad_base"][victim])
    v2 = verts.copy(); v2[qb * 4 + 1, 6] ^= 1                       # one light byte of one vertex
    bad, result, _ = o.check_batch(coords, desc, v2, idx)
    assert bad == 1 and result[victim] == 2
    i2 = idx.copy(); i2[qb * 6 + 3] += 1                            # one index
    bad, result, _ = o.check_batch(coords, desc, verts, i2)
    assert bad == 1 and result[victim] == 4
    d2 = desc.copy(); d2["vert_count"][victim] -= 4; d2["quads"][victim] -= 1
    bad, result, _ = o.check_batch(coords, d2, verts, idx)
    assert bad == 1 and result[victim] & 1
    d2 = desc.copy(); d2["flags"][victim] = 2                      # an unexpected flag
    assert o.check_batch(coords, d2, verts, idx)[0] == 1
    d2 = desc.copy(); d2["quad_base"][victim] = len(idx) // 6       # a segment outside the arena is reported, not read
    bad, result, _ = o.check_batch(coords, d2, verts, idx)
    assert bad == 1 and result[victim] == 8


def test_checker_expects_the_overflow_flag_above_the_cap():
    w = util.cap_edge_world(pairs=2, singles=1664)                  # 10 004 quads in chunk (1, 1, 1)
    coords = np.array([(1, 1, 1), (1, 0, 1)], np.int32)
    o = ob.OracleWorld(w)
    desc = np.zeros(2, DESC)
    desc["quads"][0], desc["flags"][0] = 10004, 1
    verts, idx = np.zeros((4, 12), np.uint8), np.zeros(6, np.uint16)
    assert o.check_batch(coords, desc, verts, idx)[0] == 0
    desc["flags"][0] = 0
    assert o.check_batch(coords, desc, verts, idx)[0] == 1

"""CPU model of the device relight after a block edit (gcedit_kernel.cu): reset the full-height box of radius 16 around the
edited column to the state a from-scratch fill starts from, relax it for 14 sweeps against the unchanged cells around it,
and the window equals the oracle's from-scratch fill of the edited blocks.  Checks the algorithm (how far light can change,
the seeding rule, the fixed boundary) without a GPU; the kernels themselves are checked in tests/test_gpu_edit.py."""
import numpy as np
import pytest

from oracle import binding as ob
import gcutil as util
from test_edit_oracle import random_edit

RADIUS, SWEEPS = 16, 14


def to_window(w, arr):
    a = arr.reshape(w.size_z, w.size_x, 4, 32, 64, 32)
    return np.ascontiguousarray(a.transpose(0, 3, 2, 4, 1, 5).reshape(w.size_z * 32, 256, w.size_x * 32))     # [Z][Y][X]


def shifted(a, axis, d, fill):
    """a displaced so that out[p] = a[p + d along axis] (cells from outside read `fill`)."""
    out = np.full_like(a, fill)
    src = [slice(None)] * 3; dst = [slice(None)] * 3
    if d > 0:
        src[axis] = slice(d, None); dst[axis] = slice(0, -d)
    else:
        src[axis] = slice(0, d); dst[axis] = slice(-d, None)
    out[tuple(dst)] = a[tuple(src)]
    return out


def relight_box(blocks, surface, sun, light, wx, wz, step, emitted, size):
    """blocks / sun / light: [Z][Y][X] window arrays (sun, light edited in place); surface: [Z][X]."""
    Z, Y, X = blocks.shape
    z0, z1 = max(wz - RADIUS, 0), min(wz + RADIUS + 1, Z)
    x0, x1 = max(wx - RADIUS, 0), min(wx + RADIUS + 1, X)
    st = step[blocks & 255].astype(np.int32)                                   # 255 = opaque
    ys = np.arange(Y).reshape(1, Y, 1)
    inner = np.zeros((Z, X), bool)
    inner[32:Z - 32, 32:X - 32] = True                                         # pre-processed (non-edge) columns
    sky = (ys >= surface[:, None, :]) & (st != 255)                            # cells that read 15 ...
    sky_src = sky & inner[:, None, :]                                          # ... and were ever queued (SetLightNodes)
    # ComputeMaxY (Lighting.cpp:28-67): own surface and the four neighbours', capped at 255
    s = surface.astype(np.int32)
    my = s.copy()
    my[:, 1:] = np.maximum(my[:, 1:], s[:, :-1]); my[:, :-1] = np.maximum(my[:, :-1], s[:, 1:])
    my[1:, :] = np.maximum(my[1:, :], s[:-1, :]); my[:-1, :] = np.maximum(my[:-1, :], s[1:, :])
    my = np.minimum(my, 255)
    em = emitted[blocks & 255].astype(np.int32)
    seed = np.where((em > 1) & inner[:, None, :] & (ys <= my[:, None, :]), em, 0)
    box = np.zeros(blocks.shape, bool)
    box[z0:z1, :, x0:x1] = True
    sun[box] = 0
    light[box] = seed[box]
    sun_cand = box & (st != 255) & (ys < surface[:, None, :])
    light_cand = box & (st != 255)
    for arr, cand, is_sun in ((sun, sun_cand, True), (light, light_cand, False)):
        for _ in range(SWEEPS):
            val = np.where(sky_src, 15, arr).astype(np.int32) if is_sun else arr.astype(np.int32)
            if is_sun:
                val = np.where(sky & ~sky_src, 0, val)                         # a sky cell of an edge column offers nothing
            offer = np.where(st == 255, 0, val - st)                           # what each cell offers its neighbours
            best = arr.astype(np.int32)
            for axis in range(3):
                for d in (1, -1):
                    best = np.maximum(best, shifted(offer, axis, d, 0))
            upd = cand & (best > arr) & (best > 1)
            if not upd.any():
                break
            arr[upd] = best[upd]


@pytest.mark.parametrize("kind,seed", [("forest", 2), ("mixed", 3), ("islands", 8)])
def test_box_relight_equals_from_scratch_fill(kind, seed):
    size = 5
    kw = dict(radius=90, falloff=50) if kind in ("islands", "mixed") else {}
    w = util.HostWorld(size, size).generate(kind, seed, **kw)
    w.surface[:] = ob.compute_surface(w)
    w.sun[:], w.light[:] = ob.light_fill(w, w.surface)
    step, emitted = ob.default_light_rules()
    rng = np.random.default_rng(seed)
    palette = [0, 3, 14, 17, 8, 22]
    for e in range(6):
        wx, wy, wz, b = random_edit(rng, w, size)
        if e % 2:
            b = palette[e % len(palette)]
        w.set_block(wx, wy, wz, b)
        w.surface[:] = ob.compute_surface(w)
        blocks, sun, light = to_window(w, w.blocks), to_window(w, w.sun), to_window(w, w.light)
        surface = w.surface.reshape(size, size, 32, 32).transpose(0, 2, 1, 3).reshape(size * 32, size * 32)   # [Z][X]
        relight_box(blocks, surface, sun, light, wx, wz, np.asarray(step), np.asarray(emitted), size)
        want_sun, want_light = ob.light_fill(w, w.surface)
        assert np.array_equal(sun, to_window(w, want_sun)), "edit %d %r: sunlight" % (e, (wx, wy, wz, b))
        assert np.array_equal(light, to_window(w, want_light)), "edit %d %r: blockLight" % (e, (wx, wy, wz, b))
        w.sun[:], w.light[:] = want_sun, want_light

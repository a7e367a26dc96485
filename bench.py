#!/usr/bin/env python
"""bench.py — chunks meshed/s of the B200 chunk mesher (and the reference CPU mesher beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload forest4096|...]

One JSON line on stdout (rank 0).  A *step* is one pass of the hot path over one batch: every chunk of the
workload is meshed once (all vertex + index streams written).  N = 1 workload: BASELINE.json configs[1], 4096
seeded forest chunks (36x36-column window, inner 32x32 columns x 4 meshed).  N > 1: BASELINE.json configs[4] —
one 128 x 128-column forest (65 536 chunks) split into z-slabs over the ranks (strong scaling); every step each
rank meshes its slab and trades the boundary voxel planes with its slab neighbours over NVLink peer memory
inside the same kernel (gcmesh_submit_exchange); the weak-scaling run (4096 chunks per rank) is a second leg
of the same line.  Before the first exchange the margin rows a neighbour owns are overwritten with junk on the
device, and after the timed region EVERY chunk of the last submission is compared byte for byte with the CPU
oracle ("parity"): a broken exchange or kernel fails the run.

  value          whole-job chunks/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e            same metric through the host-facing C ABI: H2D of the window from pinned host memory + mesh + D2H of
                 descriptors and the vertex/index streams, all inside the timed region
  roofline       algorithmic bytes (262 424 + 60*quads per chunk) / mean kernel duration vs measured HBM peak
  cpu_baseline   the reference's own mesher (oracle/_ref, compiled verbatim) — or its plain-C port — on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

ALG_BYTES_PER_CHUNK = 262144 + 256 + 24   # three arrays once + surface/4 + counters
ALG_BYTES_PER_QUAD = 60                   # 4 x 12 B vertices + 6 x 2 B indices
FALLBACK_HBM_GBS = 6650.0                 # B200_PROFILING.md fallback

WORKLOADS = {
    # name: (kind, meshed columns x, meshed rows z per rank, seed, radius)
    "forest4096": ("forest", 32, 32, 1337, None),
    "forest1024": ("forest", 16, 16, 1337, None),
    "islands4096": ("islands", 32, 32, 4242, 512),
    "checker1024": ("checker", 16, 16, 99, None),
    "noise1024": ("noise", 16, 16, 77, None),
    # BASELINE config 5: one 128 x 128-column forest (65 536 chunks) split into z-slabs over the ranks (strong scaling);
    # rows per rank = 128 / N (the entry holds the total)
    "forest65536": ("forest", 128, 128, 1337, None),
    # BASELINE config 3 for several GPUs: one fixed-size island (radius 2048 blocks = 64 columns, fall-off over the last 64 blocks)
    # filling a 128 x 128-column window, every chunk re-meshed, split into z-slabs over the ranks (strong scaling)
    "islands65536": ("islands", 128, 128, 4242, 2048),
}
STRONG = {"forest65536", "islands65536"}


def rows_per_rank(workload: str, world_size: int) -> int:
    mz = WORKLOADS[workload][2]
    return mz // world_size if workload in STRONG else mz


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS), help="default: forest4096 at N = 1, forest65536 (BASELINE config 5) at N > 1")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", choices=["fused", "p2p", "nccl"], default="fused", help="N > 1: how slab neighbours trade boundary planes")
    ap.add_argument("--no-parity", action="store_true", help="skip the byte comparison of every chunk of the timed output with the oracle")
    ap.add_argument("--no-surface-hint", action="store_true", help="e2e: scan every block brick (GC_WORLD_SURFACE_BOUNDS_BLOCKS off)")
    ap.add_argument("--no-surface-bounded", action="store_true", help="skip the GC_WORLD_SURFACE_BOUNDS_MESH leg")
    ap.add_argument("--no-config5-leg", action="store_true", help="N = 1 default run: skip the forest65536 (BASELINE config 5) single-GPU leg")
    ap.add_argument("--no-weak-leg", action="store_true", help="N > 1 default run: skip the weak-scaling (forest4096 per rank) second leg")
    ap.add_argument("--no-preprocess", action="store_true", help="skip the surface + light-fill leg (SURVEY 8(f1))")
    ap.add_argument("--no-stress", action="store_true", help="skip the uncapped 196 608-quads-per-chunk leg (BASELINE config 4, stress variant)")
    return ap.parse_args()


def make_world(workload: str, rank: int, world_size: int):
    """The rank's window: meshed rows + 2 margin rows/columns each side (one halo ring for the mesher, one more so
    the halo ring's light is final, WorldRender.cpp:255-277).  Noise is sampled in world space, so overlapping rows
    of neighbouring ranks hold identical voxels."""
    from gamecraft_b200.worldgen import HostWorld, INT_MAX

    kind, mx, mz, seed, radius = WORKLOADS[workload]
    mz = rows_per_rank(workload, world_size)
    size_x, size_z = mx + 4, mz + 4
    origin = (-(size_x // 2), rank * mz - (world_size * mz + 4) // 2)
    t0 = time.time()
    host = HostWorld(size_x, size_z, origin=origin)
    host.build(kind, seed, radius=radius if radius else INT_MAX)
    coords = host.interior_chunks(2)
    return host, coords, time.time() - t0


def pin_to_local_cpus(local_rank: int, world_size: int):
    """One process per GPU: keep the rank's threads (zero-scan, stream feeding) and the pages they first touch (pinned host
    buffers) on the CPUs next to its GPU (the device's local_cpulist in sysfs).  A single-node host is left alone."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        local = cpus & allowed
        if local and len(local) < len(allowed):
            os.sched_setaffinity(0, local)
            return sorted(local)
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def dram_traffic_per_launch(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the mesh kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def preprocess_leg(ctx, host, pinned, args):
    """Surface map + sunlight / blockLight fill of the whole window on the device (gcmesh_world_compute_surface,
    gcmesh_world_light), checked against the host-filled arrays of the same world and timed next to the host fills."""
    import torch
    from gamecraft_b200 import mesher
    from gamecraft_b200.worldgen import HostWorld

    pw = mesher.World(ctx, host.size_x, host.size_z).upload(pinned)
    ctx.synchronize()

    def timed(f, k=5):
        f(); ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            f()
        e1.record()
        ctx.synchronize()
        return e0.elapsed_time(e1) / k

    surface_ms = timed(pw.compute_surface)
    light_ms = timed(pw.light)
    stats = pw.light_stats()
    got = HostWorld(host.size_x, host.size_z)
    pw.download(got)
    pw.close()
    same = bool(np.array_equal(got.surface, host.surface) and np.array_equal(got.sun, host.sun) and np.array_equal(got.light, host.light))
    lit = (host.size_x - 2) * (host.size_z - 2)
    n_slots = host.size_x * host.size_z * 4
    alg = n_slots * (131072 + 2 * 65536) + host.size_x * host.size_z * 1024     # block ids read once, both light arrays and the surface written once
    peak, peak_src = hbm_peak()
    ms = surface_ms + light_ms
    out = {"metric": "columns_preprocessed_per_s", "value": lit / (ms * 1e-3), "unit": "columns/s", "columns": lit,
           "surface_ms": surface_ms, "light_ms": light_ms, "gpu_launches": 1 + stats["launches"], "light_stats": stats,
           "identical_to_host_fill": same,
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes": alg, "peak_source": peak_src, "traffic": None}}
    if not args.no_cpu_baseline:
        scratch = HostWorld(host.size_x, host.size_z)
        scratch.blocks[:] = host.blocks
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        scratch.compute_surface(threads).compute_light(threads)
        sec = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": lit / sec, "unit": "columns/s", "cores": threads, "kind": "port",
                               "sample": "the same window: gcworld.c's ComputeSurface + SetLightNodes restatement (tested equal to the "
                                         "reference's), nine colour classes of columns in parallel"}
        try:
            from oracle import binding as ob
            if ob.have_ref():
                k = min(10, host.size_x, host.size_z)
                small = HostWorld(k, k)
                cols = [(cz + (host.size_z - k) // 2) * host.size_x + cx + (host.size_x - k) // 2 for cz in range(k) for cx in range(k)]
                for i, c in enumerate(cols):
                    small.blocks[4 * i:4 * i + 4] = host.blocks[4 * c:4 * c + 4]
                r = ob.RefWorld(small)
                t0 = time.perf_counter()
                r.compute_surface(); r.compute_light()
                sec = time.perf_counter() - t0
                r.close()
                out["cpu_reference"] = {"value": (k - 2) * (k - 2) / sec, "unit": "columns/s", "cores": 1, "kind": "reference",
                                        "sample": "a %dx%d-column cut of the same world through the reference's own ComputeSurface + SetLightNodes "
                                                  "(oracle/_ref), one thread, z-major column order" % (k, k)}
        except Exception as e:   # the verbatim build is optional
            out["cpu_reference"] = {"unavailable": str(e)}
    return out


def load_path_leg(ctx, world, batch, host, coords, total_quads, args):
    """The reference's chunk-load path on the device (SURVEY 8(f4) + 8(f1) + the mesher): saved region records (the RLE
    format of SaveGroup / LoadGroupFromDisk) in host memory -> block ids -> surface + light -> meshes back in host memory."""
    import torch
    from gamecraft_b200 import mesher

    every = np.asarray([(cx, cy, cz) for cz in range(host.size_z) for cx in range(host.size_x) for cy in range(4)], np.int32)
    loc, data = world.encode_rle(every)                       # untimed: the saved form of this window, written by the device encoder
    data = torch.from_numpy(data.copy()).pin_memory().numpy()
    lw = mesher.World(ctx, host.size_x, host.size_z)
    h_verts = torch.empty((total_quads * 4 + 4, 12), dtype=torch.uint8).pin_memory().numpy()
    h_idx = torch.empty(total_quads * 6 + 6, dtype=torch.uint16).pin_memory().numpy()

    def one():
        lw.upload_rle(loc, data)
        lw.compute_surface().light()
        batch.submit(lw, coords).wait()
        batch.download(h_verts, h_idx)

    for _ in range(2):
        one()
    ctx.synchronize()
    k = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(k):
        one()
    ctx.synchronize()
    sec = (time.perf_counter() - t0) / k
    desc, q = batch.results()
    ok = int(q) == int(total_quads) and lw.rle_refused() == 0
    # the same path with the meshes handed to device-resident (graphics) buffers instead of host memory (output
    # side): one vertex buffer and one element buffer per mesh type per chunk, carved out of two device allocations
    desc = desc.copy()
    d_vb = torch.empty(int(q) * 48 + 16, dtype=torch.uint8, device="cuda")
    d_ib = torch.empty(int(q) * 12 + 16, dtype=torch.uint8, device="cuda")
    targets = np.zeros(len(desc), mesher.DEVICE_MESH_DTYPE)
    targets["chunk"] = np.arange(len(desc))
    targets["vertices"] = d_vb.data_ptr() + desc["quad_base"].astype(np.uint64) * 48
    for t in range(5):
        targets["indices"][:, t] = d_ib.data_ptr() + (desc["quad_base"].astype(np.uint64) * 6 + desc["index_offset"][:, t].astype(np.uint64)) * 2

    def one_display():
        lw.upload_rle(loc, data)
        lw.compute_surface().light()
        batch.submit(lw, coords).wait()
        batch.fill_device(targets)

    for _ in range(2):
        one_display()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        one_display()
    ctx.synchronize()
    sec_display = (time.perf_counter() - t0) / k
    # (a chunk's sizes do not change between the steps, so the buffer layout carved above stays valid; where its streams lie
    # in the arenas does change — the call resolves that from the current descriptors)
    now, _ = batch.results()
    v_host, i_host = batch.download()
    v_host = np.asarray(v_host).reshape(-1)
    v_dev = d_vb.cpu().numpy()
    same_bytes = True
    for c in range(len(now)):
        nb = int(now["vert_count"][c]) * 12
        if nb and not np.array_equal(v_dev[int(desc["quad_base"][c]) * 48:int(desc["quad_base"][c]) * 48 + nb],
                                     v_host[int(now["quad_base"][c]) * 48:int(now["quad_base"][c]) * 48 + nb]):
            same_bytes = False
    lw.close()
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import binding as ob
            from gamecraft_b200.worldgen import HostWorld
            if ob.have_ref():
                # the same path through the reference's own code on one host thread, on a 10x10-column cut of the window:
                # LoadGroupFromDisk (decode) + ComputeSurface + SetLightNodes + BuildChunkAsync of the 8x8x4 inner chunks
                kk = min(10, host.size_x, host.size_z)
                small = HostWorld(kk, kk)
                cols = [(cz + (host.size_z - kk) // 2) * host.size_x + cx + (host.size_x - kk) // 2 for cz in range(kk) for cx in range(kk)]
                for i, c in enumerate(cols):
                    small.blocks[4 * i:4 * i + 4] = host.blocks[4 * c:4 * c + 4]
                r = ob.RefWorld(small)
                for cz in range(kk):
                    for cx in range(kk):
                        r.rle_save_group(cx, cz)              # untimed: puts the records into the reference's region store
                inner = small.interior_chunks(1)
                t0 = time.perf_counter()
                for cz in range(kk):
                    for cx in range(kk):
                        r.rle_load_group(cx, cz)
                r.compute_surface(); r.compute_light()
                sec_mesh, _ = r.bench(inner, 1)
                sec_cpu = time.perf_counter() - t0
                r.close()
                cpu = {"value": len(inner) / sec_cpu, "unit": "chunks/s", "cores": 1, "kind": "reference",
                       "sample": "a %dx%d-column cut: LoadGroupFromDisk + ComputeSurface + SetLightNodes for every column, BuildChunkAsync for the "
                                 "%d inner chunks, the reference's own code (oracle/_ref), one thread" % (kk, kk, len(inner))}
        except Exception as e:
            cpu = {"unavailable": str(e)}
    return {"metric": "chunks_meshed_per_s", "value": len(coords) / sec, "unit": "chunks/s", "ms_per_step": 1e3 * sec, "steps": k,
            "cpu_reference": cpu,
            "h2d_bytes_per_step": int(data.nbytes + loc.nbytes), "d2h_bytes_per_step": int(len(coords) * 64 + total_quads * 60),
            "records": int(len(loc)), "same_quads_as_host_prepared_world": bool(ok),
            "to_device_buffers": {"value": len(coords) / sec_display, "unit": "chunks/s", "ms_per_step": 1e3 * sec_display,
                                  "same_vertex_bytes_as_download": same_bytes, "d2h_bytes_per_step": int(len(coords) * 64),
                                  "call": "... + gcmesh_batch_fill_device (one vertex and up to five element buffers per chunk on the device) "
                                          "instead of gcmesh_batch_download"},
            "call": "gcmesh_world_upload_rle (every chunk of the window as its region record) + gcmesh_world_compute_surface + "
                    "gcmesh_world_light + gcmesh_submit + gcmesh_wait + gcmesh_batch_download"}


def stream_path_leg(ctx, batch, host, coords, total_quads, workload, gen_s):
    """Columns born in HBM (SURVEY 8(f2) + 8(f1) + the mesher): terrain generated on the device -> surface -> light -> meshes,
    nothing crossing PCIe but the per-chunk descriptors.  Same window as the host-prepared one of the main line."""
    from gamecraft_b200 import mesher
    from gamecraft_b200.worldgen import INT_MAX

    kind, mx, mz, seed, radius = WORKLOADS[workload]
    if kind not in ("forest", "islands", "flat"):
        return None
    radius = radius if radius else INT_MAX
    lw = mesher.World(ctx, host.size_x, host.size_z)

    def gen():
        lw.generate(kind, seed, origin=host.origin, radius=radius)

    def one():
        gen()                                                 # (block ids and the surface map)
        lw.light()
        batch.submit(lw, coords).wait()

    for _ in range(2):
        one()
    ctx.synchronize()
    k = 5
    t0 = time.perf_counter()
    for _ in range(k):
        one()
    ctx.synchronize()
    sec = (time.perf_counter() - t0) / k
    desc, q = batch.results()
    t0 = time.perf_counter()
    for _ in range(k):
        gen()
    ctx.synchronize()
    gen_sec = (time.perf_counter() - t0) / k
    from gamecraft_b200.worldgen import HostWorld
    got = HostWorld(host.size_x, host.size_z)
    lw.download(got, blocks=True, sun=False, light=False, surface=True)
    same = bool(np.array_equal(got.blocks, host.blocks)) and bool(np.array_equal(got.surface, host.surface))
    # every biome of the reference (Environment.cpp:83-143) on the same window: generate (block ids + surface map), ms per window
    biomes = {}
    try:
        for b in ("forest", "islands", "snow", "desert", "volcanic", "flat", "grid", "void", "dungeon"):
            rb = 2048 if b == "islands" else (INT_MAX if b != "grid" else 512)
            lw.generate(b, seed, origin=host.origin, radius=rb)
            ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                lw.generate(b, seed, origin=host.origin, radius=rb)
            ctx.synchronize()
            biomes[b] = 1e3 * (time.perf_counter() - t0) / 3
    except Exception as err:                                  # (a leg of the line, never the line itself)
        biomes["error"] = str(err)
    lw.close()
    n_cols = host.size_x * host.size_z
    return {"metric": "chunks_meshed_per_s", "value": len(coords) / sec, "unit": "chunks/s", "ms_per_step": 1e3 * sec, "steps": k,
            "generate_ms": 1e3 * gen_sec, "columns_generated_per_s": n_cols / gen_sec, "generate_ms_per_biome": biomes,
            "host_producer": {"seconds": gen_s, "columns_per_s": n_cols / gen_s if gen_s else None, "cores": os.cpu_count(),
                              "what": "gamecraft_b200/worldgen on the host cores: generate + surface + light of the same window (the run's own input preparation)"},
            "blocks_identical_to_host_producer": same, "same_quads_as_host_prepared_world": bool(int(q) == int(total_quads)),
            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(len(coords) * 64),
            "call": "gcmesh_world_generate (block ids + surface map) + gcmesh_world_light + gcmesh_submit + gcmesh_wait"}


def edit_path_leg(ctx, world, host, args):
    """The interactive caller on the device: SetBlock -> overflow pre-check -> relight -> rebuild list, then the
    remesh of the flagged chunks with their meshes read back — dig the top block of a column, then put it back."""
    from gamecraft_b200 import mesher

    rng = np.random.default_rng(7)
    small = mesher.Batch(ctx, 64, 64 * 4096)
    spots = []
    for _ in range(12):
        wx, wz = int(rng.integers(64, 32 * (host.size_x - 2))), int(rng.integers(64, 32 * (host.size_z - 2)))
        s = int(host.surface[(wz >> 5) * host.size_x + (wx >> 5)][(wz & 31) * 32 + (wx & 31)])
        if 1 <= s <= 254:
            spots.append((wx, s - 1, wz, int(host.get_block(wx, s - 1, wz))))
    t_edit, t_total, flagged, applied = [], [], [], 0
    for i, (wx, wy, wz, old) in enumerate(spots + spots):
        block = 0 if i < len(spots) else old
        ctx.synchronize()
        t0 = time.perf_counter()
        status, rebuild = world.set_blocks([(wx, wy, wz, block)])
        t1 = time.perf_counter()
        if len(rebuild):
            small.submit(world, rebuild[:64]).wait()
            small.download()
        t2 = time.perf_counter()
        if i >= 2:                                            # the first calls allocate the scratch
            t_edit.append(t1 - t0); t_total.append(t2 - t0); flagged.append(len(rebuild))
        applied += int(status[0] == mesher.EDIT_APPLIED)
    # the same edits with GC_WORLD_EDIT_REPLAY: the reference's RecomputeLight replayed node for node on one device thread
    t_replay, flagged_replay = [], []
    try:
        world.set_option(mesher.WORLD_EDIT_REPLAY, 1)
        for i, (wx, wy, wz, old) in enumerate(spots + spots):
            block = 0 if i < len(spots) else old
            ctx.synchronize()
            t0 = time.perf_counter()
            status, rebuild = world.set_blocks([(wx, wy, wz, block)])
            t1 = time.perf_counter()
            if i >= 2:
                t_replay.append(t1 - t0); flagged_replay.append(len(rebuild))
    except mesher.GcMeshError as err:
        print("edit replay leg: %s" % err, file=sys.stderr)
    finally:
        world.set_option(mesher.WORLD_EDIT_REPLAY, 0)
    small.close()
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import binding as ob
            from gamecraft_b200.worldgen import HostWorld
            if ob.have_ref():
                kk = min(10, host.size_x, host.size_z)
                ox, oz = (host.size_x - kk) // 2, (host.size_z - kk) // 2
                cut = HostWorld(kk, kk)
                for cz in range(kk):
                    for cx in range(kk):
                        c, i = (cz + oz) * host.size_x + cx + ox, cz * kk + cx
                        cut.blocks[4 * i:4 * i + 4] = host.blocks[4 * c:4 * c + 4]
                r = ob.RefWorld(cut)
                r.compute_surface(); r.compute_light()
                for cz in range(kk):
                    for cx in range(kk):
                        for cy in range(4):
                            r.set_chunk_built(cx, cy, cz, True, 0)
                r.download(cut)
                r.pending_updates(clear=True)
                secs = []
                for k in range(16):
                    wx, wz = int(rng.integers(64, 32 * (kk - 2))), int(rng.integers(64, 32 * (kk - 2)))
                    s = int(cut.surface[(wz >> 5) * kk + (wx >> 5)][(wz & 31) * 32 + (wx & 31)])
                    if not 1 <= s <= 254:
                        continue
                    t0 = time.perf_counter()
                    r.set_block(wx, s - 1, wz, 0)
                    pend = r.pending_updates(clear=True)
                    todo = [(cx, cy, cz) for cz in range(1, kk - 1) for cx in range(1, kk - 1) for cy in range(4) if pend[cz, cx, cy]]
                    if todo:
                        r.bench(np.asarray(todo, np.int32), 1)
                    secs.append(time.perf_counter() - t0)
                r.close()
                cpu = {"ms_per_edit_and_remesh": 1e3 * float(np.median(secs)), "cores": 1, "kind": "reference",
                       "sample": "%d edits on a %dx%d-column cut: the reference's own SetBlock (RecomputeLight, FlagChunkForUpdate) + "
                                 "BuildChunkAsync of the chunks it left pending (oracle/_ref), one thread" % (len(secs), kk, kk)}
        except Exception as e:
            cpu = {"unavailable": str(e)}
    return {"ms_per_edit": 1e3 * float(np.median(t_edit)), "ms_per_edit_and_remesh": 1e3 * float(np.median(t_total)),
            "edits": len(t_edit), "applied": applied, "chunks_flagged_per_edit": float(np.mean(flagged)), "cpu_reference": cpu,
            "replay_mode": ({"ms_per_edit": 1e3 * float(np.median(t_replay)), "chunks_flagged_per_edit": float(np.mean(flagged_replay)),
                             "what": "GC_WORLD_EDIT_REPLAY: byte-identical to the reference's RecomputeLight (its queues replayed on one device thread)"}
                            if t_replay else None),
            "call": "gcmesh_world_set_blocks (one edit: a CUDA graph replay of 19 small kernels) + gcmesh_submit of the flagged chunks + "
                    "gcmesh_wait + gcmesh_batch_download"}


def stress_leg(ctx, args):
    """BASELINE config 4, the stress variant: full 3-D checkerboard chunks — 196 608 quads each, twenty
    times the reference's cap — through an uncapped batch (GC_BATCH_UNCAPPED: 32-bit indices).  No parity claim against the
    reference (it asserts at its cap); checked against the oracle's traversal with the cap lifted on a sample of chunks."""
    import torch
    from gamecraft_b200 import mesher
    from gamecraft_b200.worldgen import HostWorld

    cols = 8                                                     # 8 x 8 meshed columns x 4 = 256 chunks, 50.3 M quads, 3.6 GB of output
    host = HostWorld(cols + 2, cols + 2).build("stress", 1)
    coords = host.interior_chunks(1)
    n = len(coords)
    world = mesher.World(ctx, host.size_x, host.size_z).upload(host)
    batch = mesher.Batch(ctx, n, n * 196608, uncapped=True)
    for _ in range(3):
        batch.submit(world, coords)
    batch.wait()
    k = max(1, min(args.steps, 10))
    ms = []
    for _ in range(k):
        batch.submit(world, coords).wait()
        ms.append(batch.elapsed_ms())
    desc, q = batch.results()
    kms = float(np.mean(ms))
    out_bytes = q * (48 + 24)
    alg = n * ALG_BYTES_PER_CHUNK + out_bytes
    peak, peak_src = hbm_peak()
    checked, bad = 0, 0
    if not args.no_parity:
        from oracle import binding as ob

        o = ob.OracleWorld(host)
        verts, idx = batch.download()
        for i in range(0, n, max(n // 8, 1)):
            d = desc[i]
            want = o.build_chunk_uncapped(*coords[i])
            qb, nq = int(d["quad_base"]), int(d["quads"])
            ok = nq == want.quads and np.array_equal(verts[qb * 4: qb * 4 + nq * 4], want.vertices)
            for t in range(5):
                ok = ok and np.array_equal(idx[qb * 6 + int(d["index_offset"][t]): qb * 6 + int(d["index_offset"][t]) + int(d["index_count"][t])], want.indices[t])
            checked += 1; bad += 0 if ok else 1
        del verts, idx
    flags = int((desc["flags"] != 0).sum())
    batch.close(); world.close()
    return {"workload": "stress196608", "chunks": n, "quads_per_chunk": q / n, "kernel_ms": kms, "chunks_per_s": n / (kms * 1e-3), "quads_per_s": q / (kms * 1e-3),
            "output_gbs": out_bytes / (kms * 1e-3) / 1e9, "index_bits": 32, "chunks_flagged": flags,
            "roofline": {"bound": "hbm", "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (kms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg, "note": "262 424 B in + 72 B per quad out (48 B of vertices, six uint32 indices)", "traffic": None},
            "checked_against_uncapped_oracle": {"chunks": checked, "mismatches": bad},
            "note": "no reference behaviour (the reference asserts at 10 000 quads per chunk, WorldRender.cpp:558): reported separately, no parity claim"}


def cpu_reference_run(host, coords, seconds: float, threads: int | None = None):
    """Times the reference CPU mesher on the host cores: oracle/_ref (the reference compiled verbatim) when that
    library travelled with the repo, else the plain-C port.  Returns the cpu_baseline object."""
    from oracle import binding as ob

    threads = threads or os.cpu_count() or 1
    if ob.have_ref() and host.size_x == host.size_z:
        impl, kind = ob.RefWorld(host), "reference"
    else:
        impl, kind = ob.OracleWorld(host), "port"
    n = len(coords)
    sec, quads = impl.bench(coords, threads)                    # warm-up pass (page-in, allocator)
    reps, total, total_q = 0, 0.0, 0
    while total < seconds and reps < 50:
        sec, quads = impl.bench(coords, threads)
        total += sec; total_q += quads; reps += 1
    one_n = max(1, min(n, int(n * 2.0 / max(total / reps * threads, 1e-6)) if reps else n))
    sec1, _ = impl.bench(coords[:one_n], 1)                      # single-thread figure on a bounded prefix
    return {
        "value": n * reps / total, "unit": "chunks/s", "cores": threads, "kind": kind,
        "sample": "%d x the full %d-chunk list, %d threads pulling chunks from one atomic counter (Async.cpp job queue shape)" % (reps, n, threads),
        "quads_per_s": total_q / total, "one_thread_chunks_per_s": one_n / sec1,
    }


def run_reference_arm(args):
    """--impl reference: the reference's own BuildChunkAsync (oracle/_ref, compiled verbatim) — or its plain-C port where that
    library did not travel — on all host threads, over the same chunk list as the N = 1 line of this arm.  Every step is
    timed on its own; `value` is the rate of the median step (a noisy neighbour on the host shows up as best << worst)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # this arm's configuration at --gpus N: forest4096 at N = 1, forest65536 (BASELINE config 5) at N > 1.  The CPU is given a
    # bounded sample of it: for forest65536 the 32 x 32 meshed columns at the centre of the same world (same generator, seed and
    # world-space noise: forest4096's window is that cut) — a full 65 536-chunk step would take the host most of a second
    named = args.workload or ("forest4096" if args.gpus <= 1 else "forest65536")
    workload = "forest4096" if named == "forest65536" else ("islands4096" if named == "islands65536" else named)
    host, coords, gen_s = make_world(workload, 0, 1)
    threads = os.cpu_count() or 1
    from oracle import binding as ob

    impl, kind = (ob.RefWorld(host), "reference") if (ob.have_ref() and host.size_x == host.size_z) else (ob.OracleWorld(host), "port")
    n = len(coords)
    for _ in range(max(args.warmup, 1)):
        impl.bench(coords, threads)
    secs, q = [], 0
    for _ in range(args.steps):
        sec, quads = impl.bench(coords, threads)
        secs.append(sec); q += quads
    med, best, worst, mean = float(np.median(secs)), float(np.min(secs)), float(np.max(secs)), float(np.mean(secs))
    value = n / med
    line = {
        "impl": "reference", "metric": "chunks_meshed_per_s", "value": value, "unit": "chunks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * med, "higher_is_better": True,
        "scaling": "strong" if named in STRONG else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": reference_arm_config(named, workload, args.gpus, host, coords),
        "quads_per_s": q / args.steps / med,
        "step_ms": {"median": 1e3 * med, "best": 1e3 * best, "worst": 1e3 * worst, "mean": 1e3 * mean,
                    "note": "value = chunks / median step; each step spawns its workers (one per host thread, pinned to the process's CPUs in turn)"},
        "best_step_chunks_per_s": n / best, "mean_chunks_per_s": n / mean,
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": kind,
                         "sample": ("every step = the full %d-chunk list on %d host threads" % (n, threads)) if named == workload else
                                   ("every step = %d chunks of %s (the 32 x 32 meshed columns at the centre of its world) on %d host threads" % (n, named, threads))},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_arm_config(named, sampled, n_gpus, host, coords):
    """`config` of the reference arm: this arm's configuration at the same --gpus (so the two lines name one workload), with the
    CPU's bounded sample spelled out when it is a cut of it."""
    if named == sampled:
        return workload_config(sampled, 1, host, coords)
    kind, mx, mz, seed, radius = WORKLOADS[named]
    rows = rows_per_rank(named, max(n_gpus, 1))
    return {"workload": named, "world": kind, "seed": seed, "chunks_per_gpu": mx * rows * 4, "window_columns_per_gpu": [mx + 4, rows + 4],
            "chunk_voxels": "32x64x32", "parallelism": "z-slab per GPU (the GPU arm); the CPU arm runs on the host's cores",
            "cpu_sample": {"workload": sampled, "chunks": int(len(coords)), "window_columns": [host.size_x, host.size_z]}}


def workload_config(workload, n_gpus, host, coords, exchange=None):
    kind, mx, mz, seed, radius = WORKLOADS[workload]
    how = {"fused": "boundary voxel planes stored into the neighbours' windows over NVLink peer memory by the mesh kernel itself "
                    "(its first work items; boundary-row chunks wait for the neighbour's arrival flag): one kernel per step, no collective",
           "p2p": "boundary voxel planes stored straight into the neighbours' windows over NVLink peer memory by an exchange kernel ahead of the mesh kernel (no collective)",
           "nccl": "NCCL send/recv of packed boundary voxel planes"}.get(exchange, "")
    return {
        "workload": workload, "world": kind, "seed": seed, "chunks_per_gpu": int(len(coords)),
        "window_columns_per_gpu": [host.size_x, host.size_z], "chunk_voxels": "32x64x32",
        "parallelism": ("z-slab per GPU, " + how) if n_gpus > 1 else "single GPU",
        "l2": "inputs (%.0f MB per step) exceed the 126 MB L2; no explicit flush" % (len(coords) * 262144 / 1e6),
    }


def scribble_margin_rows(world, host, rows, seed):
    """Overwrites whole rows of columns of the DEVICE window with junk (random block ids, light nibbles and surface heights).
    bench.py does this to the margin rows a slab neighbour owns before the first exchange: the rank's window was generated
    with correct margins (world-space noise), so without this a broken exchange would go unseen."""
    rng = np.random.default_rng(seed)
    jb = rng.integers(0, 24, 65536, dtype=np.uint16)
    js = rng.integers(0, 16, 65536, dtype=np.uint8)
    jl = rng.integers(0, 16, 65536, dtype=np.uint8)
    jf = rng.integers(0, 256, 1024, dtype=np.uint8)
    for cz in rows:
        for cx in range(host.size_x):
            for cy in range(4):
                world.upload_chunk(cx, cy, cz, np.roll(jb, cx * 4 + cy), np.roll(js, cx), np.roll(jl, cy))
            world.upload_surface(cx, cz, jf)
    world.ctx.synchronize()


_T0 = time.time()


def log(rank, msg):
    """Progress on stderr (GCBENCH_VERBOSE=1): which stage a rank is in, seconds since the process started."""
    if os.environ.get("GCBENCH_VERBOSE"):
        print("[bench %6.1fs rank %d] %s" % (time.time() - _T0, rank, msg), file=sys.stderr, flush=True)


def measure(args, workload, torch, dist, world_size, rank, local_rank, full_legs):
    """One workload: device-resident timing, parity of every chunk against the oracle, kernel-only timing, and (full_legs) the
    end-to-end and §8(f) legs.  Returns the JSON line (rank 0) or None."""
    from gamecraft_b200 import mesher

    log(rank, "building the %s window" % workload)
    host, coords, gen_s = make_world(workload, rank, world_size)
    n = len(coords)
    log(rank, "window %dx%d columns, %d chunks (%.1f s)" % (host.size_x, host.size_z, n, gen_s))

    # an explicit (non-default) stream: the library launches on it, torch's events and NCCL ops are ordered on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = mesher.Context(local_rank, stream.cuda_stream)
    world = mesher.World(ctx, host.size_x, host.size_z)

    # pinned host copies: the e2e leg uploads from these every step
    want_e2e = full_legs and not args.no_e2e
    if want_e2e:
        pin = [torch.from_numpy(a).pin_memory() for a in (host.blocks, host.sun, host.light, host.surface)]

        class Pinned:
            size_x, size_z = host.size_x, host.size_z
            blocks, sun, light, surface = [t.numpy() for t in pin]
    else:
        Pinned = host

    world.upload(Pinned)
    ctx.synchronize()
    log(rank, "window on the device")

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- slab halo exchange (N > 1): boundary voxel planes to / from the z-neighbours ----
    #   fused (default): gcmesh_submit_exchange — the mesh kernel itself stores the planes into the neighbours' windows
    #   p2p: gcmesh_world_halo_push (an exchange kernel over peer memory) ahead of gcmesh_submit
    #   nccl: pack -> NCCL send/recv -> unpack (also the fallback when the IPC handles cannot be opened)
    halo = None
    exchange_kind = None
    top_row, bottom_row = host.size_z - 3, 2              # last / first meshed row of columns
    if world_size > 1:
        exchange_kind = args.exchange
        if exchange_kind in ("fused", "p2p"):
            handles = [None] * world_size
            dist.all_gather_object(handles, world.halo_export())
            ok = 1
            try:
                if rank + 1 < world_size:
                    world.halo_connect_ipc(1, handles[rank + 1], bottom_row - 1)
                if rank > 0:
                    world.halo_connect_ipc(0, handles[rank - 1], top_row + 1)
            except mesher.GcMeshError as e:
                print("rank %d: peer-memory exchange unavailable (%s); using NCCL send/recv" % (rank, e), file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                exchange_kind = "nccl"
        if exchange_kind == "nccl":
            hb = world.halo_bytes()
            halo = {k: torch.empty(hb, dtype=torch.uint8, device="cuda") for k in ("send_up", "send_down", "recv_up", "recv_down")}
        # the margin rows that belong to a neighbour are junk on the device from here on: only the exchange can make the
        # boundary-row chunks come out right (parity below compares them with the oracle of the correct window)
        scribbled = []
        if rank > 0:
            scribbled += list(range(0, bottom_row))
        if rank + 1 < world_size:
            scribbled += list(range(top_row + 1, host.size_z))
        scribble_margin_rows(world, host, scribbled, 1000 + rank)
        barrier()

    def exchange():
        if exchange_kind == "p2p":
            world.halo_push(bottom_row, top_row)
            return 1
        ops = []
        if rank + 1 < world_size:
            world.pack_halo(top_row, 31, halo["send_up"].data_ptr())
            ops += [dist.P2POp(dist.isend, halo["send_up"], rank + 1), dist.P2POp(dist.irecv, halo["recv_up"], rank + 1)]
        if rank > 0:
            world.pack_halo(bottom_row, 0, halo["send_down"].data_ptr())
            ops += [dist.P2POp(dist.isend, halo["send_down"], rank - 1), dist.P2POp(dist.irecv, halo["recv_down"], rank - 1)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if rank + 1 < world_size:
            world.unpack_halo(top_row + 1, 0, halo["recv_up"].data_ptr())
        if rank > 0:
            world.unpack_halo(bottom_row - 1, 31, halo["recv_down"].data_ptr())
        return 2 * ((rank + 1 < world_size) + (rank > 0))

    def step(b):
        if world_size > 1 and exchange_kind == "fused":
            b.submit_exchange(world, coords, bottom_row, top_row)
            return 1
        launches = exchange() if world_size > 1 else 0
        b.submit(world, coords)
        return launches + 1

    log(rank, "exchange: %s" % exchange_kind)
    # size the arenas from a first pass (with its exchange: every rank steps the same number of times)
    probe = mesher.Batch(ctx, n, n * mesher.MAX_QUADS // 8 + 1)
    step(probe); probe.wait()
    desc, _ = probe.results()
    total_quads = int(desc["quads"].sum())
    again = torch.tensor([1 if (desc["flags"] != 0).any() else 0], dtype=torch.int32, device="cuda")
    if world_size > 1:
        dist.all_reduce(again, op=dist.ReduceOp.MAX)
    if int(again.item()):
        probe.close()
        probe = mesher.Batch(ctx, n, total_quads + 1)
        step(probe); probe.wait()
        desc, _ = probe.results()
    flags_bad = int((desc["flags"] != 0).sum())
    probe.close()

    # the first submission of a list has no quad history to order the work by (the reference's streaming caller meshes a chunk
    # once, WorldRender.cpp:177-180): time it on fresh batches
    cold_ms = []
    if world_size == 1:
        for _ in range(3):
            cold = mesher.Batch(ctx, n, total_quads + 1)
            cold.submit(world, coords).wait()
            cold_ms.append(cold.elapsed_ms())
            cold.close()

    log(rank, "%d quads; timing" % total_quads)
    batch = mesher.Batch(ctx, n, total_quads + 1)
    batch.set_timing(False)                              # nothing but the launch between back-to-back kernels

    # ---- device-resident timing ----
    for _ in range(max(args.warmup, 3)):
        step(batch)
    batch.wait()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    launches = 0
    for _ in range(args.steps):
        launches += step(batch)
    ev1.record()
    batch.wait()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()

    log(rank, "timed region done: %.3f ms per step on this rank" % (total_ms / args.steps))
    # ---- parity: every chunk of the last timed submission, byte for byte, against the oracle of this rank's (correct) window ----
    parity = None
    desc, tq = batch.results()
    assert int(desc["quads"].sum()) == total_quads
    if not args.no_parity:
        from oracle import binding as ob

        t0 = time.perf_counter()
        verts, idx = batch.download()
        cores = os.cpu_count() or 1
        bad, result, quads_checked = ob.OracleWorld(host).check_batch(coords, desc.copy(), verts, idx, threads=max(cores // world_size, 1))
        edge = (coords[:, 2] == bottom_row) | (coords[:, 2] == top_row)
        mine = [float(n), float(bad), float(quads_checked), float(int(edge.sum()) if world_size > 1 else 0), float(int((result[edge] != 0).sum()) if world_size > 1 else 0)]
        del verts, idx
        acc = torch.tensor(mine, dtype=torch.float64, device="cuda")
        per_rank = [None] * world_size
        if world_size > 1:
            dist.all_gather_object(per_rank, int(bad))
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        else:
            per_rank = [int(bad)]
        a = [int(v) for v in acc.tolist()]
        parity = {"chunks": a[0], "mismatches": a[1], "quads_compared": a[2], "mismatches_per_rank": per_rank,
                  "checker": "oracle/gc_mesh_cpu.c (plain-C restatement, pinned to the reference compiled verbatim): every chunk of the last timed "
                             "submission rebuilt on the host and compared byte for byte (counts, vertex stream, five index streams)",
                  "seconds": time.perf_counter() - t0}
        if world_size > 1:
            parity["boundary_row_chunks"] = a[3]
            parity["boundary_row_mismatches"] = a[4]
            parity["halo_rows_scribbled_before_first_exchange"] = True

    log(rank, "parity: %s" % (parity and {k: parity[k] for k in ("chunks", "mismatches")}))
    # mean duration of the mesh kernel alone (the library's own events bracket exactly the launch)
    batch.set_timing(True)
    kernel_ms = []
    for _ in range(min(args.steps, 10)):
        step(batch); batch.wait()
        kernel_ms.append(batch.elapsed_ms())
    batch.set_timing(False)

    # ---- the same submission with GC_WORLD_SURFACE_BOUNDS_MESH: the window's surface maps are ComputeSurface of its block ids (they
    # are: the producer computed them, as PreprocessGroup does before any chunk is meshed, WorldRender.cpp:240-262), so the kernel
    # takes bricks at or above their column's highest surface entry for AIR from the 1 KB map instead of reading 128 KB of ids.
    # Reported beside the headline, which stays the mesher as a function of the block ids alone. ----
    bounded = None
    if not args.no_surface_bounded:
        world.set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 1)
        for _ in range(3):
            step(batch)
        batch.wait()
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            step(batch)
        b1.record()
        batch.wait()
        barrier()
        bt = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device="cuda")
        bbad = None
        if not args.no_parity:
            from oracle import binding as ob

            bdesc, _ = batch.results()
            verts, idx = batch.download()
            bbad = ob.OracleWorld(host).check_batch(coords, bdesc.copy(), verts, idx, threads=max((os.cpu_count() or 1) // world_size, 1))[0]
            del verts, idx
        bb = torch.tensor([float(bbad or 0)], dtype=torch.float64, device="cuda")
        if world_size > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        world.set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 0)
        bounded = {"ms_per_step": float(bt.item()) / args.steps, "steps": args.steps,
                   "parity_mismatches": None if bbad is None else int(bb.item()),
                   "option": "GC_WORLD_SURFACE_BOUNDS_MESH: bricks at or above their column's highest surface entry are AIR by ComputeSurface's "
                             "definition (Lighting.cpp:5-17); the kernel answers them from the surface map and does not read their block ids"}

    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    km = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device="cuda")
    cq = torch.tensor([float(n), float(total_quads)], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
        dist.all_reduce(cq, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    all_chunks, all_quads = float(cq[0].item()), float(cq[1].item())
    ms_per_step = total_ms / args.steps
    value = all_chunks / (ms_per_step * 1e-3)

    # ---- end to end through the host-facing C ABI: host arrays in pinned memory -> meshes back in host memory ----
    # Every step: classify the host brick arrays zero / non-zero (multi-threaded scan inside the library), move the
    # non-zero ones over PCIe (GPU copy kernel pulling from the pinned arrays), mesh, copy descriptors + vertex + index
    # streams back.  "e2e_dense" is the same with a plain full upload (every array of every chunk, cudaMemcpyAsync).
    e2e = None
    e2e_dense = None
    log(rank, "device-resident legs done")
    if want_e2e:
        h_verts = torch.empty((total_quads * 4 + 4, 12), dtype=torch.uint8).pin_memory().numpy()
        h_idx = torch.empty(total_quads * 6 + 6, dtype=torch.uint16).pin_memory().numpy()
        dense_h2d = sum(a.nbytes for a in (Pinned.blocks, Pinned.sun, Pinned.light, Pinned.surface))
        d2h = n * 64 + total_quads * 60

        # the ranks of one box share its host cores: each takes its share of them for the zero-scan
        cores = os.cpu_count() or 2
        scan_threads = max((cores - 1) // world_size, 1) if world_size > 1 else 0

        def timed(one, ksteps=None):
            ksteps = ksteps or max(1, min(args.steps, 10))
            for _ in range(2):
                one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                one()
            barrier()
            sec = (time.perf_counter() - t0) / ksteps
            te = torch.tensor([sec], dtype=torch.float64, device="cuda")
            if world_size > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item()), ksteps

        def staged(upload):
            def one():
                upload()
                batch.submit(world, coords).wait()
                batch.download(h_verts, h_idx)
            return one

        def pipelined():
            # gcmesh_mesh_host: scan / PCIe pull / mesh / read-back overlapped strip by strip.  With slabs on several GPUs the
            # boundary planes arrive with the host arrays themselves (every rank's window holds its halo rows), so no exchange.
            batch.mesh_host(world, Pinned, coords, h_verts, h_idx, threads=scan_threads)

        # the windows' surface maps are ComputeSurface of their block ids (as ChunkGroup::surface always is): the scan may take the
        # bricks above them for AIR unread (GC_WORLD_SURFACE_BOUNDS_BLOCKS), and so may the kernel (GC_WORLD_SURFACE_BOUNDS_MESH)
        world.set_option(mesher.WORLD_SURFACE_BOUNDS_BLOCKS, 0 if args.no_surface_hint else 1)
        world.set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 0 if args.no_surface_hint else 1)
        hint_txt = ("every block brick is scanned" if args.no_surface_hint else
                    "GC_WORLD_SURFACE_BOUNDS_BLOCKS / _MESH: block bricks above their column's surface map are AIR by ComputeSurface's definition and are not read")
        threads_txt = "%d host threads%s" % (scan_threads or max(cores - 1, 1), " per rank" if world_size > 1 else "")

        def e2e_parity():
            # the meshes that came back over PCIe, every chunk, against the oracle
            from oracle import binding as ob

            d2, _ = batch.results()
            bad2, _, _ = ob.OracleWorld(host).check_batch(coords, d2.copy(), h_verts, h_idx, threads=max(cores // world_size, 1))
            acc = torch.tensor([float(bad2)], dtype=torch.float64, device="cuda")
            if world_size > 1:
                dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            return int(acc.item())

        # (1) the host hands in block ids, both light arrays and the surface maps — BuildChunkAsync's inputs
        sec, ksteps = timed(pipelined)
        h2d = world.upload_bytes()
        e2e_host_light = {"value": all_chunks / sec, "unit": "chunks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                          "ms_per_step": 1e3 * sec, "steps": ksteps, "gpu_launches_per_step": batch.launches() + world.upload_launches(),
                          "call": "gcmesh_mesh_host(blocks, sunlight, blockLight, surface): zero/non-zero scan of the host arrays on %s (one core is left to the "
                                  "thread that feeds the streams; %s), copy-engine / GPU pull of the non-zero brick arrays, meshing and read-back pipelined "
                                  "in strips of 2 rows on three streams" % (threads_txt, hint_txt)}
        if not args.no_parity:
            e2e_host_light["parity_mismatches"] = e2e_parity()

        # (2) the host hands in block ids and the surface maps only: the light is filled on the device, group of rows by group of
        # rows (PreprocessGroup's fill, WorldRender.cpp:240-262 — the light the host arrays of (1) hold), so neither light array is
        # scanned or crosses the link
        class BlocksOnly:
            size_x, size_z = Pinned.size_x, Pinned.size_z
            blocks, surface = Pinned.blocks, Pinned.surface
            sun = None
            light = None

        e2e_device_light = None
        try:
            sec, ksteps = timed(lambda: batch.mesh_host(world, BlocksOnly, coords, h_verts, h_idx, threads=scan_threads))
            e2e_device_light = {"value": all_chunks / sec, "unit": "chunks/s", "h2d_bytes_per_step": int(world.upload_bytes()), "d2h_bytes_per_step": int(d2h),
                                "ms_per_step": 1e3 * sec, "steps": ksteps, "gpu_launches_per_step": batch.launches() + world.upload_launches(),
                                "call": "gcmesh_mesh_host(blocks, NULL, NULL, surface): block-id bricks classified on %s (%s) and moved by the copy engines; "
                                        "sunlight and blockLight filled on the device per group of 6 rows (a scan kernel + one persistent relaxation kernel "
                                        "each); meshing and read-back pipelined behind it" % (threads_txt, hint_txt)}
            if not args.no_parity:
                e2e_device_light["parity_mismatches"] = e2e_parity()
        except mesher.GcMeshError as err:
            e2e_device_light = None
            print("rank %d: device-light e2e unavailable: %s" % (rank, err), file=sys.stderr)
        world.set_option(mesher.WORLD_SURFACE_BOUNDS_MESH, 0)
        # headline: the call with the fewest bytes on the link whose meshes equal the oracle's; the other one is kept beside it
        if e2e_device_light and e2e_device_light.get("parity_mismatches", 0) == 0 and e2e_device_light["value"] > e2e_host_light["value"]:
            e2e = dict(e2e_device_light); e2e["light_arrays_in"] = e2e_host_light
        else:
            e2e = dict(e2e_host_light); e2e["light_on_device"] = e2e_device_light
        if world_size == 1:
            sec, ksteps = timed(staged(lambda: world.upload_sparse(Pinned, threads=scan_threads)))
            e2e_staged = {"value": all_chunks / sec, "unit": "chunks/s", "ms_per_step": 1e3 * sec, "steps": ksteps,
                          "call": "gcmesh_world_upload_sparse + gcmesh_submit + gcmesh_wait + gcmesh_batch_download, one after the other"}
            sec, ksteps = timed(staged(lambda: world.upload(Pinned)), 3)
            e2e_dense = {"value": all_chunks / sec, "unit": "chunks/s", "h2d_bytes_per_step": int(dense_h2d), "d2h_bytes_per_step": int(d2h),
                         "ms_per_step": 1e3 * sec, "steps": ksteps, "call": "gcmesh_world_upload_all (every array of every chunk) + submit + wait + download"}
            e2e_dense["staged_sparse"] = e2e_staged

    # ---- column pre-processing on the device: surface map + light flood fill from the block ids ----
    preprocess = load_path = stream_path = edit_path = stress = None
    if full_legs and not args.no_stress and world_size == 1:
        stress = stress_leg(ctx, args)
    if full_legs and not args.no_preprocess and world_size == 1:
        preprocess = preprocess_leg(ctx, host, Pinned, args)
        load_path = load_path_leg(ctx, world, batch, host, coords, total_quads, args)
        stream_path = stream_path_leg(ctx, batch, host, coords, total_quads, workload, gen_s)
        edit_path = edit_path_leg(ctx, world, host, args)     # last: it edits the device world

    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        kms_isolated = float(km.item())
        # The timed region holds nothing but the K launches of the mesh kernel when a step is one launch (single GPU, or the
        # exchange fused into the kernel): the kernel's average duration over the region is then the region's time / K.  Launches
        # are programmatically dependent (a launch's set-up runs beside its predecessor's tail), so a launch bracketed by
        # events and waited for on its own — the "isolated" figure — lasts a few microseconds longer than its share of the region.
        one_launch_per_step = launches == args.steps
        kms = ms_per_step if one_launch_per_step else kms_isolated
        alg_bytes = n * ALG_BYTES_PER_CHUNK + total_quads * ALG_BYTES_PER_QUAD
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        traffic = dram_traffic_per_launch(workload) if world_size == 1 else None
        line = {
            "metric": "chunks_meshed_per_s", "value": value, "unit": "chunks/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if workload in STRONG else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(workload, args.gpus, host, coords, exchange_kind),
            "quads_per_s": all_quads / (ms_per_step * 1e-3), "quads_per_chunk": total_quads / n, "chunks_flagged": flags_bad,
            "parity": parity,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "gc_mesh_kernel", "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_isolated": kms_isolated,
                         "kernel_ms_note": ("kernel_ms: the timed region (CUDA events on the launching stream around %d back-to-back launches and nothing else) / %d; "
                                            "kernel_ms_isolated: mean over %d launches each bracketed by events and waited for%s" %
                                            (args.steps, args.steps, len(kernel_ms), "; rank 0's launch moves rank 0's bytes, the duration is the slowest rank's" if world_size > 1 else ""))
                                           if one_launch_per_step else
                                           ("mean over %d launches, CUDA events around the launch on the launching stream (a step holds other kernels too)" % len(kernel_ms)),
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         # light is only read where quads are, so fewer bytes move than the algorithmic figure credits;
                         # the physical rate (ncu DRAM bytes of one launch / this run's kernel time) is reported next to it
                         "physical_gbs": (traffic / (kms * 1e-3) / 1e9) if traffic else None,
                         "physical_frac": (traffic / (kms * 1e-3) / 1e9 / peak) if traffic else None},
            "gpu_launches": launches, "clocks": clocks, "kernel_info": ctx.kernel_info(), "worldgen_s": gen_s,
            # what a step holds beside the isolated kernel time: negative when back-to-back launches overlap (programmatic dependent launch)
            "between_kernels_us": 1e3 * (ms_per_step - kms_isolated),
        }
        if bounded:
            bounded["value"] = all_chunks / (bounded["ms_per_step"] * 1e-3)
            bounded["unit"] = "chunks/s"
            line["surface_bounded"] = bounded
        if cold_ms:
            line["cold_order"] = {"kernel_ms": float(np.median(cold_ms)), "chunks_per_s": n / (float(np.median(cold_ms)) * 1e-3),
                                  "note": "first submission of the list on a fresh batch: no quad history, work order = lower chunks first"}
        if world_size > 1:
            line["exchange"] = {"kind": exchange_kind, "collective": "none (peer-memory stores + flags)" if exchange_kind != "nccl" else "NCCL send/recv",
                                "process_group": "NCCL (barrier and all-reduce of the timings only)"}
        if e2e:
            line["e2e"] = e2e
            if e2e_dense:
                line["e2e_dense"] = e2e_dense
        if full_legs and not args.no_cpu_baseline and world_size == 1:
            line["cpu_baseline"] = cpu_reference_run(host, coords, args.cpu_seconds)
        for k, v in (("stress", stress), ("preprocess", preprocess), ("load_path", load_path), ("stream_path", stream_path), ("edit_path", edit_path)):
            if v:
                line[k] = v

    log(rank, "legs done")
    halo_timed_out = world.halo_status()[0] if exchange_kind in ("p2p", "fused") else False
    if world_size > 1:
        dist.barrier()                                    # nobody unmaps a window a neighbour may still write to
    batch.close(); world.close(); ctx.close()
    if halo_timed_out:
        raise SystemExit("rank %d: a halo exchange timed out waiting for a neighbour" % rank)
    if parity and parity["mismatches"]:
        raise SystemExit("rank %d: PARITY FAILURE — %d of %d chunks differ from the oracle (per rank: %s)" %
                         (rank, parity["mismatches"], parity["chunks"], parity["mismatches_per_rank"]))
    return line


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size != args.gpus and world_size > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world_size))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mesher has no CPU path")
    torch.cuda.set_device(local_rank)
    pin_to_local_cpus(local_rank, world_size)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # N = 1: BASELINE config 2 (4096 forest chunks).  N > 1: BASELINE config 5 (65 536 forest chunks, one world split into
    # z-slabs: strong scaling) with the weak-scaling run (4096 chunks per rank) as a second leg of the same line.
    workload = args.workload or ("forest4096" if world_size == 1 else "forest65536")
    line = measure(args, workload, torch, dist, world_size, rank, local_rank, full_legs=True)
    if world_size > 1 and not args.workload and not args.no_weak_leg:
        weak = measure(args, "forest4096", torch, dist, world_size, rank, local_rank, full_legs=False)
        if rank == 0:
            line["weak_scaling_leg"] = {k: weak[k] for k in ("value", "unit", "ms_per_step", "scaling", "config", "quads_per_s", "parity", "roofline", "gpu_launches", "between_kernels_us")}
    if world_size == 1 and not args.workload and not args.no_config5_leg:
        # the N = 1 point of the strong-scaling curve the N > 1 runs report: BASELINE config 5 on one GPU (17 GB resident)
        c5 = measure(args, "forest65536", torch, dist, world_size, rank, local_rank, full_legs=False)
        line["config5_single_gpu_leg"] = {k: c5[k] for k in ("value", "unit", "ms_per_step", "scaling", "config", "quads_per_s", "parity", "roofline", "gpu_launches",
                                                               "between_kernels_us", "surface_bounded") if k in c5}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

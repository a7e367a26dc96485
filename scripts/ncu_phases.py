"""Instruction / stall-sample share per kernel phase from an .ncu-rep (source-line attribution grouped by line ranges)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': h = r; ie = r.index('Instructions Executed'); isamp = r.index('# Samples'); continue
    if h and len(r) > ie and r[0].isdigit():
        agg[(cur, int(r[0]))] = (int(r[ie]) if r[ie].isdigit() else 0, int(r[isamp]) if r[isamp].isdigit() else 0)
import os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def first_line(path, pattern):
    for i, line in enumerate(open(os.path.join(ROOT, path)), 1):
        if re.search(pattern, line):
            return i
    raise SystemExit("marker not found: " + pattern)
K = "gamecraft_b200/csrc/gcmesh_kernel.cu"
C = "gamecraft_b200/csrc/mesh_core.cuh"
k_p0 = first_line(K, r"=+ P0:"); k_p1 = first_line(K, r"=+ P1: stage"); k_halo = first_line(K, r"---- y-halo rows")
k_p1x = first_line(K, r"=+ P1x:"); k_p2 = first_line(K, r"=+ P2:"); k_p3 = first_line(K, r"=+ P3: emission")
k_p3a = first_line(K, r"---- 3a:"); k_p3b = first_line(K, r"---- 3b:")
c_p1 = first_line(C, r"Phase 1: raw voxels"); c_p1x = first_line(C, r"GC_HD void p1_xhalo_cell"); c_light = first_line(C, r"GC_HD uint32_t light_pair_of")
c_p2 = first_line(C, r"Phase 2: face masks"); c_p3a = first_line(C, r"Phase 3a:"); c_p3b = first_line(C, r"GC_HD void quad_index_words")
def grp(f, l):
    if f == 'gcmesh_kernel.cu':
        if l < k_p0: return 'setup / helpers (mbarrier waits, barriers)'
        if l < k_p1: return 'P0'
        if l < k_halo: return 'P1 (kernel)'
        if l < k_p1x: return 'y-halo + z-halo slabs'
        if l < k_p2: return 'P1x'
        if l < k_p3: return 'P2 (kernel)'
        if l < k_p3a: return 'P3 setup (surface tile, batches)'
        if l < k_p3b: return 'P3a (kernel)'
        return 'P3b (kernel)'
    if f == 'mesh_core.cuh':
        if l < c_p1: return 'core helpers (byte_perm, transpose, planes store)'
        if l < c_p1x: return 'P1 (core: uniformity, class rows)'
        if l < c_light: return 'P1x (core)'
        if l < c_p2: return 'P3b light_pair_of'
        if l < c_p3a: return 'P2 / P3a row faces + counts'
        if l < c_p3b: return 'P3a p3a_row'
        return 'P3b (core)'
    return f
g = {}
for (f, l), v in agg.items():
    a = g.setdefault(grp(f, l), [0, 0]); a[0] += v[0]; a[1] += v[1]
tot = sum(v[0] for v in g.values()) or 1; ts = sum(v[1] for v in g.values()) or 1
print("%-52s %8s %9s" % ("phase", "inst %", "samples %"))
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0]):
    print("%-52s %8.1f %9.1f" % (k, 100 * v[0] / tot, 100 * v[1] / ts))

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
def grp(f, l):
    if f == 'gcmesh_kernel.cu':
        if l < 260: return 'setup / helpers (mbarrier waits, barriers)'
        if l < 268: return 'P0'
        if l < 367: return 'P1 (kernel)'
        if l < 416: return 'y-halo + z-halo slabs'
        if l < 471: return 'P1x'
        if l < 572: return 'P2 (kernel)'
        if l < 605: return 'P3 setup (surface tile, batches)'
        if l < 630: return 'P3a (kernel)'
        return 'P3b (kernel)'
    if f == 'mesh_core.cuh':
        if l < 148: return 'core helpers (byte_perm, transpose, planes store)'
        if l < 226: return 'P1 (core: uniformity, class rows)'
        if l < 247: return 'P1x (core)'
        if l < 255: return 'P3b light_pair_of'
        if l < 313: return 'P2 / P3a row faces + counts'
        if l < 367: return 'P3a p3a_row'
        return 'P3b (core)'
    return f
g = {}
for (f, l), v in agg.items():
    a = g.setdefault(grp(f, l), [0, 0]); a[0] += v[0]; a[1] += v[1]
tot = sum(v[0] for v in g.values()) or 1; ts = sum(v[1] for v in g.values()) or 1
print("%-52s %8s %9s" % ("phase", "inst %", "samples %"))
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0]):
    print("%-52s %8.1f %9.1f" % (k, 100 * v[0] / tot, 100 * v[1] / ts))

"""Summarise an .ncu-rep: key raw metrics + instruction / stall-sample share per source line.
usage: python scripts/ncu_summary.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.max']
for k in keep:
    if k in hdr:
        i = hdr.index(k); print("%-70s %-12s %s" % (k, units[i], r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': h = r; ie = r.index('Instructions Executed'); isamp = r.index('# Samples'); continue
    if h and len(r) > ie and r[0].isdigit():
        agg[(cur, int(r[0]))] = (int(r[ie]) if r[ie].isdigit() else 0, int(r[isamp]) if r[isamp].isdigit() else 0, r[1])
tot = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print("total inst", tot, "samples", ts)
for (f, l), v in sorted(agg.items()):
    if v[0] > tot * thr / 100 or v[1] > ts * thr / 100:
        print(f"{f[:16]:16s} {l:4d} {100*v[0]/tot:5.2f}% inst {100*v[1]/ts:5.2f}% samp  {v[2].strip()[:105]}")

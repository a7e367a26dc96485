"""Turn the raw output of scripts/gpu_collect.sh (gpurun_out/) into the tracked files under profiles/."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = os.path.join(G, "prof_final.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_tma_ld.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'sm__cycles_elapsed.max']
with open(os.path.join(P, tag + "_final_ncu_full_summary.csv"), "w") as f:
    w = csv.writer(f); w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
    for k in keep:
        if k in hdr:
            i = hdr.index(k); w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
r = rows[2]
mul = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
def val(k): i = hdr.index(k); return float(r[i]) * mul.get(units[i], 1)
traffic = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
tj = os.path.join(P, "traffic.json")
td = json.load(open(tj)) if os.path.exists(tj) else {}
td["forest4096"] = traffic
td["_source"] = ("profiles/%s_final_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum of gc_mesh_kernel, one launch "
                 "(ncu --set full --clock-control none)" % tag)
json.dump(td, open(tj, "w"), indent=1)
for src, dst in (("launches_final.csv", "_final_launches.csv"), ("phase_forest.txt", "_phase_cycles_forest4096.txt"), ("phase_checker.txt", "_phase_cycles_checker4096.txt"),
                 ("BENCH_local.json", "_bench_forest4096.json"), ("BENCH_ref_local.json", "_bench_reference_arm.json"), ("bench_islands4096.json", "_bench_islands4096.json"),
                 ("bench_checker1024.json", "_bench_checker1024.json"), ("bench_noise1024.json", "_bench_noise1024.json"), ("pytest_gpu.txt", "_pytest_gpu.txt")):
    if os.path.exists(os.path.join(G, src)): shutil.copy(os.path.join(G, src), os.path.join(P, tag + dst))
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, "0.5"], capture_output=True, text=True).stdout
open(os.path.join(P, tag + "_final_ncu_source_hotspots.txt"), "w").write(summ)
print("traffic", traffic)

#!/bin/bash
# Run on the GPU box (via gpurun): everything profiles/ is built from.  Output goes to gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/BENCH_local.json 2> gpurun_out/BENCH_local.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/BENCH_ref_local.json 2>/dev/null
for w in islands4096 checker1024 noise1024; do timeout 300 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-preprocess --no-stress > gpurun_out/bench_$w.json 2>/dev/null; done
timeout 100 python scripts/gpu_prof.py forest > gpurun_out/phase_forest.txt 2>&1
timeout 100 python scripts/gpu_prof.py checker > gpurun_out/phase_checker.txt 2>&1
for k in forest islands; do timeout 200 python scripts/gpu_light_bench.py $k 36 2>&1 | tail -4 > gpurun_out/light_bench_$k.txt; done
# ncu passes only after the same command has exited 0 without ncu
timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-preprocess --no-stress --no-surface-bounded --no-config5-leg > gpurun_out/plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-preprocess --no-stress --no-surface-bounded --no-config5-leg > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gc_mesh -s 4 -c 1 -o gpurun_out/prof_final -f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-preprocess --no-stress --no-surface-bounded --no-config5-leg > gpurun_out/ncu_f.log 2>&1
for k in forest islands; do timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/light_launches_$k.csv python scripts/gpu_light_profile.py $k > /dev/null 2>&1; done
cat gpurun_out/pytest_gpu.txt
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/BENCH_local.json") if l.startswith("{")][-1]
print("value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY

#!/bin/bash
# round 2, first call: where round 1's HEAD stands on today's box + ncu captures of the dense worlds
set -u
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nproc > $O/host.txt; lscpu | head -25 >> $O/host.txt; free -g >> $O/host.txt; nvidia-smi topo -m >> $O/host.txt 2>&1; numactl -H >> $O/host.txt 2>&1
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
for i in 1 2 3; do timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref_$i.json 2>/dev/null; done
for w in checker1024 noise1024; do
  timeout 200 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-preprocess > $O/bench_$w.json 2>/dev/null
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gc_mesh -s 4 -c 1 -o $O/prof_$w -f python bench.py --workload $w --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-preprocess > $O/ncu_$w.log 2>&1
done
cat $O/pytest_gpu.txt

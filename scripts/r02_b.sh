#!/bin/bash
# round 2, second call (1 GPU): tests + default bench + reference arm at the new HEAD
set -u
O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $O/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref.json 2>$O/ref.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-surface-hint --no-preprocess --no-cpu-baseline > $O/bench_nohint.json 2> $O/bench_nohint.err
cat $O/pytest_gpu.txt; tail -3 $O/bench.err

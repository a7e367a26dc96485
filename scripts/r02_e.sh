#!/bin/bash
set -u
O=gpurun_out/r02e; mkdir -p $O
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_halo_p2p.py -x -q 2>&1 | tail -1; done > $O/halo_loop.txt
timeout 600 python scripts/gpu_fused_debug.py > $O/fused_debug.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > $O/pytest_gpu.txt
GCMESH_TRACE=1 HINT=1 timeout 120 python scripts/gpu_trace.py > $O/trace_hint.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-preprocess > $O/bench.json 2> $O/bench.err
cat $O/halo_loop.txt $O/fused_debug.txt $O/pytest_gpu.txt; tail -22 $O/trace_hint.txt

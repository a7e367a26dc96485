#!/bin/bash
set -u
O=gpurun_out/r02m; mkdir -p $O
GCMESH_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-preprocess --no-stress --no-cpu-baseline --no-parity > $O/bench.json 2> $O/bench.err; echo rc $?
grep -n "total\|strip  0\|strip 17\|last kernel" $O/bench.err | tail -40
NOLIGHT=1 HINT=1 GCMESH_TRACE=1 timeout 120 python scripts/gpu_trace.py 2>&1 | grep "total\|^call" | tail -4

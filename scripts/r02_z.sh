#!/bin/bash
set -u
for lib in gamecraft_b200/csrc/variants/libgcmesh_t384.so gamecraft_b200/csrc/variants/libgcmesh_t416.so gamecraft_b200/csrc/libgcmesh.so; do
  for w in forest4096 checker1024 noise1024; do
    GCMESH_LIB=$PWD/$lib timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-preprocess --no-stress 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $lib)', d['config']['workload'], 'ms', round(d['ms_per_step'],4), 'bounded', round(d['surface_bounded']['ms_per_step'],4), 'parity', d['parity']['mismatches'], d['kernel_info'])"
  done
done

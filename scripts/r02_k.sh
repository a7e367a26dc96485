#!/bin/bash
set -u
O=gpurun_out/r02k; mkdir -p $O
for n in 6 8; do
  for k in forest islands; do
    echo "ctas/sm $n $k"; GCMESH_LIGHT_CTAS_PER_SM=$n timeout 300 python scripts/gpu_light_bench.py $k 36 2>&1 | grep "device light"
  done
done 2>&1 | tee $O/sweep.txt

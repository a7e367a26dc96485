#!/bin/bash
set -u
O=gpurun_out/r02y; mkdir -p $O
for k in forest islands; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gc_relax_all\|gc_light_scan -c 2 -o $O/light_$k -f python scripts/gpu_light_profile.py $k > $O/ncu_$k.log 2>&1; echo rc $?
done
ls -la $O

#!/bin/bash
# light fill: timings on forest / islands windows + ncu launch lists
set -u
O=gpurun_out/r02i; mkdir -p $O
for k in forest islands; do
  timeout 300 python scripts/gpu_light_bench.py $k 36 > $O/light_$k.txt 2>&1; tail -6 $O/light_$k.txt
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/light_launches_$k.csv python scripts/gpu_light_profile.py $k > $O/ncu_$k.log 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open("$O/light_launches_$k.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
for r in rows[1:][-12:]:
    print(r[ki][:60], r[vi])
PY
done

#!/bin/bash
set -u
for n in 4 8 16 36; do echo "size $n"; timeout 300 python scripts/gpu_light_bench.py forest $n 2>&1 | grep "device light\|sweeps"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_light8.csv python scripts/gpu_light_bench.py forest 8 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02r_light8.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
for r in rows[1:][-8:]:
    print(r[ki][:50], r[vi])
PY

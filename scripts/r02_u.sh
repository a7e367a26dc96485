#!/bin/bash
set -u
O=gpurun_out/r02u; mkdir -p $O
for v in plain halo; do
  if [ $v = halo ]; then export GCMESH_FORCE_HALO_KERNEL=1; else unset GCMESH_FORCE_HALO_KERNEL; fi
  for w in forest4096 forest65536; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-preprocess --no-stress --no-surface-bounded > $O/bench_${v}_$w.json 2> $O/bench_${v}_$w.err
  python - <<PY
import json
d = [json.loads(l) for l in open("$O/bench_${v}_$w.json") if l.startswith("{")][-1]
print("$v $w", "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "parity", d["parity"]["mismatches"], "cold", (d.get("cold_order") or {}).get("kernel_ms"))
PY
  done
done

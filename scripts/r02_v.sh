#!/bin/bash
set -u
O=gpurun_out/r02v; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_halo_p2p.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3
export GCMESH_FORCE_HALO_KERNEL=1
for w in forest4096 forest65536; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-preprocess --no-stress --no-surface-bounded > $O/bench_halo_$w.json 2> $O/bench_halo_$w.err
  python - <<PY
import json
d = [json.loads(l) for l in open("$O/bench_halo_$w.json") if l.startswith("{")][-1]
print("halo $w", "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "parity", d["parity"]["mismatches"], "cold", (d.get("cold_order") or {}).get("kernel_ms"))
PY
done

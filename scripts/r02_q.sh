#!/bin/bash
# quick kernel A/B: parity tests of the mesher + device-resident bench line only
set -u
O=gpurun_out/r02q; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for w in forest4096 checker1024 noise1024 islands4096; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-preprocess --no-stress > $O/bench_$w.json 2> $O/bench_$w.err
python - <<PY
import json
d = [json.loads(l) for l in open("$O/bench_$w.json") if l.startswith("{")][-1]
print("$w", "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"]["mismatches"], "bounded", (d.get("surface_bounded") or {}).get("ms_per_step"), "cold", (d.get("cold_order") or {}).get("kernel_ms"))
PY
done

#!/bin/bash
set -u
O=gpurun_out/r02w; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
for v in pdl nopdl; do
  if [ $v = nopdl ]; then export GCMESH_NO_PDL=1; else unset GCMESH_NO_PDL; fi
  for w in forest4096 checker1024; do
  timeout 300 python bench.py --workload $w --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-preprocess --no-stress --no-surface-bounded > $O/bench_${v}_$w.json 2> $O/bench_${v}_$w.err
  python - <<PY
import json
d = [json.loads(l) for l in open("$O/bench_${v}_$w.json") if l.startswith("{")][-1]
print("$v $w", "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "parity", d["parity"]["mismatches"], "between", round(d["between_kernels_us"], 2))
PY
  done
done
unset GCMESH_NO_PDL
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-preprocess --no-stress --no-config5-leg > $O/bench_e2e.json 2> $O/bench_e2e.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02w/bench_e2e.json") if l.startswith("{")][-1]
e = d["e2e"]; print("e2e", round(e["value"]), round(e["ms_per_step"], 3), e.get("parity_mismatches"), "| other", round(e["light_arrays_in"]["ms_per_step"], 3), e["light_arrays_in"].get("parity_mismatches"))
PY

#!/bin/bash
set -u
O=gpurun_out/r02l; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "without_light or surface_bounds or mesh_host" 2>&1 | tail -30 > $O/pytest.txt; cat $O/pytest.txt
GCBENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-preprocess --no-stress > $O/bench.json 2> $O/bench.err; echo rc $?
tail -5 $O/bench.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02l/bench.json") if l.startswith("{")][-1]
e = d["e2e"]
print("e2e", e["value"], e["ms_per_step"], e["h2d_bytes_per_step"], e.get("parity_mismatches"), e["call"][:60])
o = e.get("light_arrays_in") or e.get("light_on_device")
print("other", o and (o["value"], o["ms_per_step"], o["h2d_bytes_per_step"], o.get("parity_mismatches")))
PY
HINT=1 GCMESH_TRACE=1 timeout 120 python scripts/gpu_trace.py 2>&1 | tail -24 > $O/trace.txt
NOLIGHT=1 HINT=1 GCMESH_TRACE=1 timeout 120 python scripts/gpu_trace.py 2>&1 | tail -24 > $O/trace_nolight.txt

#!/bin/bash
set -u
O=gpurun_out/r02e2; mkdir -p $O
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-stress --no-config5-leg --no-surface-bounded > $O/bench.json 2> $O/bench.err; echo rc $?; tail -2 $O/bench.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02e2/bench.json") if l.startswith("{")][-1]
print(json.dumps(d["edit_path"])[:900])
PY

#!/bin/bash
# round 2, session 3: GPU tests at HEAD (new generators, async light fill, validate_blocks, edit refusals) + default bench line
set -u
O=gpurun_out/r02h; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -60 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
GCBENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo rc $?
tail -3 $O/bench.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02h/bench.json") if l.startswith("{")][-1]
print("value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("bounded", d.get("surface_bounded")); print("preprocess", json.dumps(d.get("preprocess"))[:900])
PY

#!/bin/bash
set -u
O=gpurun_out/r02s; mkdir -p $O
T0=$(date +%s.%N); python bench.py > $O/bench.json 2> $O/bench.err; echo rc $? wall $(python -c "import time,sys; print(round(time.time()-float(sys.argv[1]),1))" $T0)
tail -2 $O/bench.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02s/bench.json") if l.startswith("{")][-1]
c = d["config5_single_gpu_leg"]
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("config5", c["value"], c["ms_per_step"], c["parity"]["mismatches"], c["roofline"]["kernel_ms"], c["roofline"]["frac"], (c.get("surface_bounded") or {}).get("ms_per_step"))
PY

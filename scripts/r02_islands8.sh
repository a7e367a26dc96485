#!/bin/bash
# BASELINE config 3 (islands65536) at HEAD: N = 8 and N = 1 on one 8-GPU box
set -u
O=gpurun_out/r02isl; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --workload islands65536 --no-weak-leg > $O/islands65536_n8.json 2> $O/islands65536_n8.err; echo "n8 rc $?"
timeout 900 python bench.py --gpus 1 --workload islands65536 --no-e2e --no-cpu-baseline --no-preprocess --no-stress > $O/islands65536_n1.json 2> $O/islands65536_n1.err; echo "n1 rc $?"
python - <<'PY'
import json
for n in (1, 8):
    d = [json.loads(l) for l in open("gpurun_out/r02isl/islands65536_n%d.json" % n) if l.startswith("{")][-1]
    e = d.get("e2e") or {}
    sb = d.get("surface_bounded") or {}
    print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"]["mismatches"], "quads/chunk", round(d["quads_per_chunk"], 1),
          "| bounded", sb.get("ms_per_step"), sb.get("parity_mismatches"), "| e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_mismatches"))
PY

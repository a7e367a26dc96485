#!/bin/bash
set -u
O=gpurun_out/r02n2c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_halo_p2p.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_halo_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-e2e > $O/scale_n2.json 2> $O/scale_n2.err; echo rc $?
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02n2c/scale_n2.json") if l.startswith("{")][-1]
w = d.get("weak_scaling_leg") or {}
print(d["config"]["workload"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "parity", d["parity"], "| weak", w.get("value"), w.get("ms_per_step"), (w.get("parity") or {}).get("mismatches"))
PY

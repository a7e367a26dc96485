#!/bin/bash
# Two GPUs, at HEAD, exactly as the driver launches it: halo tests, the default bench line and the reference arm under torchrun.
set -u
O=gpurun_out/r02n2d; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_halo_p2p.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_halo_2gpu.txt
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > $O/scale_n2.json 2> $O/scale_n2.err; echo rc $? wall $(( $(date +%s) - S )) s
S=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err; echo rc $? wall $(( $(date +%s) - S )) s
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02n2d/scale_n2.json") if l.startswith("{")][-1]
w = d.get("weak_scaling_leg") or {}
e = d.get("e2e") or {}
print(d["config"]["workload"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "parity", d["parity"].get("mismatches"), "| e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_mismatches"), "| weak", w.get("value"), w.get("ms_per_step"))
r = [json.loads(l) for l in open("gpurun_out/r02n2d/ref_n2.json") if l.startswith("{")][-1]
print("reference arm", r.get("value"), r.get("unit"), r["config"])
PY

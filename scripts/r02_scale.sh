#!/bin/bash
# the driver's scaling sequence at HEAD on one 8-GPU box: default arguments, N = 1, 2, 4, 8 back to back, the reference arm first
set -u
O=gpurun_out/r02scale; mkdir -p $O
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/ref_n1.json 2> $O/ref_n1.err; echo "ref rc $?"
for n in 1 2 4 8; do
  T0=$(date +%s)
  if [ $n = 1 ]; then timeout 900 python bench.py --gpus 1 > $O/scale_n1.json 2> $O/scale_n1.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n > $O/scale_n$n.json 2> $O/scale_n$n.err; fi
  echo "N=$n rc $? wall $(( $(date +%s) - T0 )) s"
done
python - <<'PY'
import json
v = {}
for n in (1, 2, 4, 8):
    try:
        d = [json.loads(l) for l in open("gpurun_out/r02scale/scale_n%d.json" % n) if l.startswith("{")][-1]
    except Exception as ex:
        print(n, "ERR", ex); continue
    e = d.get("e2e") or {}
    c5 = d.get("config5_single_gpu_leg") or {}
    w = d.get("weak_scaling_leg") or {}
    v[n] = d
    print(n, d["config"]["workload"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "parity", d["parity"]["mismatches"],
          "| e2e", round(e.get("value", 0)), round(e.get("ms_per_step", 0), 2), e.get("parity_mismatches"), "| weak", w.get("value"), w.get("ms_per_step"), "| c5", c5.get("value"), c5.get("ms_per_step"))
r = [json.loads(l) for l in open("gpurun_out/r02scale/ref_n1.json") if l.startswith("{")][-1]
print("reference", r["value"], r.get("step_ms"))
PY

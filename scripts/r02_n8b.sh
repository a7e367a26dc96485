#!/bin/bash
# eight GPUs, session 3: BASELINE config 5 (forest65536, the default at N > 1) with the new legs (e2e with the light on the device, surface_bounded)
set -u
O=gpurun_out/r02n8b; mkdir -p $O
nproc > $O/host.txt; free -g >> $O/host.txt; nvidia-smi topo -m >> $O/host.txt 2>&1
run() { # n name args...
  local n=$1 name=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 20 --warmup 5 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc $?"; tail -2 $O/$name.err
}
run 8 forest65536_n8
run 2 forest65536_n2
run 8 islands65536_n8 --workload islands65536 --no-e2e --no-weak-leg
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02n8b/*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        w = d.get("weak_scaling_leg") or {}
        e = d.get("e2e") or {}
        o = e.get("light_arrays_in") or e.get("light_on_device") or {}
        sb = d.get("surface_bounded") or {}
        print(f.split("/")[-1], d["n_gpus"], d["config"]["workload"], round(d["value"]), "ms", round(d["ms_per_step"], 4), "k", round(d["roofline"]["kernel_ms"], 4), "par", d["parity"]["mismatches"],
              "| e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_mismatches"), (e.get("call") or "")[:40], "| other", o.get("value"), o.get("ms_per_step"),
              "| bounded", sb.get("ms_per_step"), sb.get("parity_mismatches"), "| weak", w.get("value"), w.get("ms_per_step"))
    except Exception as ex:
        print(f, "ERR", ex)
PY

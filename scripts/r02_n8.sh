#!/bin/bash
# eight GPUs: BASELINE config 5 (forest65536) and config 3 (islands65536) at N = 8, 4, 2, 1 on one box, the halo tests one process per GPU
set -u
O=gpurun_out/r02n8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt
run() { # n name args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > $O/$name.json 2> $O/$name.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 20 --warmup 5 "$@" > $O/$name.json 2> $O/$name.err; fi
  echo "$name rc $?"
}
run 8 forest65536_n8
run 8 islands65536_n8 --workload islands65536 --no-e2e --no-weak-leg
run 4 forest65536_n4 --no-e2e --no-weak-leg
run 2 forest65536_n2 --no-e2e --no-weak-leg
run 1 forest65536_n1 --workload forest65536 --no-e2e --no-cpu-baseline --no-preprocess --no-stress
run 1 islands65536_n1 --workload islands65536 --no-e2e --no-cpu-baseline --no-preprocess --no-stress
run 8 forest65536_n8_p2p --exchange p2p --no-e2e --no-weak-leg
timeout 300 python -m pytest tests/test_gpu_halo_p2p.py -q 2>&1 | tail -4 > $O/pytest_halo.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02n8/*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        w = d.get("weak_scaling_leg") or {}
        print(f.split("/")[-1], d["n_gpus"], d["config"]["workload"], round(d["value"]), "ms", round(d["ms_per_step"], 4), "k", round(d["roofline"]["kernel_ms"], 4), "par", d["parity"]["mismatches"], "e2e", (d.get("e2e") or {}).get("value"), "weak", w.get("value"), w.get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
# two GPUs: the one-process-per-GPU halo tests (CUDA IPC peers, exchange kernel and fused form) and the N = 2 bench lines
set -u
O=gpurun_out/r02n2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt
timeout 600 python -m pytest tests/test_gpu_halo_p2p.py -v 2>&1 | tail -15 > $O/pytest_halo_2gpu.txt
cat $O/pytest_halo_2gpu.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $T bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo rc $?
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p --no-e2e --no-weak-leg > $O/bench_n2_p2p.json 2> $O/bench_n2_p2p.err; echo rc $?
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl --no-e2e --no-weak-leg > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; echo rc $?
tail -5 $O/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n2", "bench_n2_p2p", "bench_n2_nccl"):
    try:
        d = json.loads(open("gpurun_out/r02n2/%s.json" % f).read())
        print(f, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["parity"]["mismatches"], d["parity"].get("boundary_row_chunks"), d.get("e2e", {}).get("value"), (d.get("weak_scaling_leg") or {}).get("value"), (d.get("weak_scaling_leg") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY

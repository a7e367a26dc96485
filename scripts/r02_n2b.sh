#!/bin/bash
set -u
O=gpurun_out/r02n2; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
GCBENCH_VERBOSE=1 timeout 900 $T bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo rc $?
tail -40 $O/bench_n2.err; ls -la $O; head -c 600 $O/bench_n2.json

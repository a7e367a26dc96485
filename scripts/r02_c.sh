#!/bin/bash
set -u
O=gpurun_out/r02c; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_halo_p2p.py -x -q -k "local_peers and True" > $O/sanitizer.txt 2>&1
GCMESH_TRACE=1 HINT=1 timeout 120 python scripts/gpu_trace.py > $O/trace_hint.txt 2>&1
GCMESH_TRACE=1 timeout 120 python scripts/gpu_trace.py > $O/trace_nohint.txt 2>&1
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/ref_pin.json 2>/dev/null
GCREF_NO_PIN=1 timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/ref_nopin.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/ref_pin2.json 2>/dev/null
GCREF_NO_PIN=1 timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/ref_nopin2.json 2>/dev/null
grep -E "Invalid|at |ERROR SUMMARY|gcmesh_kernel.cu" $O/sanitizer.txt | head -30

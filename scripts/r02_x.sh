#!/bin/bash
set -u
O=gpurun_out/r02x; mkdir -p $O
timeout 200 python scripts/gpu_sanitize_small.py > $O/plain.txt 2>&1; echo "plain rc $?"; tail -3 $O/plain.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/gpu_sanitize_small.py > $O/memcheck.txt 2>&1; echo "memcheck rc $?"; tail -4 $O/memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/gpu_sanitize_small.py > $O/racecheck.txt 2>&1; echo "racecheck rc $?"; tail -4 $O/racecheck.txt

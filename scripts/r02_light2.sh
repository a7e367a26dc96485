#!/bin/bash
# light fill with the straight-down start: light tests, the calls that fill light per group, the fill's time
set -u
O=gpurun_out/r02light2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_light.py tests/test_gpu_edit.py tests/test_gpu_generate.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_light.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "light or host" 2>&1 | tail -3 | tee -a $O/pytest_light.txt
for k in forest islands; do timeout 200 python scripts/gpu_light_bench.py $k 36 2>&1 | tail -4 | tee $O/light_bench_$k.txt; done

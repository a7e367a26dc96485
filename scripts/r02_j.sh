#!/bin/bash
set -u
O=gpurun_out/r02j; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_light.py tests/test_gpu_edit.py tests/test_gpu_generate.py -q -m gpu 2>&1 | tail -15 > $O/pytest.txt; cat $O/pytest.txt
for k in forest islands; do
  timeout 300 python scripts/gpu_light_bench.py $k 36 > $O/light_$k.txt 2>&1; tail -4 $O/light_$k.txt
done

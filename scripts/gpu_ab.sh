#!/bin/bash
# development A/B: bench every libgcmesh variant under gamecraft_b200/csrc/variants on the listed workloads
for lib in gamecraft_b200/csrc/variants/*.so; do
  for w in ${@:-forest4096}; do
    GCMESH_LIB=$PWD/$lib timeout 200 python bench.py --workload $w --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $lib)', d['config']['workload'], round(d['value']), round(d['roofline']['kernel_ms'],4), d['kernel_info'])"
  done
done

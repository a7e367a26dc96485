#!/bin/bash
set -u
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python scripts/gpu_fused_debug.py > $O/fused_debug.txt 2>&1
timeout 300 python scripts/gpu_pcie_probe.py > $O/pcie.txt 2>&1
for c in 16 128 296; do GCMESH_PULL_CTAS=$c GCMESH_PULL_CTAS_RUN=$c timeout 300 python scripts/gpu_pcie_probe.py 2>&1 | grep "pull kernel" >> $O/pcie.txt; done
cat $O/fused_debug.txt $O/pcie.txt

#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "without_light" 2>&1 | tail -12

#!/usr/bin/env bash
# 8-lane qkrope backward: parity + kbench
set -uo pipefail
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -3
echo "== kbench"; KB_B=64 timeout 300 python tools/kbench.py 2>&1 | grep -E "adarms|qkrope"

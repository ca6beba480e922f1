#!/usr/bin/env bash
# adarms_fwd: rows per warp (block churn vs balance), R = 1 (no prefetch hack), 2, 4, 14
set -uo pipefail
L=voicebox-pytorch_b200/lib
for v in sm100a r2 r4 r14 sm100a; do
  echo "-- $v"
  VBX_LIB=$L/libvbx_$v.so timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "resid_norm" 2>&1 | tail -1
  VBX_LIB=$L/libvbx_$v.so KB_B=64 timeout 300 python tools/kbench.py 2>&1 | grep -E "adarms_fwd"
done

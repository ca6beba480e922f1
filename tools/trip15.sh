#!/usr/bin/env bash
# qkrope backward: U vectors in flight per thread x resident blocks
set -uo pipefail
L=voicebox-pytorch_b200/lib
for v in sm100a u4b2 u2b3 u2b4 u8b2; do
  echo "-- $v"
  VBX_LIB=$L/libvbx_$v.so timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k attention 2>&1 | tail -1
  VBX_LIB=$L/libvbx_$v.so KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope_bwd"
done

#!/usr/bin/env bash
# staged rope backward (token-major): bulk-copy piece size; and the register kernel under the same layout
set -uo pipefail
L=voicebox-pytorch_b200/lib
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k attention 2>&1 | tail -1
for v in sm100a p256 p512 p4096; do
  echo "-- $v"; VBX_LIB=$L/libvbx_$v.so KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope_bwd"
done
echo "-- regs"; VBX_QKROPE_BWD=regs KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope_bwd"

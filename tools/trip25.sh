#!/usr/bin/env bash
# token-major q^ / k^ / dq^ / dk^ between the rope and attention kernels vs head-major (libvbx_headmajor.so), same box
set -uo pipefail
L=voicebox-pytorch_b200/lib
echo "== tests (token-major)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
for v in sm100a headmajor sm100a headmajor; do
  echo "-- $v"; VBX_LIB=$L/libvbx_$v.so KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "attn_|qkrope"
done

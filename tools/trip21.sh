#!/usr/bin/env bash
# packed fp32 (FFMA2 / FADD2 / FMUL2, 3-input max) in the attention softmax loops: parity, repeatability, kbench A/B vs the scalar build
set -uo pipefail
L=voicebox-pytorch_b200/lib
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_at_size.py -m gpu -q 2>&1 | tail -2
for v in sm100a scalar sm100a scalar; do
  echo "-- $v"; VBX_LIB=$L/libvbx_$v.so KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "attn_"
done

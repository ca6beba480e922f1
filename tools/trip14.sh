#!/usr/bin/env bash
# grouped token order in the rope kernels: parity + kbench A/B (VBX_QKROPE_ORDER=strided = round-1 order)
set -uo pipefail
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -3
echo "== tests (strided order)"; VBX_QKROPE_ORDER=strided timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k attention 2>&1 | tail -2
echo "== kbench grouped"; KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope|attn"
echo "== kbench strided"; VBX_QKROPE_ORDER=strided KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope"

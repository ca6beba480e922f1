#!/usr/bin/env bash
# staged (bulk-copy) qkrope backward vs the register kernel
set -uo pipefail
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -2
echo "== kbench staged"; KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope"
echo "== kbench regs"; VBX_QKROPE_BWD=regs KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope"

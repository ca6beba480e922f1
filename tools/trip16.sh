#!/usr/bin/env bash
# qkrope forward: U vectors in flight x resident blocks; backward product = U4 / 2 blocks
set -uo pipefail
L=voicebox-pytorch_b200/lib
for v in sm100a f4b2 f1b2 f1b3; do
  echo "-- $v"
  VBX_LIB=$L/libvbx_$v.so timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k attention 2>&1 | tail -1
  VBX_LIB=$L/libvbx_$v.so KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "qkrope"
done
echo "== model tests"; timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_at_size.py -m gpu -q 2>&1 | tail -2
echo "== bench"; timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --optimizer flat 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"

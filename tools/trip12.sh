#!/usr/bin/env bash
# L2 bulk prefetch in the row kernels: with / without, standalone (kbench, B=64) and in the step
set -uo pipefail
OUT=gpurun_out/t12
mkdir -p "$OUT"
L=voicebox-pytorch_b200/lib
echo "== tests (kernels)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -2
echo "== kbench with prefetch"; KB_B=64 timeout 300 python tools/kbench.py 2>&1 | grep -E "adarms|qkrope|geglu|convpos"
echo "== kbench without"; KB_B=64 VBX_LIB=$L/libvbx_nopf.so timeout 300 python tools/kbench.py 2>&1 | grep -E "adarms|qkrope"
echo "== bench step A/B"
for v in sm100a nopf; do
VBX_LIB=$L/libvbx_$v.so timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --optimizer flat > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"
done
python - <<'PY'
import json
for n in ('sm100a', 'nopf'):
    try:
        d = json.load(open(f'gpurun_out/t12/bench_{n}.json'))
        print(n, round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'launches', d['gpu_launches'])
        for k in d.get('kernels', []):
            pass
    except Exception as e:
        print(n, 'failed:', e)
PY

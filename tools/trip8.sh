#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/t8
mkdir -p "$OUT"
echo "== new kernel tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ff2_dgrad or gemm or ff1" > "$OUT/tests_gemm.log" 2>&1; tail -4 "$OUT/tests_gemm.log"
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q -rs > "$OUT/tests.log" 2>&1; tail -4 "$OUT/tests.log"; grep -E "^FAILED" "$OUT/tests.log"
echo "== bench A/B: fused FF backward on / off (same box)"
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --optimizer flat > "$OUT/bench_on.json" 2> "$OUT/bench_on.err"
VBX_FUSED_FF_BWD=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --optimizer flat > "$OUT/bench_off.json" 2> "$OUT/bench_off.err"
python - <<'PY'
import json
for n in ('on', 'off'):
    try:
        d = json.load(open(f'gpurun_out/t8/bench_{n}.json'))
        print(n, round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'launches', d['gpu_launches'])
    except Exception as e:
        print(n, 'failed:', e)
PY
echo "== step profile"; timeout 300 python tools/step_profile.py > "$OUT/step_profile.txt" 2>&1; head -30 "$OUT/step_profile.txt"

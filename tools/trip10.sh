#!/usr/bin/env bash
# FF backward epilogue after the TMA-load rewrite: parity, attribution, then the step A/B
set -uo pipefail
OUT=gpurun_out/t10
mkdir -p "$OUT"
L=voicebox-pytorch_b200/lib
echo "== tests product"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ff2_dgrad or operand_pack" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -k "operand_pack or pack" 2>&1 | tail -3
for v in sm100a abl1 abl4 abl8 abl13; do
  echo "-- $v"; VBX_GEMM_BENCH=bwd VBX_LIB=$L/libvbx_$v.so timeout 200 python tools/gemm_bench.py 2>&1 | tail -1
done
echo "== bench A/B: fused FF backward on / off (same box)"
VBX_FUSED_FF_BWD=1 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --optimizer flat > "$OUT/bench_on.json" 2> "$OUT/bench_on.err"
VBX_FUSED_FF_BWD=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --optimizer flat > "$OUT/bench_off.json" 2> "$OUT/bench_off.err"
python - <<'PY'
import json
for n in ('on', 'off'):
    try:
        d = json.load(open(f'gpurun_out/t10/bench_{n}.json'))
        print(n, round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'launches', d['gpu_launches'])
    except Exception as e:
        print(n, 'failed:', e)
PY

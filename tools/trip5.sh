#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/t5
mkdir -p "$OUT"
L=$PWD/voicebox-pytorch_b200/lib
echo "== repeatability (product with the O_FULL phase fix)"; timeout 120 python tools/debug_attn_repeat.py 2>&1 | tee "$OUT/repeat_product.txt"
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q -rs > "$OUT/tests.log" 2>&1; tail -4 "$OUT/tests.log"; grep -E "^FAILED" "$OUT/tests.log"
echo "== kbench all kernels (B=64)"; KB_B=64 KB_ITERS=8 timeout 300 python tools/kbench.py > "$OUT/kbench_all.txt" 2>&1; cat "$OUT/kbench_all.txt"
echo "== stagger variants (fwd only)"
for v in stag1200 stag2000 stag1600p1; do
  VBX_LIB=$L/libvbx_$v.so timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > "$OUT/tests_$v.log" 2>&1; echo "$v: $(tail -1 $OUT/tests_$v.log)"
  VBX_LIB=$L/libvbx_$v.so timeout 100 python tools/debug_attn_repeat.py 2>&1 | head -2
  VBX_LIB=$L/libvbx_$v.so KB_ONLY=attn KB_B=64 KB_ITERS=8 timeout 200 python tools/kbench.py 2>&1 | grep -i "attn_fwd" | tee "$OUT/kbench_$v.txt"
done
echo "== trace"; VBX_LIB=$L/libvbx_trace.so timeout 200 python tools/trace_attn.py > "$OUT/trace_attention.txt" 2>&1; tail -12 "$OUT/trace_attention.txt"
echo "== bench default (fused FF1 on) with flat optimizer, full line"; timeout 700 python bench.py --optimizer flat > "$OUT/bench_flat.json" 2> "$OUT/bench_flat.err"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --no-sdpa > "$OUT/bench_torchopt.json" 2> "$OUT/bench_torchopt.err"
python - <<'PY'
import json
for n in ('flat', 'torchopt'):
    try:
        d = json.load(open(f'gpurun_out/t5/bench_{n}.json'))
        print(n, round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'launches', d.get('gpu_launches'))
        if n == 'flat':
            for k, v in d['kernels'].items(): print('  ', k, round(v['avg_us'], 1), round(v['frac'], 3))
            print('   sample', d['sample']['value'], d['sample']['ms_per_ode_step'], 'cpu', d['cpu_baseline']['value'])
    except Exception as e:
        print(n, 'failed:', e)
PY

#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/t6
mkdir -p "$OUT"
L=$PWD/voicebox-pytorch_b200/lib
export VBX_ATTN_FWD=3
echo "== v3 attention tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > "$OUT/tests_attn_v3.log" 2>&1; tail -3 "$OUT/tests_attn_v3.log"
echo "== v3 repeatability"; timeout 120 python tools/debug_attn_repeat.py 2>&1 | tee "$OUT/repeat_v3.txt"
echo "== v3 kbench"; KB_ONLY=attn KB_B=64 KB_ITERS=8 timeout 200 python tools/kbench.py 2>&1 | grep -i "attn" | tee "$OUT/kbench_v3.txt"
echo "== v3 + poly1"; VBX_LIB=$L/libvbx_poly1.so KB_ONLY=attn KB_B=64 KB_ITERS=8 timeout 200 python tools/kbench.py 2>&1 | grep -i "attn_fwd" | tee "$OUT/kbench_v3_poly1.txt"
echo "== v3 trace"; VBX_LIB=$L/libvbx_trace.so timeout 200 python tools/trace_attn.py > "$OUT/trace_attention_v3.txt" 2>&1; tail -24 "$OUT/trace_attention_v3.txt"
echo "== v3 model tests"; timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_at_size.py -m gpu -q > "$OUT/tests_model_v3.log" 2>&1; tail -3 "$OUT/tests_model_v3.log"
unset VBX_ATTN_FWD
echo "== v2 kbench (same box)"; KB_B=64 KB_ITERS=8 timeout 300 python tools/kbench.py 2>&1 | grep -i "attn\|adarms" | tee "$OUT/kbench_v2.txt"
echo "== bench: fwd v3 vs v2 (same box)"
VBX_ATTN_FWD=3 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sdpa --optimizer flat --sample-steps 16 > "$OUT/bench_v3.json" 2> "$OUT/bench_v3.err"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sdpa --optimizer flat --sample-steps 16 > "$OUT/bench_v2.json" 2> "$OUT/bench_v2.err"
python - <<'PY'
import json
for n in ('v3', 'v2'):
    try:
        d = json.load(open(f'gpurun_out/t6/bench_{n}.json'))
        k = d['kernels']
        print(n, round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'attn fwd', round(k['vbx_attn_fwd']['avg_us'], 1), 'bwd', round(k['vbx_attn_bwd']['avg_us'], 1),
              'adarms', round(k['vbx_adarms_fwd']['avg_us'], 1), round(k['vbx_adarms_bwd']['avg_us'], 1), 'sample', round(d['sample']['value'], 2))
    except Exception as e:
        print(n, 'failed:', e)
PY

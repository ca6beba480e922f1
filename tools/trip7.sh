#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/t7
mkdir -p "$OUT"
echo "== full suite (fwd v1 default)"; timeout 1200 python -m pytest tests -m gpu -q -rs > "$OUT/tests.log" 2>&1; tail -4 "$OUT/tests.log"; grep -E "^FAILED" "$OUT/tests.log"
echo "== bench: eager vs whole-step CUDA graph (same box)"
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --graph off > "$OUT/bench_eager.json" 2> "$OUT/bench_eager.err"
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --graph on > "$OUT/bench_graph.json" 2> "$OUT/bench_graph.err"
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sdpa --no-sample --graph off --optimizer flat > "$OUT/bench_eager_flat.json" 2> "$OUT/bench_eager_flat.err"
python - <<'PY'
import json
for n in ('eager', 'graph', 'eager_flat'):
    try:
        d = json.load(open(f'gpurun_out/t7/bench_{n}.json'))
        print(n, round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'e2e', round(d['e2e']['value']), 'eager_ms', d.get('eager_ms_per_step'), d['config'].get('step_graph'), 'launches', d['gpu_launches'])
    except Exception as e:
        print(n, 'failed:', e)
PY
tail -3 "$OUT/bench_graph.err"

#!/usr/bin/env bash
# ONE gpurun trip for the host-side (PyTorch-level) experiments on the training step: model parity with the flag on, then
# the bench line with and without it.   gpurun --timeout 600 -- 'bash tools/step_experiments.sh'
# Flags: VBX_BATCHED_GB=1       all adaptive norms' gamma/beta from one batched GEMM (ops.batched_affine)
#        --optimizer flat       FlatAdam: clip + Adam as one vbx_adam_step launch over the flat buffers
set -uo pipefail
OUT=gpurun_out/stepexp
mkdir -p "$OUT"
VBX_BATCHED_GB=1 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q > "$OUT/tests_batched_gb.log" 2>&1
tail -1 "$OUT/tests_batched_gb.log"
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k flat_adam > "$OUT/tests_flat_adam.log" 2>&1
tail -1 "$OUT/tests_flat_adam.log"
timeout 240 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample > "$OUT/bench_base.json" 2> "$OUT/bench_base.err"
VBX_BATCHED_GB=1 timeout 240 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample > "$OUT/bench_batched_gb.json" 2> "$OUT/bench_batched_gb.err"
timeout 240 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --optimizer flat > "$OUT/bench_flat_adam.json" 2> "$OUT/bench_flat_adam.err"
VBX_BATCHED_GB=1 timeout 240 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --optimizer flat > "$OUT/bench_both.json" 2> "$OUT/bench_both.err"
python - <<'PY'
import json
for n in ('base', 'batched_gb', 'flat_adam', 'both'):
    try:
        d = json.load(open(f'gpurun_out/stepexp/bench_{n}.json'))
        print(n, d['ms_per_step'], 'ms/step', d['value'], d['unit'], 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(n, 'failed:', e)
PY

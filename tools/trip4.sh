#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/t4
mkdir -p "$OUT"
L=$PWD/voicebox-pytorch_b200/lib
echo "== repeatability: product, then bisect variants"
timeout 120 python tools/debug_attn_repeat.py 2>&1 | tee "$OUT/repeat_product.txt"
for d in 1 2 4 8; do echo "-- dbg$d"; VBX_LIB=$L/libvbx_dbg$d.so timeout 120 python tools/debug_attn_repeat.py 2>&1 | tee "$OUT/repeat_dbg$d.txt"; done
echo "-- fwd v1"; VBX_ATTN_FWD_V1=1 timeout 120 python tools/debug_attn_repeat.py 2>&1 | tee "$OUT/repeat_v1.txt"
echo "== bench: packed (kernel-mode wgrad) / unpacked / packed+fused FF1 / packed+flat"
for cfg in "packed:" "unpacked:VBX_PACKED=0" "fusedff1:VBX_FUSED_FF1=1" ; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --no-sdpa > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
done
VBX_FUSED_FF1=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --no-sdpa --optimizer flat > "$OUT/bench_fusedff1_flat.json" 2> "$OUT/bench_fusedff1_flat.err"
python - <<'PY'
import json
for n in ('packed', 'unpacked', 'fusedff1', 'fusedff1_flat'):
    try:
        d = json.load(open(f'gpurun_out/t4/bench_{n}.json'))
        print(n, round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(n, 'failed:', e)
PY
echo "== step profile packed"; VBX_FUSED_FF1=1 timeout 300 python tools/step_profile.py > "$OUT/step_profile_packed.txt" 2>&1; head -34 "$OUT/step_profile_packed.txt"
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q -rs > "$OUT/tests.log" 2>&1; tail -6 "$OUT/tests.log"; grep -E "^FAILED" "$OUT/tests.log"

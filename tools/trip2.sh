#!/usr/bin/env bash
# Round-2 trip 2: forward attention v2 + 3-issuer backward, operand packs, tcgen05 GEMM -- parity, A/B timings, bench.
set -uo pipefail
OUT=gpurun_out/t2
mkdir -p "$OUT"
L=$PWD/voicebox-pytorch_b200/lib
echo "== attention + gemm + pack tests first (small, catch deadlocks early)"
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or umma" > "$OUT/tests_attn.log" 2>&1; tail -3 "$OUT/tests_attn.log"
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm or ff1" > "$OUT/tests_gemm.log" 2>&1; tail -3 "$OUT/tests_gemm.log"
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q -rs --durations=10 > "$OUT/tests.log" 2>&1; tail -6 "$OUT/tests.log"
echo "== kbench attention: product (fwd v2, 3 issuers) / fwd v1 / poly1 / 1 issuer"
KB_ONLY=attn KB_B=64 KB_ITERS=5 timeout 200 python tools/kbench.py 2>&1 | grep -i "attn\|qkrope" > "$OUT/kbench_product.txt"; cat "$OUT/kbench_product.txt"
VBX_ATTN_FWD_V1=1 KB_ONLY=attn KB_B=64 KB_ITERS=5 timeout 200 python tools/kbench.py 2>&1 | grep -i attn > "$OUT/kbench_fwdv1.txt"; cat "$OUT/kbench_fwdv1.txt"
for v in poly1 iss1; do
  VBX_LIB=$L/libvbx_$v.so timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > "$OUT/tests_$v.log" 2>&1; echo "$v: $(tail -1 $OUT/tests_$v.log)"
  VBX_LIB=$L/libvbx_$v.so KB_ONLY=attn KB_B=64 KB_ITERS=5 timeout 200 python tools/kbench.py 2>&1 | grep -i attn > "$OUT/kbench_$v.txt"; cat "$OUT/kbench_$v.txt"
done
KB_ONLY=attn KB_B=16 KB_ITERS=5 timeout 200 python tools/kbench.py 2>&1 | grep -i attn > "$OUT/kbench_product_b16.txt"; cat "$OUT/kbench_product_b16.txt"
echo "== trace"; VBX_LIB=$L/libvbx_trace.so timeout 200 python tools/trace_attn.py > "$OUT/trace_attention.txt" 2>&1; tail -5 "$OUT/trace_attention.txt"
echo "== gemm bench"; timeout 300 python tools/gemm_bench.py > "$OUT/gemm_bench.txt" 2>&1; cat "$OUT/gemm_bench.txt"
echo "== bench (packed default)"; timeout 700 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/t2/bench.json'))
    print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
    for k, v in d['kernels'].items(): print(' ', k, round(v['avg_us'], 1), round(v['frac'], 3))
    print(' sample', d['sample']['value'], d['sample']['ms_per_ode_step'])
    print(' cpu', d['cpu_baseline'])
except Exception as e:
    print('bench failed', e)
PY
echo "== bench A/B"; 
VBX_PACKED=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --no-sdpa > "$OUT/bench_unpacked.json" 2> "$OUT/bench_unpacked.err"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --no-sdpa --optimizer flat > "$OUT/bench_flat.json" 2> "$OUT/bench_flat.err"
VBX_FUSED_FF1=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample --no-sdpa > "$OUT/bench_fusedff1.json" 2> "$OUT/bench_fusedff1.err"
python - <<'PY'
import json
for n in ('unpacked', 'flat', 'fusedff1'):
    try:
        d = json.load(open(f'gpurun_out/t2/bench_{n}.json'))
        print(n, d['ms_per_step'], 'ms/step', d['value'], 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(n, 'failed:', e)
PY
echo "== reference arm"; timeout 500 python bench.py --impl reference --steps 3 --warmup 2 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; tail -c 700 "$OUT/bench_ref.json"

#!/usr/bin/env bash
# round-2 evidence: full GPU suite, micro-benchmarks, the default bench line + reference arm, step profile, ncu launch list, ncu --set full
set -uo pipefail
OUT=gpurun_out/t17
mkdir -p "$OUT"
echo "== full GPU suite"; timeout 1500 python -m pytest tests -m gpu -q -rs > "$OUT/tests.log" 2>&1; tail -3 "$OUT/tests.log"; grep -E "^FAILED|^ERROR" "$OUT/tests.log"
echo "== kbench"; KB_B=64 timeout 300 python tools/kbench.py > "$OUT/kbench.txt" 2>&1; cat "$OUT/kbench.txt" | tail -16
echo "== gemm bench"; timeout 300 python tools/gemm_bench.py > "$OUT/gemm_bench.txt" 2>&1; cat "$OUT/gemm_bench.txt"
echo "== bench default"; timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; cat "$OUT/bench_default.json" | cut -c1-600
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; cat "$OUT/bench_reference.json" | cut -c1-600
echo "== step profile"; timeout 300 python tools/step_profile.py > "$OUT/step_profile.txt" 2>&1; sed -n 3,12p "$OUT/step_profile.txt" | cut -c1-150
echo "== ncu launch list (one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --profile-only > "$OUT/launches_bench.log" 2>&1
wc -l "$OUT/launches.csv"
echo "== ncu full: kbench kernels"
KB_B=64 KB_ITERS=1 KB_WARM=1 timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'adarms_|qkrope_|attn_fwd|attn_bwd|convpos|geglu_fwd_kernel|geglu_bwd_kernel' -c 40 -o "$OUT/kbench_full" -f python tools/kbench.py > "$OUT/ncu_kbench.log" 2>&1
echo "== ncu full: gemm kernels"
VBX_GEMM_BENCH=bwd timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_geglu_bwd' -c 2 -o "$OUT/gemm_bwd_full" -f python tools/gemm_bench.py > "$OUT/ncu_gemm_bwd.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_kernel' -c 6 -o "$OUT/gemm_fwd_full" -f python tools/gemm_bench.py > "$OUT/ncu_gemm_fwd.log" 2>&1
ls -la "$OUT"

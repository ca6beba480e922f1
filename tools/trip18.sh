#!/usr/bin/env bash
# round-2 evidence, part 1 (small files): default bench line + reference arm, step profile, ncu launch list of one step
set -uo pipefail
OUT=gpurun_out/t18
mkdir -p "$OUT"
echo "== bench default"; timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; cut -c1-300 "$OUT/bench_default.json"
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; cut -c1-200 "$OUT/bench_reference.json"
echo "== durpred"; timeout 600 python bench.py --workload durpred --no-cpu-baseline > "$OUT/bench_durpred.json" 2> "$OUT/bench_durpred.err"; cut -c1-300 "$OUT/bench_durpred.json"
echo "== step profile"; timeout 300 python tools/step_profile.py > "$OUT/step_profile.txt" 2>&1; sed -n 3,6p "$OUT/step_profile.txt" | cut -c1-150
echo "== ncu launch list (one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --profile-only > "$OUT/launches_bench.log" 2>&1
wc -l "$OUT/launches.csv"
echo "== ncu full (no source): row + attention kernels"
KB_B=64 KB_ITERS=1 KB_WARM=0 timeout 1200 ncu --set full --clock-control none \
  -k regex:'adarms_|qkrope_|attn_fwd|attn_bwd|convpos' -c 16 -o "$OUT/kbench_full" -f python tools/kbench.py > "$OUT/ncu_kbench.log" 2>&1
VBX_GEMM_BENCH=bwd timeout 600 ncu --set full --clock-control none -k regex:'gemm_geglu_bwd' -c 1 -o "$OUT/gemm_bwd_full" -f python tools/gemm_bench.py > "$OUT/ncu_gemm_bwd.log" 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gemm_bf16_kernel --launch-skip 39 -c 1 -o "$OUT/gemm_ff1_full" -f python tools/gemm_bench.py > "$OUT/ncu_gemm_ff1.log" 2>&1
du -sh "$OUT"; ls -la "$OUT"

#!/usr/bin/env bash
# Round-2 trip 1 (ONE gpurun call): full GPU suite incl. the new at-size parity tests, the default bench (fixed CPU baseline,
# SDPA baseline, 64-step graphed sampling), the reference arm, the MMA issue microbenchmark, every gated attention variant,
# and the host-side step experiments.  Everything lands under gpurun_out/t1/.
set -uo pipefail
OUT=gpurun_out/t1
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -rs --durations=15 > "$OUT/tests.log" 2>&1; tail -5 "$OUT/tests.log"
echo "== bench"; timeout 600 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; tail -c 600 "$OUT/bench.json"
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 2 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; tail -c 400 "$OUT/bench_ref.json"
echo "== umma bench"; VBX_LIB=$PWD/voicebox-pytorch_b200/lib/libvbx_trace.so timeout 120 python tools/umma_bench.py > "$OUT/umma_bench.txt" 2>&1; tail -14 "$OUT/umma_bench.txt"
echo "== attention variants"; timeout 1100 bash tools/attn_diagnose.sh > "$OUT/attn_diagnose.log" 2>&1; tail -30 "$OUT/attn_diagnose.log"
echo "== step experiments"; timeout 900 bash tools/step_experiments.sh > "$OUT/step_experiments.log" 2>&1; tail -8 "$OUT/step_experiments.log"
echo "== durpred"; timeout 200 python bench.py --workload durpred --steps 10 --warmup 3 > "$OUT/bench_durpred.json" 2> "$OUT/bench_durpred.err"; tail -c 400 "$OUT/bench_durpred.json"

#!/usr/bin/env bash
# final revision: GPU suite, default bench line, reference arm (same box)
set -uo pipefail
OUT=gpurun_out/t22
mkdir -p "$OUT"
echo "== full GPU suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2
echo "== kbench attn"; KB_B=64 KB_ONLY=attn timeout 300 python tools/kbench.py 2>&1 | grep -E "attn_|qkrope"
echo "== bench default"; timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; cut -c1-260 "$OUT/bench_default.json"
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; cut -c1-200 "$OUT/bench_reference.json"

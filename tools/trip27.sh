#!/usr/bin/env bash
# last check of the round: GPU suite, smoke(), default bench line
set -uo pipefail
OUT=gpurun_out/t27
mkdir -p "$OUT"
echo "== full GPU suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench default"; timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; cut -c1-260 "$OUT/bench_default.json"

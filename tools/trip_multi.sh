#!/usr/bin/env bash
# Multi-GPU experiments on ONE box (gpurun --gpus N): gradient exchange placement and NCCL CTA budget, whole-step CUDA graph + NCCL.
set -uo pipefail
N=${1:-4}
OUT=gpurun_out/multi_n$N
mkdir -p "$OUT"
run() {  # name, env..., -- bench args
  local name=$1; shift
  local envs=()
  while [[ $# -gt 0 && "$1" != "--" ]]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $N --steps 6 --warmup 3 --no-sample --no-cpu-baseline --no-sdpa "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  python - "$OUT/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], round(d['ms_per_step'], 2), 'ms/step', round(d['value']), d['config'].get('step_graph'), d['config'].get('allreduce'))
except Exception as e:
    print(sys.argv[2], 'failed:', e)
PY
}
run after FOO=1 -- --allreduce after
run overlap_default FOO=1 -- --allreduce overlap
run overlap_cta8 NCCL_MAX_CTAS=8 -- --allreduce overlap
run overlap_cta16 NCCL_MAX_CTAS=16 -- --allreduce overlap
run after_graph FOO=1 -- --allreduce after --graph on
tail -2 "$OUT/after_graph.err"

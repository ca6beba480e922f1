#!/usr/bin/env bash
# N = 2 check of the final revision (flat Adam default, fused FF backward): default exchange and the whole-step graph
set -uo pipefail
OUT=gpurun_out/t19
mkdir -p "$OUT"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-sample --no-sdpa "${@:3}" > "$OUT/$2.json" 2> "$OUT/$2.err"; python -c "
import json,sys
try:
    d=json.load(open('$OUT/$2.json')); print('$2', round(d['ms_per_step'],2), 'ms/step', round(d['value']), d['config'].get('optimizer'), d['config'].get('step_graph'))
except Exception as e:
    print('$2 failed', e); print(open('$OUT/$2.err').read()[-1500:])
"; }
run 29511 n2_default
run 29512 n2_torchopt --optimizer torch
run 29513 n2_graph --graph on

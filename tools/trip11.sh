#!/usr/bin/env bash
# FF backward epilogue with packed-fp32 math; ncu --set full of the HBM-bound row kernels (why are they at 0.46-0.7 of HBM?)
set -uo pipefail
OUT=gpurun_out/t11
mkdir -p "$OUT"
echo "== tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ff2_dgrad" 2>&1 | tail -2
echo "== bwd GEMM bench"; VBX_GEMM_BENCH=bwd timeout 200 python tools/gemm_bench.py 2>&1 | tail -1
echo "== ncu full: row kernels"
KB_B=64 KB_ITERS=1 KB_WARM=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'adarms_bwd|adarms_fwd|qkrope_bwd|qkrope_fwd' -c 8 -o "$OUT/rowkernels" -f python tools/kbench.py > "$OUT/ncu_rows.log" 2>&1
tail -3 "$OUT/ncu_rows.log"; ls -la "$OUT"

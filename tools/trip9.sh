#!/usr/bin/env bash
# FF backward epilogue attribution: product, ablations, and the smem-readback column-sum variant
set -uo pipefail
OUT=gpurun_out/t9
mkdir -p "$OUT"
L=voicebox-pytorch_b200/lib
echo "== tests product"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ff2_dgrad" 2>&1 | tail -2
echo "== tests ss0"; VBX_LIB=$L/libvbx_ss0.so timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ff2_dgrad" 2>&1 | tail -2
for v in sm100a abl1 abl2 abl4 abl8 abl3 abl7 abl15 ss0 ss1 ss8; do
  echo "-- $v"; VBX_GEMM_BENCH=bwd VBX_LIB=$L/libvbx_$v.so timeout 200 python tools/gemm_bench.py 2>&1 | tail -1
done

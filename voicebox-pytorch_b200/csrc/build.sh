#!/usr/bin/env bash
# Builds libvbx_sm100a.so in-tree (voicebox-pytorch_b200/lib/).  nvcc cross-compiles for sm_100a without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/../build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v)
OBJS=()
for f in api norm_ffn cfm_ode convpos qkrope optim pack gemm attn; do
  src="$HERE/$f.cu"; obj="$HERE/../build/$f.o"
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$HERE/common.cuh" -nt "$obj" || "$HERE/umma.cuh" -nt "$obj" || "$HERE/../../include/vbx.h" -nt "$obj" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" 2> "$HERE/../build/$f.ptxas.log" || { cat "$HERE/../build/$f.ptxas.log"; exit 1; }
  fi
  OBJS+=("$obj")
done
"$NVCC" -shared -o "$OUT/libvbx_sm100a.so" "${OBJS[@]}" -lcudart
echo "built $OUT/libvbx_sm100a.so"

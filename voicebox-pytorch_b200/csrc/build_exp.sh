#!/usr/bin/env bash
# Builds lib/libvbx_exp.so: the same library with EXPERIMENTAL kernel variants compiled in (default: -DVBX_EXP_TAIL=1,
# override with VBX_EXP_DEFS="-D... -D..."; VBX_EXP_OUT names the output, e.g. the clock64 trace build of tools/trace_attn.py:
# VBX_EXP_DEFS=-DVBX_TRACE VBX_EXP_OUT=libvbx_trace.so).  Never loaded by default -- select it with VBX_LIB=<path> (see _lib.py) to run
# the GPU tests / tools/kbench.py against it next to libvbx_sm100a.so.  An experiment graduates by flipping its macro's
# default in the source once it is parity-green and faster on the B200.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
bash "$HERE/build.sh" > /dev/null
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
DEFS=(${VBX_EXP_DEFS:--DVBX_EXP_TAIL=1})
OUT="${VBX_EXP_OUT:-libvbx_exp.so}"
TAG="${OUT%.so}"
SRC="${VBX_EXP_SRC:-attn}"     # which translation unit gets the experimental defines (attn or gemm)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "${DEFS[@]}" \
  -c "$HERE/$SRC.cu" -o "$HERE/../build/${SRC}_$TAG.o" 2> "$HERE/../build/${SRC}_$TAG.ptxas.log" || { cat "$HERE/../build/${SRC}_$TAG.ptxas.log"; exit 1; }
OBJS=()
for f in api norm_ffn cfm_ode convpos qkrope optim pack gemm attn; do [ "$f" = "$SRC" ] || OBJS+=("$HERE/../build/$f.o"); done
"$NVCC" -shared -o "$HERE/../lib/$OUT" "${OBJS[@]}" "$HERE/../build/${SRC}_$TAG.o" -lcudart
echo "built $HERE/../lib/$OUT (${DEFS[*]})"

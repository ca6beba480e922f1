#!/usr/bin/env bash
# ONE gpurun trip that refreshes every attention diagnostic (run from the repo root ON THE GPU BOX, after building here:
#   voicebox-pytorch_b200/csrc/build_exp.sh                                                   -> lib/libvbx_exp.so
#   VBX_EXP_DEFS=-DVBX_TRACE VBX_EXP_OUT=libvbx_trace.so voicebox-pytorch_b200/csrc/build_exp.sh
#   VBX_EXP_DEFS=-DVBX_EXP_POLY=1 VBX_EXP_OUT=libvbx_poly1.so voicebox-pytorch_b200/csrc/build_exp.sh   (and =2)
#   VBX_EXP_DEFS=-DVBX_EXP_DSBUF=1 VBX_EXP_OUT=libvbx_dsbuf.so voicebox-pytorch_b200/csrc/build_exp.sh
#   (every lib/libvbx_*.so other than the product and trace builds is tested and timed)
# ):  gpurun --timeout 900 -- 'bash tools/attn_diagnose.sh'
# Everything lands in gpurun_out/diag/.  Numbers printed under ncu are never bench values.
set -uo pipefail
OUT=gpurun_out/diag
mkdir -p "$OUT"
LIBDIR="$PWD/voicebox-pytorch_b200/lib"
PY=python

# 1. product library: parity, then the attention lines of the micro-benchmark at the bench geometry (B=64)
# (the product library is covered by the full suite)
KB_ONLY=attn KB_B=64 KB_ITERS=5 timeout 200 $PY tools/kbench.py 2>&1 | grep -i attn > "$OUT/kbench_product.txt"
# 2. every experimental library that was shipped: same parity tests, same benchmark
for lib in "$LIBDIR"/libvbx_*.so; do
  [[ -f "$lib" ]] || continue
  case "$(basename "$lib")" in libvbx_sm100a.so|libvbx_trace*.so) continue;; esac
  tag=$(basename "$lib" .so)
  VBX_LIB="$lib" timeout 300 $PY -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" > "$OUT/tests_$tag.log" 2>&1
  echo "$tag: $(tail -1 "$OUT/tests_$tag.log")" >> "$OUT/summary.txt"
  VBX_LIB="$lib" KB_ONLY=attn KB_B=64 KB_ITERS=5 timeout 200 $PY tools/kbench.py 2>&1 | grep -i attn > "$OUT/kbench_$tag.txt"
done
# 3. clock64 timeline of one CTA of each attention kernel (trace build)
for lib in "$LIBDIR"/libvbx_trace*.so; do
  [[ -f "$lib" ]] || continue
  VBX_LIB="$lib" timeout 200 $PY tools/trace_attn.py > "$OUT/$(basename "$lib" .so | sed s/libvbx_//)_attention.txt" 2>&1
done
# 4. one full ncu capture of each attention kernel at B=64, with the source page exported for tools/ncu_source_top.py
KB_ONLY=attn KB_B=64 KB_ITERS=1 KB_WARM=1 timeout 600 ncu --set full --import-source on --clock-control none \
  --kernel-name regex:'attn_(fwd|bwd)_kernel' --launch-skip 2 --launch-count 2 -o "$OUT/attn_full" -f \
  $PY tools/kbench.py > "$OUT/ncu_run.log" 2>&1
if [[ -f "$OUT/attn_full.ncu-rep" ]]; then
  ncu -i "$OUT/attn_full.ncu-rep" --page source --csv --kernel-name attn_fwd_kernel > "$OUT/fwd_src.csv" 2>/dev/null
  ncu -i "$OUT/attn_full.ncu-rep" --page source --csv --kernel-name attn_bwd_kernel > "$OUT/bwd_src.csv" 2>/dev/null
  ncu -i "$OUT/attn_full.ncu-rep" --page raw --csv > "$OUT/attn_raw.csv" 2>/dev/null
fi
cat "$OUT/summary.txt" "$OUT"/kbench_*.txt 2>/dev/null

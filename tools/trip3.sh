#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/t3
mkdir -p "$OUT"
L=$PWD/voicebox-pytorch_b200/lib
echo "== determinism"; timeout 300 python tools/debug_determinism.py > "$OUT/determinism.txt" 2>&1; cat "$OUT/determinism.txt"
VBX_ATTN_FWD_V1=1 timeout 300 python tools/debug_determinism.py > "$OUT/determinism_v1.txt" 2>&1; echo "-- with fwd v1"; cat "$OUT/determinism_v1.txt"
echo "== gemm tests (cluster multicast on by default)"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm or ff1" > "$OUT/tests_gemm.log" 2>&1; tail -3 "$OUT/tests_gemm.log"
echo "== gemm bench cluster2"; timeout 300 python tools/gemm_bench.py > "$OUT/gemm_bench_cluster2.txt" 2>&1; cat "$OUT/gemm_bench_cluster2.txt"
echo "== gemm bench cluster1"; VBX_GEMM_CLUSTER=1 timeout 300 python tools/gemm_bench.py > "$OUT/gemm_bench_cluster1.txt" 2>&1; cat "$OUT/gemm_bench_cluster1.txt"
echo "== attention kbench (warm-ups excluded now): product, fwd v1, stagger variants"
KB_ONLY=attn KB_B=64 KB_ITERS=8 timeout 200 python tools/kbench.py 2>&1 | grep -i "attn" > "$OUT/kbench_product.txt"; cat "$OUT/kbench_product.txt"
VBX_ATTN_FWD_V1=1 KB_ONLY=attn KB_B=64 KB_ITERS=8 timeout 200 python tools/kbench.py 2>&1 | grep -i "attn_fwd" > "$OUT/kbench_fwdv1.txt"; cat "$OUT/kbench_fwdv1.txt"
for v in stag1200 stag2000 stag1600p1; do
  VBX_LIB=$L/libvbx_$v.so timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > "$OUT/tests_$v.log" 2>&1; echo "$v: $(tail -1 $OUT/tests_$v.log)"
  VBX_LIB=$L/libvbx_$v.so KB_ONLY=attn KB_B=64 KB_ITERS=8 timeout 200 python tools/kbench.py 2>&1 | grep -i "attn_fwd" > "$OUT/kbench_$v.txt"; cat "$OUT/kbench_$v.txt"
done
echo "== step profiles: packed vs unpacked"
timeout 300 python tools/step_profile.py > "$OUT/step_profile_packed.txt" 2>&1; head -40 "$OUT/step_profile_packed.txt"
VBX_PACKED=0 timeout 300 python tools/step_profile.py > "$OUT/step_profile_unpacked.txt" 2>&1; head -30 "$OUT/step_profile_unpacked.txt"
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q -rs > "$OUT/tests.log" 2>&1; tail -6 "$OUT/tests.log"

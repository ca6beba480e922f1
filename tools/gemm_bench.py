"""Development tool: the hand-written tcgen05 GEMM (csrc/gemm.cu) against the cuBLASLt kernel torch picks, on the four Linear
shapes of a cfg3 layer (T = 64 * 1040 tokens) and the cfg4 sampling geometry (T = 16 * 2064), plus the fused FF1 + GEGLU epilogue
against cuBLASLt GEMM + vbx_geglu_fwd.  CUDA events, 3 warm-ups, inputs larger than L2 where the real layer's are."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voicebox_pytorch_b200 as vbx  # noqa: E402
from voicebox_pytorch_b200 import ops  # noqa: E402

BF16 = torch.bfloat16


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def bench_ff_bwd(T, out):
    """FF2 dgrad GEMM with the GEGLU-backward epilogue (one launch) against the pair it replaces (cuBLASLt dgrad + vbx_geglu_bwd)."""
    K, Fp = 1024, 2752
    from voicebox_pytorch_b200._lib import call, ptr, stream
    dy = torch.randn(T, K, device='cuda').to(BF16)
    w2 = (torch.randn(K, Fp, device='cuda') / Fp ** 0.5).to(BF16)          # second Linear's weight [out, in]
    w2t = w2.t().contiguous()                                              # [Fp, K]: the dgrad GEMM's B operand, K-major
    h = torch.randn(T, 2 * Fp, device='cuda').to(BF16)
    dh = torch.empty_like(h)
    db = torch.zeros(2 * Fp, device='cuda')

    def pair():
        dg = dy @ w2
        call('vbx_geglu_bwd', ptr(h), ptr(dg), ptr(dh), ptr(db), T, Fp, stream())

    def fused():
        call('vbx_ff2_dgrad_geglu_bwd', ptr(dy), ptr(w2t), ptr(h), ptr(dh), ptr(db), T, Fp, K, stream())

    t_gemm = timeit(lambda: dy @ w2)
    t_pair = timeit(pair)
    t_fused = timeit(fused)
    out[f'ff2_dgrad_geglu_bwd_T{T}'] = dict(cublaslt_dgrad_us=t_gemm, pair_us=t_pair, fused_us=t_fused)
    print(f'ff2 dgrad + geglu bwd T={T:6d}  cuBLASLt dgrad {t_gemm:8.1f} us, + geglu_bwd kernel {t_pair:8.1f} us   fused {t_fused:8.1f} us '
          f' ratio {t_pair / t_fused:5.2f}', flush=True)


out = {}
if os.environ.get('VBX_GEMM_BENCH') == 'bwd':
    bench_ff_bwd(64 * 1040, out)
    sys.exit(0)
for T in (64 * 1040, 16 * 2064):
    for name, K, N, bias in (('to_qkv', 1024, 3072, False), ('to_out', 1024, 1024, False), ('ff2', 2752, 1024, True)):
        a = torch.randn(T, K, device='cuda').to(BF16)
        w = (torch.randn(N, K, device='cuda') / K ** 0.5).to(BF16)
        b = torch.randn(N, device='cuda').to(BF16) if bias else None
        t_lib = timeit(lambda: F.linear(a, w, b))
        t_own = timeit(lambda: ops.gemm_bf16(a, w, b))
        fl = 2.0 * T * K * N
        out[f'{name}_T{T}'] = dict(cublaslt_us=t_lib, tcgen05_us=t_own, cublaslt_tflops=fl / t_lib / 1e6, tcgen05_tflops=fl / t_own / 1e6)
        print(f'{name:8s} T={T:6d} K={K:5d} N={N:5d}  cuBLASLt {t_lib:8.1f} us ({fl / t_lib / 1e6:7.1f} TF/s)   tcgen05 {t_own:8.1f} us '
              f'({fl / t_own / 1e6:7.1f} TF/s)  ratio {t_lib / t_own:5.2f}', flush=True)
    K, Fp = 1024, 2752
    x = torch.randn(T, K, device='cuda').to(BF16)
    w1 = (torch.randn(2 * Fp, K, device='cuda') / K ** 0.5).to(BF16)
    b1 = torch.randn(2 * Fp, device='cuda').to(BF16)
    with torch.no_grad():
        t_lib = timeit(lambda: ops.geglu(F.linear(x, w1, b1)))
        t_gemm = timeit(lambda: F.linear(x, w1, b1))
        t_h = timeit(lambda: ops.ff1_geglu(x, w1, b1, True))
        t_noh = timeit(lambda: ops.ff1_geglu(x, w1, b1, False))
    fl = 2.0 * T * K * 2 * Fp
    out[f'ff1_geglu_T{T}'] = dict(cublaslt_gemm_us=t_gemm, cublaslt_plus_geglu_us=t_lib, fused_with_h_us=t_h, fused_no_h_us=t_noh,
                                  fused_with_h_tflops=fl / t_h / 1e6, fused_no_h_tflops=fl / t_noh / 1e6)
    print(f'ff1+geglu T={T:6d}  cuBLASLt GEMM {t_gemm:8.1f} us, + geglu kernel {t_lib:8.1f} us   fused (h written) {t_h:8.1f} us '
          f'({fl / t_h / 1e6:7.1f} TF/s)   fused (no h) {t_noh:8.1f} us ({fl / t_noh / 1e6:7.1f} TF/s)', flush=True)
    bench_ff_bwd(T, out)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'gemm_bench.json'), 'w'), indent=1)

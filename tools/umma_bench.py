"""Development tool (trace build): cycles per tcgen05.mma as a function of shape, operand source and the number of INDEPENDENT
accumulator tiles the back-to-back MMAs rotate over (1 = every MMA depends on the previous one, as in a K loop)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voicebox_pytorch_b200 as vbx  # noqa
lib = vbx._lib.load()
lib.vbx_debug_umma_bench.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 7
out = torch.zeros(2, dtype=torch.int64, device='cuda')
iters = 200
for (n, a_mn, b_mn, a_t, name) in [(64, 0, 1, 0, 'SS KxMN N=64'), (64, 0, 1, 1, 'TS TMEMxMN N=64'), (64, 1, 1, 0, 'SS MNxMN N=64'),
                                    (128, 0, 0, 0, 'SS KxK N=128'), (32, 0, 0, 0, 'SS KxK N=32'), (16, 0, 0, 0, 'SS KxK N=16')]:
    for nacc in (1, 2, 4):
        if n == 128 and nacc > 2:
            continue
        lib.vbx_debug_umma_bench(out.data_ptr(), n, a_mn, b_mn, a_t, iters, 148, nacc)
        torch.cuda.synchronize()
        o = out.cpu().tolist()
        print(f'{name:18s} nacc={nacc}  issue {o[0] / (8 * iters):6.1f} clk/MMA   retire {o[1] / (8 * iters):6.1f} clk/MMA')

# several issuing warps, each on its own accumulator: per-warp and aggregate rate (is the N <= 64 floor an issue-path limit?)
lib.vbx_debug_umma_bench_mw.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 7
out6 = torch.zeros(6, dtype=torch.int64, device='cuda')
for (n, a_mn, b_mn, a_t, name) in [(64, 0, 1, 0, 'SS KxMN N=64'), (64, 0, 1, 1, 'TS TMEMxMN N=64'), (16, 0, 0, 0, 'SS KxK N=16'),
                                    (128, 0, 0, 0, 'SS KxK N=128')]:
    for nw in (1, 2, 3):
        out6.zero_()
        rc = lib.vbx_debug_umma_bench_mw(out6.data_ptr(), n, a_mn, b_mn, a_t, iters, 148, nw)
        torch.cuda.synchronize()
        o = out6.cpu().tolist()
        retire = max(o[1::2])
        print(f'{name:18s} issuers={nw}  per-warp issue {[round(v / (8 * iters), 1) for v in o[0:2 * nw:2]]} clk/MMA   '
              f'aggregate {retire / (8 * iters * nw):6.1f} clk/MMA (all retired after {retire} clk)  rc={rc}')

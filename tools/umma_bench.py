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

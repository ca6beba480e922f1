"""Development tool: per-tile clock64() timeline of CTA (1,0,0) of the attention kernels.  Needs the trace build:
   VBX_EXP_DEFS=-DVBX_TRACE VBX_EXP_OUT=libvbx_trace.so voicebox-pytorch_b200/csrc/build_exp.sh ; run with VBX_LIB=<that path>."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voicebox_pytorch_b200 as vbx  # noqa: E402
from voicebox_pytorch_b200 import ops  # noqa: E402

lib = vbx._lib.load()
lib.vbx_debug_set_trace.argtypes = [ctypes.c_void_p]
B, H, N = 16, 16, 1040
dev = 'cuda'
qkv = torch.randn(B, N, 3 * H * 64, device=dev).to(torch.bfloat16)
z = torch.zeros(N, 32, device=dev)
cosv, sinv = z.cos().contiguous(), z.sin().contiguous()
gq, gk = torch.ones(H, 1, 64, device=dev), torch.ones(H, 1, 64, device=dev)
qkv1 = qkv.clone().requires_grad_()
do = torch.randn(B, N, H * 64, device=dev).to(torch.bfloat16)
for it in range(3):
    trace = torch.zeros(8 * 16 * 8, dtype=torch.int64, device=dev)
    lib.vbx_debug_set_trace(trace.data_ptr())
    o = ops.attention(qkv1, cosv, sinv, gq.requires_grad_(), gk.requires_grad_(), None, 10., H)
    o.backward(do)
    torch.cuda.synchronize()
t = trace.cpu().view(8, 16, 8)
names = {0: ('bwd MMA A (S^T, dP^T)', ['pre', 'QD_FULL', 'ST_FREE', 'S issued', 'preDS', 'DS_FULL', 'G issued']),
         5: ('bwd MMA B (dV, dK)', ['QD_FULL', 'DS_FULL', 'issued']),
         6: ('bwd MMA C (dQ)', ['pre', 'DS_FULL', 'DQ_FREE', 'issued']),
         1: ('bwd compute t0', ['top', 'bar1', 'ST_FULL', 'math done', 'flushed', 'stored', 'arrived']),
         2: ('bwd producer', ['QD_EMPTY ok']),
         3: ('fwd MMA', ['pre', 'K_FULL', 'S_FREE', 'S issued', 'V_FULL', 'P_FULL', 'PV issued']),
         4: ('fwd softmax t0 (v2)', ['top', 'S_FULL', 'max exchanged', '-', 'exp+P stored', 'P arrived'])}
for role, (nm, pts) in names.items():
    base = int(t[role][t[role] > 0].min()) if (t[role] > 0).any() else 0
    print(f'== {nm}  (clk relative to first stamp of this role; columns: {pts})')
    for tile in range(10):
        row = t[role, tile, :len(pts)]
        if (row > 0).any():
            print(f'  tile {tile}: ' + '  '.join(f'{int(v) - base:7d}' if v > 0 else '      -' for v in row))

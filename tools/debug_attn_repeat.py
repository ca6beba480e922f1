"""Development tool: run-to-run repeatability of the attention forward at size (bitwise), 12 repetitions, two batch sizes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voicebox_pytorch_b200 as vbx  # noqa: E402
from voicebox_pytorch_b200 import ops  # noqa: E402

torch.manual_seed(0)
for (B, H, N) in ((2, 16, 1040), (16, 16, 1040), (4, 16, 2064)):
    qkv = torch.randn(B, N, 3 * H * 64, device='cuda').to(torch.bfloat16)
    z = torch.zeros(N, 32, device='cuda')
    cosv, sinv = z.cos().contiguous(), z.sin().contiguous()
    gq = torch.ones(H, 1, 64, device='cuda')
    with torch.no_grad():
        outs = [ops.attention(qkv, cosv, sinv, gq, gq, None, 10., H) for _ in range(12)]
    bad = [i for i, o in enumerate(outs[1:], 1) if not torch.equal(outs[0], o)]
    worst = max(float((outs[0].float() - o.float()).abs().max()) for o in outs[1:])
    rows = 0
    if bad:
        d = (outs[0].float() - outs[bad[0]].float()).abs().reshape(B, N, H, 64).amax(-1)
        rows = int((d > 1e-3).sum())
    print(f'B={B} N={N}: repeatable={not bad}  differing runs {bad}  max diff {worst:.4f}  (token, head) pairs differing in first bad run: {rows}')

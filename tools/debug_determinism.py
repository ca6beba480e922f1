"""Development tool: is the forward path run-to-run deterministic at size, and where do ConditionalFlowMatcherWrapper.sample
(euler, one interval) and y0 + VoiceBox.forward(y0) differ?  (trip 2: test_full_depth_cfg3_roundtrip_properties part b)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voicebox_pytorch_b200 as vbx  # noqa: E402
from voicebox_pytorch_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
dev = 'cuda'
torch.manual_seed(0)
B, H, N = 2, 16, 1040
qkv = torch.randn(B, N, 3 * H * 64, device=dev).to(BF16)
z = torch.zeros(N, 32, device=dev)
cosv, sinv = z.cos().contiguous(), z.sin().contiguous()
gq = torch.ones(H, 1, 64, device=dev)
with torch.no_grad():
    outs = [ops.attention(qkv, cosv, sinv, gq, gq, None, 10., H) for _ in range(4)]
print('attention fwd bitwise repeatable:', all(torch.equal(outs[0], o) for o in outs[1:]),
      'max diff', max(float((outs[0].float() - o.float()).abs().max()) for o in outs[1:]))

for depth in (2, 24):
    torch.manual_seed(0)
    vb = vbx.VoiceBox(dim=1024, depth=depth, heads=16, condition_on_text=False)
    with torch.no_grad():
        for n, p in vb.named_parameters():
            if 'to_gamma.weight' in n or 'to_beta.weight' in n:
                p.normal_(0, 0.02)
    w = vbx.ConditionalFlowMatcherWrapper(voicebox=vb).cuda()
    w.odeint_kwargs['method'] = 'euler'
    Bm, Nm, D = 2, 1024, 1024
    cond = torch.randn(Bm, Nm, D, device=dev)
    cm = torch.zeros(Bm, Nm, dtype=torch.bool, device=dev)
    cm[:, 300:] = True
    y0 = torch.randn(Bm, Nm, D, device=dev)
    vb.eval()
    with torch.no_grad():
        t0 = torch.zeros((), device=dev)
        f = [vb(y0, times=t0, cond=cond, cond_token_ids=None, cond_mask=cm, cond_drop_prob=0.) for _ in range(3)]
    print(f'depth {depth}: forward bitwise repeatable:', all(torch.equal(f[0], x) for x in f[1:]),
          'max diff', max(float((f[0] - x).abs().max()) for x in f[1:]), 'max |f|', float(f[0].abs().max()))
    real = torch.randn_like
    torch.randn_like = lambda ref_, **kw: y0.clone()
    try:
        out = w.sample(cond=cond, cond_mask=cm, steps=2)
        out2 = w.sample(cond=cond, cond_mask=cm, steps=2)
    finally:
        torch.randn_like = real
    expect = y0 + f[0].to(BF16).float()
    d = (out - expect).abs()
    print(f'depth {depth}: sample repeatable:', torch.equal(out, out2), ' sample vs y0+f: max diff', float(d.max()), 'of', float(expect.abs().max()))
    for b in range(Bm):
        rows = d[b].amax(dim=-1)
        nz = (rows > 1e-3).nonzero().flatten()
        print(f'   batch {b}: rows differing {nz.numel()} of {Nm}', (int(nz.min()), int(nz.max())) if nz.numel() else '')
    # same comparison with times passed as a 1-element device tensor view (what the sampler passes)
    with torch.no_grad():
        tw = torch.tensor([0., 1.], device=dev)
        g = vb(y0, times=tw[0], cond=cond, cond_token_ids=None, cond_mask=cm, cond_drop_prob=0.)
    print(f'depth {depth}: forward(times=0-dim zeros) vs forward(times=tw[0]):', torch.equal(g, f[0]), float((g - f[0]).abs().max()))

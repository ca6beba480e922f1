"""Fixed-grid ODE sampling loop (euler / midpoint) for ConditionalFlowMatcherWrapper.sample.

Restates what torchdiffeq's FixedGridODESolver does at the reference's call site (vp.py:1295; the package is an
un-vendored dependency, oracle/voicebox_oracle.py:odeint_fixed_grid is the checker): the grid is `linspace(0,1,steps)`,
per interval  f0 = f(t0, y);  euler: y += dt f0;  midpoint: y += dt f(t0 + dt/2, y + f0 dt/2).

B200 shape of the loop: the state `y` stays fp32 in HBM; every stage combine is ONE kernel (`vbx_ode_axpy`) that also
rewrites the x-half of the next evaluation's bf16 `to_embed` input and the midpoint time, reading dt from the DEVICE grid
(no host round trip per step; the loop is CUDA-graph capturable).  The conditioning half of that input
(cond * ~cond_mask, vp.py:1035) is constant over the trajectory and is written once.
"""
import torch

from . import ops

METHODS = ('euler', 'midpoint')


def odeint_fixed(vb, *, cond, cond_mask=None, cond_token_ids=None, self_attn_mask=None, steps=3, cond_scale=1.,
                 method='midpoint'):
    from .modules import _voicebox_body, exists
    assert method in METHODS
    B, N, D = cond.shape
    device = cond.device
    y = torch.randn_like(cond)                             # vp.py:1289
    t = torch.linspace(0, 1, steps, device=device)        # vp.py:1290
    fast = (not vb.condition_on_text) and cond_scale == 1. and isinstance(vb.proj_in, torch.nn.Identity)

    if not fast:
        # generic path (text conditioning / classifier-free guidance): the reference's own control flow on the public API
        y = y.float()
        for i in range(steps - 1):
            f0 = vb.forward_with_cond_scale(y, times=t[i], cond_token_ids=cond_token_ids, cond=cond, cond_scale=cond_scale,
                                            cond_mask=cond_mask, self_attn_mask=self_attn_mask)
            if method == 'euler':
                ops.ode_axpy(y, f0.to(torch.bfloat16), t, i, i + 1, half=False, y_out=y)
            else:
                t_mid = torch.empty((1,), device=device, dtype=torch.float32)
                y_mid = ops.ode_axpy(y, f0.to(torch.bfloat16), t, i, i + 1, half=True, y_out=torch.empty_like(y), t_out=t_mid)
                f1 = vb.forward_with_cond_scale(y_mid, times=t_mid[0], cond_token_ids=cond_token_ids, cond=cond,
                                                cond_scale=cond_scale, cond_mask=cond_mask, self_attn_mask=self_attn_mask)
                ops.ode_axpy(y, f1.to(torch.bfloat16), t, i, i + 1, half=False, y_out=y)
        return y

    y = y.float().contiguous()
    if not exists(cond_mask):  # eval default: conditioning fully masked (vp.py:1028-1030)
        cond_mask = torch.ones((B, N), device=device, dtype=torch.bool)
    emb = torch.empty((B, N, 2 * D), device=device, dtype=torch.bfloat16)
    ops.embed_concat(y, cond, cond_mask, out=emb)
    y_mid = torch.empty_like(y) if method == 'midpoint' else None
    t_mid = torch.empty((1,), device=device, dtype=torch.float32)
    for i in range(steps - 1):
        f0 = _voicebox_body(vb, emb, t[i], self_attn_mask)
        if method == 'euler':
            ops.ode_axpy(y, f0, t, i, i + 1, half=False, y_out=y, emb=emb)
        else:
            ops.ode_axpy(y, f0, t, i, i + 1, half=True, y_out=y_mid, emb=emb, t_out=t_mid)
            f1 = _voicebox_body(vb, emb, t_mid[0], self_attn_mask)
            ops.ode_axpy(y, f1, t, i, i + 1, half=False, y_out=y, emb=emb)
    return y

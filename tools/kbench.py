"""Per-kernel micro-benchmark on one B200 (CUDA events, L2-sized-or-larger tensors, 3 warm-ups): achieved GB/s or TFLOP/s of
every vbx kernel at the cfg3 geometry (B=64/16, N'=1040, D=1024, H=16) next to MEASURED_PEAKS.json.  Development tool; the
judged numbers come from bench.py."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voicebox_pytorch_b200 as vbx  # noqa: E402
from voicebox_pytorch_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
except Exception:
    pass
HBM = peaks.get('hbm_gbs', 6650.0)
TF = peaks.get('bf16_tflops', 1590.0)


ITERS = int(os.environ.get('KB_ITERS', 10))
WARM = int(os.environ.get('KB_WARM', 3))


def timeit(fn, iters=None, warm=None):
    iters = ITERS if iters is None else iters
    warm = WARM if warm is None else warm
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, gbytes=None, tflop=None):
    if gbytes is not None:
        print(f'{name:28s} {ms * 1e3:9.1f} us  {gbytes / ms * 1e3:8.0f} GB/s  ({gbytes / ms * 1e3 / HBM * 100:5.1f}% of measured HBM)')
    else:
        print(f'{name:28s} {ms * 1e3:9.1f} us  {tflop / ms * 1e3:8.1f} TFLOP/s ({tflop / ms * 1e3 / TF * 100:5.1f}% of measured bf16)')
    sys.stdout.flush()


def main():
    B = int(os.environ.get('KB_B', 16))
    only = os.environ.get('KB_ONLY', '')  # 'attn': only the attention / qk-rope lines (tools/attn_diagnose.sh)
    torch.manual_seed(0)
    if only != 'attn':
        bench_elementwise(B)
    bench_attention(B)


def bench_elementwise(B):
    N, R, D, H = 1024, 16, 1024, 16
    Np = N + R
    T = B * Np
    dev = 'cuda'
    x = torch.randn(B, Np, D, device=dev)
    br = torch.randn(B, Np, D, device=dev).to(BF16)
    g = torch.rand(B, D, device=dev) + 0.5
    bt = torch.randn(B, D, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: ops.resid_norm(x, br, g, bt))
    report('adarms_fwd', ms, gbytes=T * D * 12 / 1e9)
    x1, br1, g1, bt1 = x.clone().requires_grad_(), br.clone().requires_grad_(), g.clone().requires_grad_(), bt.clone().requires_grad_()
    xo, h = ops.resid_norm(x1, br1, g1, bt1)
    dxo, dh = torch.randn_like(xo), torch.randn_like(h)
    ms = timeit(lambda: torch.autograd.backward([xo, h], [dxo, dh], retain_graph=True))
    report('adarms_bwd (+zeros)', ms, gbytes=T * D * (4 + 2 + 4 + 4 + 2) / 1e9)

    Fp = 2752
    hh = torch.randn(T, 2 * Fp, device=dev).to(BF16)
    with torch.no_grad():
        ms = timeit(lambda: ops.geglu(hh))
    report('geglu_fwd', ms, gbytes=T * Fp * 6 / 1e9)
    h1 = hh.clone().requires_grad_()
    out = ops.geglu(h1)
    d = torch.randn_like(out)
    ms = timeit(lambda: out.backward(d, retain_graph=True))
    report('geglu_bwd', ms, gbytes=T * Fp * 10 / 1e9)

    xe = torch.randn(B, N, D, device=dev).to(BF16)
    w = torch.randn(D, 1, 31, device=dev) * 0.1
    bias = torch.randn(D, device=dev)
    reg = torch.randn(R, D, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: ops.convpos_residual_pack(xe, w, bias, None, reg))
    report('convpos_fwd (bf16->f32)', ms, gbytes=B * N * D * 6 / 1e9)
    xe1, w1, b1, r1 = xe.clone().requires_grad_(), w.clone().requires_grad_(), bias.clone().requires_grad_(), reg.clone().requires_grad_()
    y = ops.convpos_residual_pack(xe1, w1, b1, None, r1)
    dy = torch.randn_like(y)
    ms = timeit(lambda: y.backward(dy, retain_graph=True))
    report('convpos_bwd', ms, gbytes=B * N * D * (2 + 2 + 4 + 2) / 1e9)

    x0, x1_ = torch.randn(B, N, D, device=dev), torch.randn(B, N, D, device=dev)
    times = torch.rand(B, device=dev)
    cm = torch.rand(B, N, device=dev) < 0.8
    ms = timeit(lambda: ops.cfm_embed(x0, x1_, times, cm, 0.))
    report('cfm_embed', ms, gbytes=B * N * D * 12 / 1e9)
    pred = torch.randn(B, N, D, device=dev).to(BF16).requires_grad_()
    loss = ops.masked_mse(pred, cm, x0=x0, x1=x1_, sigma=0.)
    with torch.no_grad():
        ms = timeit(lambda: ops.masked_mse(pred, cm, x0=x0, x1=x1_, sigma=0.))
    report('masked_mse_fwd', ms, gbytes=B * N * D * 10 * 0.8 / 1e9)
    ms = timeit(lambda: loss.backward(retain_graph=True))
    report('masked_mse_bwd', ms, gbytes=B * N * D * (10 * 0.8 + 2) / 1e9)
    y0 = torch.randn(B, N, D, device=dev)
    f = torch.randn(B, N, D, device=dev).to(BF16)
    t = torch.linspace(0, 1, 5, device=dev)
    emb = torch.empty(B, N, 2 * D, device=dev, dtype=BF16)
    yo = torch.empty_like(y0)
    ms = timeit(lambda: ops.ode_axpy(y0, f, t, 0, 1, half=True, y_out=yo, emb=emb))
    report('ode_axpy (+emb)', ms, gbytes=B * N * D * 12 / 1e9)



def bench_attention(B):
    N, R, D, H = 1024, 16, 1024, 16
    Np = N + R
    dev = 'cuda'
    qkv = torch.randn(B, Np, 3 * H * 64, device=dev).to(BF16)
    inv_freq = 1.0 / (50000 ** (torch.arange(0, 64, 2, device=dev).float() / 64))
    pos = torch.cat((torch.full((R,), -10000, device=dev), torch.arange(N, device=dev))).float()
    fr = pos[:, None] * inv_freq[None]
    cosv, sinv = fr.cos().contiguous(), fr.sin().contiguous()
    gq, gk = torch.ones(H, 1, 64, device=dev), torch.ones(H, 1, 64, device=dev)
    iters = ITERS
    qkv1 = qkv.clone().requires_grad_()
    gq1, gk1 = gq.clone().requires_grad_(), gk.clone().requires_grad_()
    do = torch.randn(B, Np, H * 64, device=dev).to(BF16)
    for it in range(iters + WARM):
        if it == WARM:   # warm-ups (module load, attribute calls) stay out of the averages
            torch.cuda.synchronize()
            vbx._lib.profile_start(['vbx_qkrope_fwd', 'vbx_attn_fwd', 'vbx_attn_bwd', 'vbx_qkrope_bwd'])
        o = ops.attention(qkv1, cosv, sinv, gq1, gk1, None, 10., H)
        o.backward(do)
    torch.cuda.synchronize()
    prof = vbx._lib.profile_stop()
    fl = 4 * B * H * Np * Np * 64 / 1e12
    n, tot = prof['vbx_attn_fwd']
    report('attn_fwd', tot / n, tflop=fl)
    n, tot = prof['vbx_attn_bwd']
    report('attn_bwd (+delta)', tot / n, tflop=2.5 * fl)
    n, tot = prof['vbx_qkrope_fwd']
    report('qkrope_fwd', tot / n, gbytes=B * Np * H * 64 * 2 * 4 / 1e9)
    n, tot = prof['vbx_qkrope_bwd']
    report('qkrope_bwd', tot / n, gbytes=B * Np * H * 64 * 2 * (2 + 3 + 2) / 1e9)


if __name__ == '__main__':
    main()

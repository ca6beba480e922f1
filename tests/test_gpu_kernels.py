"""GPU (`-m gpu`): every CUDA kernel, through the C ABI, against the fp32 oracle (oracle/voicebox_oracle.py) on seeded inputs.

Tolerances (stated per test): the kernels compute in fp32 and round ONCE to bf16 on output, so against an fp32 oracle fed the
same bf16-rounded inputs the expected error is one bf16 ulp (2^-8 relative) of the output magnitude; reductions accumulated
with fp32 atomics are compared at 2e-3 relative to the result's max.  Index/mask semantics are exact."""
import math
import os

import pytest
import torch

from oracle import voicebox_oracle as O

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope='module')
def vbx():
    import voicebox_pytorch_b200 as m
    return m


def rel_err(a, b):
    return float((a.float() - b.float()).abs().max()) / max(float(b.float().abs().max()), 1e-20)


def rbf(t):  # round to bf16 and back: what the kernel actually receives
    return t.to(BF16).float()


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('variant', [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13])
def test_umma_descriptors_selftest(vbx, variant):
    """tcgen05.mma + TMEM + (bit2) TMA + (bit3) A operand in TMEM (TS mode): C = A B^T, K = 128, every operand-major combination.  bf16 products are exact in
    fp32 and only the accumulation order differs: tolerance 1e-4 relative."""
    torch.manual_seed(variant)
    A = torch.randn(128, 128, device='cuda').to(BF16)   # logical [M, K]
    B = torch.randn(128, 128, device='cuda').to(BF16)   # logical [N, K]
    a_mem = A.t().contiguous() if variant & 2 else A    # MN-major operand: memory holds [K][M]
    b_mem = B.t().contiguous() if variant & 1 else B
    C = vbx.ops.umma_selftest(a_mem, b_mem, variant)
    ref = A.float() @ B.float().t()
    assert rel_err(C, ref) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,N,D,adaptive', [(2, 216, 128, True), (3, 100, 512, False), (2, 1040, 1024, True), (1, 7, 64, True)])
def test_resid_norm_fwd_bwd(vbx, B, N, D, adaptive):
    torch.manual_seed(0)
    x = torch.randn(B, N, D, device='cuda') * 3
    br = torch.randn(B, N, D, device='cuda').to(BF16)
    if adaptive:
        g = 1 + 0.3 * torch.randn(B, D, device='cuda')
        bt = 0.3 * torch.randn(B, D, device='cuda')
    else:
        g, bt = 1 + 0.3 * torch.randn(D, device='cuda'), None
    xr, brr, gr = x.clone().requires_grad_(), br.float().requires_grad_(), g.clone().requires_grad_()
    btr = bt.clone().requires_grad_() if adaptive else None
    xo_ref = xr + brr
    normed = O.l2_normalize(xo_ref) * math.sqrt(D)
    h_ref = normed * gr[:, None] + btr[:, None] if adaptive else normed * gr
    x1, br1, g1 = x.clone().requires_grad_(), br.clone().requires_grad_(), g.clone().requires_grad_()
    bt1 = bt.clone().requires_grad_() if adaptive else None
    xo, h = vbx.ops.resid_norm(x1, br1, g1, bt1)
    assert torch.equal(xo, xo_ref.detach())                      # fp32 add of the same operands: exact
    assert rel_err(h, h_ref) < 2 ** -8
    dh = torch.randn_like(h_ref)
    dxo = torch.randn_like(xo_ref)
    (h_ref * rbf(dh)).sum().backward(retain_graph=True)
    (xo_ref * dxo).sum().backward()
    torch.autograd.backward([xo, h], [dxo, dh.to(BF16)])
    assert rel_err(x1.grad, xr.grad) < 1e-5                       # fp32 in, fp32 out
    assert rel_err(br1.grad, brr.grad) < 2 ** -8                  # bf16 copy of the same gradient
    assert rel_err(g1.grad, gr.grad) < 2e-3
    if adaptive:
        assert rel_err(bt1.grad, btr.grad) < 2e-3


def test_resid_norm_row_window_and_inplace(vbx):
    """final norm over rows [R, R+N) of a [B, R+N, D] stream (vp.py:476-479); in-place residual update under no_grad."""
    torch.manual_seed(1)
    B, R, N, D = 2, 16, 50, 128
    x = torch.randn(B, R + N, D, device='cuda')
    br = torch.randn(B, R + N, D, device='cuda').to(BF16)
    g = torch.rand(D, device='cuda') + 0.5
    with torch.no_grad():
        _, h = vbx.ops.resid_norm(x, br, g, None, row0=R, rows=N, need_x_out=False)
        ref = O.rms_norm((x + br.float())[:, R:], g)
        assert rel_err(h, ref) < 2 ** -8
        x2 = x.clone()
        xo, _ = vbx.ops.resid_norm(x2, br, g, None, inplace=True)
        assert xo.data_ptr() == x2.data_ptr() and torch.equal(x2, x + br.float())
    x1, br1 = x.clone().requires_grad_(), br.clone().requires_grad_()
    _, h = vbx.ops.resid_norm(x1, br1, g.clone().requires_grad_(), None, row0=R, rows=N)
    h.float().sum().backward()
    assert float(x1.grad[:, :R].abs().max()) == 0 and float(br1.grad[:, :R].abs().max()) == 0
    xr = x.clone().requires_grad_()
    O.rms_norm((xr + br.float())[:, R:], g).sum().backward()
    assert rel_err(x1.grad, xr.grad) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('T,Fp', [(300, 384), (1040 * 2, 2752)])
def test_geglu_fwd_bwd(vbx, T, Fp):
    torch.manual_seed(2)
    h = (torch.randn(T, 2 * Fp, device='cuda') * 2).to(BF16)
    hr = h.float().requires_grad_()
    val, gate = hr.chunk(2, dim=-1)
    ref = torch.nn.functional.gelu(gate) * val               # exact-erf GELU, second half is the gate (vp.py:339-340)
    h1 = h.clone().requires_grad_()
    out = vbx.ops.geglu(h1)
    assert rel_err(out, ref) < 2 ** -8
    d = torch.randn_like(ref).to(BF16)
    ref.backward(d.float())
    out.backward(d)
    assert rel_err(h1.grad, hr.grad) < 2 ** -8


def test_linear_geglu_fused_node_bias_grad(vbx):
    """FF1 + GEGLU as one autograd node (vp.py:345-346): the Linear's bias gradient comes out of the GEGLU backward kernel
    (fp32 column sums of the bf16-rounded dh) instead of a separate reduction; dx / dw are the usual bf16 GEMMs."""
    torch.manual_seed(7)
    T, D, Fp = 520, 128, 192
    x = torch.randn(2, T // 2, D, device='cuda').to(BF16)
    w = (torch.randn(2 * Fp, D, device='cuda') * 0.1).to(BF16)
    b = (torch.randn(2 * Fp, device='cuda') * 0.1).to(BF16)
    xr, wr, br = (t.float().requires_grad_() for t in (x, w, b))
    lin = torch.nn.functional.linear(xr, wr, br)
    hr = lin + (rbf(lin) - lin).detach()                        # the GEMM output is bf16 in both paths (straight-through)
    val, gate = hr.chunk(2, dim=-1)
    ref = torch.nn.functional.gelu(gate) * val
    x1, w1, b1 = (t.clone().requires_grad_() for t in (x, w, b))
    out = vbx.ops.linear_geglu(x1, w1, b1)
    assert rel_err(out, ref) < 2 ** -7
    d = torch.randn_like(ref).to(BF16)
    ref.backward(d.float())
    out.backward(d)
    assert rel_err(b1.grad, br.grad) < 2e-2      # bf16 storage of the result; sums over 520 rows of bf16-rounded terms
    assert rel_err(w1.grad, wr.grad) < 2e-2
    assert rel_err(x1.grad, xr.grad) < 2e-2


def test_clip_coefficient_through_fused_adam_grad_scale():
    """bench.py applies clip_grad_norm_(0.5) (trainer.py:274-275) through the fused Adam kernel's grad_scale input: identical to
    scaling the gradients first."""
    torch.manual_seed(8)
    p1 = torch.randn(1000, device='cuda', requires_grad=True)
    p2 = p1.detach().clone().requires_grad_()
    g = torch.randn(1000, device='cuda') * 3
    o1 = torch.optim.Adam([p1], lr=3e-4, betas=(0.9, 0.99), fused=True)
    o2 = torch.optim.Adam([p2], lr=3e-4, betas=(0.9, 0.99), fused=True)
    for _ in range(3):
        norm = g.norm()
        p1.grad = g * torch.clamp(0.5 / (norm + 1e-6), max=1.0)
        o1.step()
        p2.grad = g.clone()
        o2.grad_scale = torch.clamp((norm + 1e-6) / 0.5, min=1.0)
        o2.found_inf = torch.zeros((), device='cuda')
        o2.step()
    assert torch.allclose(p1, p2, rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,N,C,K,R,masked', [(2, 200, 128, 31, 16, False), (3, 100, 128, 31, 0, True), (2, 300, 64, 7, 4, True),
                                              (1, 1024, 256, 31, 16, False)])
def test_convpos_fwd_bwd(vbx, B, N, C, K, R, masked):
    torch.manual_seed(3)
    x = torch.randn(B, N, C, device='cuda').to(BF16)
    w = torch.randn(C, 1, K, device='cuda') * 0.2
    bias = torch.randn(C, device='cuda') * 0.1
    reg = torch.randn(R, C, device='cuda') if R else None
    mask = None
    if masked:
        mask = torch.ones(B, N, dtype=torch.bool, device='cuda')
        for b in range(B):
            mask[b, N - 7 * (b + 1):] = False
    xr, wr, br = x.float().requires_grad_(), w.clone().requires_grad_(), bias.clone().requires_grad_()
    regr = reg.clone().requires_grad_() if R else None
    y_ref = O.conv_pos_embed(xr, wr, br, mask) + xr
    if R:
        y_ref = torch.cat((regr[None].expand(B, -1, -1), y_ref), dim=1)
    x1, w1, b1 = x.clone().requires_grad_(), w.clone().requires_grad_(), bias.clone().requires_grad_()
    reg1 = reg.clone().requires_grad_() if R else None
    y = vbx.ops.convpos_residual_pack(x1, w1, b1, mask, reg1)
    assert y.dtype == torch.float32 and rel_err(y, y_ref) < 1e-5      # fp32 accumulate of the same bf16 inputs
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    y.backward(dy)
    # backward uses the bf16-saved pre-activation (one bf16 ulp on gelu'(u)) -> 2^-7
    assert rel_err(x1.grad, xr.grad) < 2 ** -7
    assert rel_err(w1.grad, wr.grad) < 5e-3
    assert rel_err(b1.grad, br.grad) < 5e-3
    if R:
        assert rel_err(reg1.grad, regr.grad) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
def _ste_bf16(t):
    """forward: round to bf16 (the tensor-core operand precision, = the reference's autocast cast before its einsum);
    backward: identity."""
    return t + (rbf(t) - t).detach()


def _attn_reference(qkv, H, gq, gk, pos, inv_freq, scale, key_mask):
    B, N, _ = qkv.shape
    q, k, v = (t.reshape(B, N, H, 64).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    if gq is not None:
        q, k = O.multihead_rms_norm(q, gq), O.multihead_rms_norm(k, gk)
    ang = O.rotary_angles(pos, inv_freq)
    q, k = O.apply_rotary(ang, q), O.apply_rotary(ang, k)
    o = O.attend(_ste_bf16(q), _ste_bf16(k), v, scale, key_mask)
    return o.transpose(1, 2).reshape(B, N, H * 64), q, k


@pytest.mark.parametrize('B,H,N,R,qk_norm,masked', [
    (1, 2, 130, 0, False, False),      # 1 full tile + a 2-key tail (14 padded columns inside the narrowest GEMM)
    (1, 2, 255, 0, True, False),       # 127-key tail
    (2, 2, 161, 0, False, True),       # 33-key tail + key-padding mask
    (1, 2, 257, 0, False, False),      # 2 full tiles + a 1-key tail
    (1, 16, 2064, 16, True, False),    # cfg4 geometry (seq 2048 + 16 registers): 16 full tiles + 16
    (2, 2, 80, 16, True, False),       # single partial tile
    (2, 4, 216, 16, True, False),      # 1 full tile + 88-key tail (golden geometry)
    (1, 2, 128, 0, False, False),      # exactly one tile, no qk-norm (scale 1/8)
    (2, 2, 300, 0, True, True),        # key-padding mask (DurationPredictor path) + tail
    (2, 2, 300, 0, False, True),       # multi-tile, no qk-norm (smooth softmax: tight check of the dQ/dK/dV GEMMs)
    (1, 16, 1040, 16, True, False),    # cfg3 geometry: 8 full tiles + 16
    (2, 3, 100, 0, True, False),       # 2H = 6 does not divide a block's vector slots: the strided token order of the rope kernels
    (1, 1, 70, 0, True, False),        # one head: a rope block covers 16 / 32 tokens per pass (group > 8 tokens, single pass)
    (2, 8, 150, 16, True, False),      # 2H = 16: two / four tokens per pass
    (1, 32, 100, 0, True, False),      # 2H = 64: grouped forward, strided backward
    (3, 16, 131, 0, True, False),      # H = 16: the staged rope backward; 393 tokens = 98 stages of four + a one-token tail
])
def test_attention_fwd_bwd(vbx, B, H, N, R, qk_norm, masked):
    """qk-norm + rotary prologue + tcgen05 flash attention vs the fp32 oracle (vp.py:320-332, attend.py:119-137).
    Tolerance: q^,k^ are rounded to bf16 (as the reference's autocast does before its einsum) and P / dS to bf16 before
    the second GEMMs.  With |logit| up to ~10*64 at scale 10 one bf16 ulp of q^.k^ moves a logit by ~2, so the fp32 oracle is
    evaluated on the SAME bf16-rounded q^,k^ (straight-through rounding: that isolates the kernel from the operand rounding
    the reference also performs); outputs are then compared at 3e-2 of the output max and gradients at 5e-2."""
    torch.manual_seed(4)
    dev = 'cuda'
    qkv = torch.randn(B, N, 3 * H * 64, device=dev).to(BF16)
    gq = (1 + 0.2 * torch.randn(H, 1, 64, device=dev)) if qk_norm else None
    gk = (1 + 0.2 * torch.randn(H, 1, 64, device=dev)) if qk_norm else None
    scale = 10. if qk_norm else 64 ** -0.5
    if not qk_norm:
        qkv = (qkv.float() * 0.5).to(BF16)
    inv_freq = 1.0 / (50000 ** (torch.arange(0, 64, 2, device=dev).float() / 64))
    pos = torch.cat((torch.full((R,), -10000, device=dev, dtype=torch.long), torch.arange(N - R, device=dev)))
    fr = pos.float()[:, None] * inv_freq[None, :]
    cosv, sinv = fr.cos().contiguous(), fr.sin().contiguous()
    key_mask = None
    if masked:
        key_mask = torch.ones(B, N, dtype=torch.bool, device=dev)
        key_mask[0, N - 37:] = False
        key_mask[1, 5:9] = False

    qkv1 = qkv.clone().requires_grad_()
    gq1 = gq.clone().requires_grad_() if qk_norm else None
    gk1 = gk.clone().requires_grad_() if qk_norm else None
    o = vbx.ops.attention(qkv1, cosv, sinv, gq1, gk1, key_mask, scale, H)

    # prologue alone: q^, k^ against the oracle (one bf16 rounding)
    qkvr = qkv.float().requires_grad_()
    gqr = gq.clone().requires_grad_() if qk_norm else None
    gkr = gk.clone().requires_grad_() if qk_norm else None
    o_ref, q_ref, k_ref = _attn_reference(qkvr, H, gqr, gkr, pos, inv_freq, scale, key_mask)
    assert rel_err(o, o_ref) < 3e-2
    do = torch.randn_like(o_ref)
    o_ref.backward(rbf(do))
    o.backward(do.to(BF16))
    assert rel_err(qkv1.grad, qkvr.grad) < 5e-2
    if qk_norm:
        assert rel_err(gq1.grad, gqr.grad) < 5e-2
        assert rel_err(gk1.grad, gkr.grad) < 5e-2


def test_attention_fully_masked_keys_uniform(vbx):
    """attend.py:127-128 semantics: -finfo.max fill => a row whose keys are all masked attends uniformly."""
    torch.manual_seed(5)
    B, H, N = 1, 2, 96
    qkv = torch.randn(B, N, 3 * H * 64, device='cuda').to(BF16)
    z = torch.zeros(N, 32, device='cuda')
    mask = torch.zeros(B, N, dtype=torch.bool, device='cuda')
    with torch.no_grad():
        o = vbx.ops.attention(qkv, z.cos().contiguous(), z.sin().contiguous(), None, None, mask, 0.125, H)
    v = qkv.float().chunk(3, dim=-1)[2]
    assert rel_err(o, v.mean(dim=1, keepdim=True).expand_as(v)) < 2e-2


# ---------------------------------------------------------------------------------------------------------------------
def test_cfm_embed_mse_and_ode(vbx):
    torch.manual_seed(6)
    B, N, D, sigma = 3, 130, 128, 0.1
    x0, x1 = torch.randn(B, N, D, device='cuda'), torch.randn(B, N, D, device='cuda')
    times = torch.rand(B, device='cuda')
    cm = torch.rand(B, N, device='cuda') < 0.7
    w, flow = O.cfm_interpolate(x0, x1, times, sigma)
    emb = vbx.ops.cfm_embed(x0, x1, times, cm, sigma)
    assert rel_err(emb[..., :D], w) < 2 ** -8 and rel_err(emb[..., D:], flow * ~cm[..., None]) < 2 ** -8
    assert float(emb[..., D:][cm].abs().max()) == 0.
    emb2 = vbx.ops.embed_concat(w, flow, cm)
    assert torch.equal(emb2[..., D:], (flow * ~cm[..., None]).to(BF16)) and torch.equal(emb2[..., :D], w.to(BF16))

    pred = torch.randn(B, N, D, device='cuda').to(BF16)
    pr = pred.float().requires_grad_()
    ref = O.masked_mse(pr, flow, cm)
    ref.backward()
    for kw in (dict(target=flow), dict(x0=x0, x1=x1, sigma=sigma)):
        p1 = pred.clone().requires_grad_()
        loss = vbx.ops.masked_mse(p1, cm, **kw)
        assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())   # fp32 math on identical inputs
        (loss * 0.5).backward()                                           # non-unit upstream gradient
        assert rel_err(p1.grad, 0.5 * pr.grad) < 2 ** -8

    y = torch.randn(B, N, D, device='cuda')
    f = torch.randn(B, N, D, device='cuda').to(BF16)
    t = torch.linspace(0, 1, 5, device='cuda')
    emb = torch.zeros(B, N, 2 * D, device='cuda', dtype=BF16)
    tmid = torch.zeros(1, device='cuda')
    ym = vbx.ops.ode_axpy(y, f, t, 1, 2, half=True, y_out=torch.empty_like(y), emb=emb, t_out=tmid)
    dt = t[2] - t[1]
    assert torch.allclose(ym, y + f.float() * (0.5 * dt), atol=1e-6)
    assert torch.equal(emb[..., :D], ym.to(BF16)) and float(emb[..., D:].abs().max()) == 0
    assert float(tmid[0]) == float(t[1] + 0.5 * dt)
    y2 = y.clone()
    vbx.ops.ode_axpy(y2, f, t, 1, 2, half=False, y_out=y2)
    assert torch.allclose(y2, y + dt * f.float(), atol=1e-6)


def test_argument_errors_are_reported_not_launched(vbx):
    x = torch.randn(2, 4, 12, device='cuda')  # D % 8 != 0
    with pytest.raises(RuntimeError, match='invalid size'):
        vbx.ops.resid_norm(x, None, torch.ones(12, device='cuda'))
    with pytest.raises(RuntimeError, match='not supported'):
        vbx.ops.resid_norm(torch.randn(1, 2, 4096, device='cuda'), None, torch.ones(4096, device='cuda'))


@pytest.mark.parametrize('clip', [None, 0.5])
def test_flat_adam_matches_torch_adam(vbx, clip):
    """vbx_adam_step (fused clip + Adam over the flat buffers, trainer.py:274-278) against clip_grad_norm_ + torch.optim.Adam
    with the reference's hyper-parameters (optimizer.py:13-14), 4 steps, a bucket whose size is not a multiple of 4.
    fp32 in / fp32 out: 1e-5 relative (fused multiply-adds and one reciprocal differ from ATen's op sequence)."""
    from voicebox_pytorch_b200.dist import FlatGradBucket

    def make():
        torch.manual_seed(1)
        return torch.nn.Sequential(torch.nn.Linear(37, 30), torch.nn.Linear(30, 3)).cuda()

    a, b = make(), make()
    ref = torch.optim.Adam(a.parameters(), lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    bucket = FlatGradBucket(b)
    opt = vbx.FlatAdam(bucket, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=clip, bf16_shadow=True)
    # parameter sizes 3, 90, 30, 1110 (none a multiple of 4): every parameter still starts on a 16-byte boundary
    assert [o % 4 for o in bucket.offsets] == [0, 0, 0, 0] and bucket.offsets[1] == 4
    assert all(p.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0 for p in b.parameters())
    torch.manual_seed(2)
    for _ in range(4):
        x = torch.randn(16, 37, device='cuda')
        ref.zero_grad()
        a(x).square().sum().backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(a.parameters(), clip)
        ref.step()
        opt.zero_grad()
        b(x).square().sum().backward()
        bucket.finish()
        opt.step()
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7)
    assert torch.equal(opt.flat_p_bf16, opt.flat_p.to(BF16))            # the bf16 operand copy of the next forward
    assert all(p._version > 0 for p in b.parameters())                   # raw-pointer update is visible to version checks
    before = opt.flat_p.clone()
    opt.found_inf = torch.ones(1, device='cuda')                         # skipped step: nothing may change
    opt.step()
    assert torch.equal(opt.flat_p, before)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,K,N,with_bias', [(300, 128, 256, True), (128, 64, 64, False), (1040 * 2, 1024, 3072, False),
                                             (777, 2752, 1024, True), (4096, 1024, 1000, True)])
def test_tcgen05_gemm_plain_vs_fp32(vbx, M, K, N, with_bias):
    """csrc/gemm.cu, mode PLAIN: C = A W^T + b with fp32 accumulation in tensor memory and ONE rounding to bf16 -- against the
    fp32 product of the same bf16 operands: 2^-8 relative to the output max plus the fp32 accumulation-order noise.  Covers
    row tails (M % 128), column tails (N % 256, N % 64), K tails (K % 64) and a single-tile problem."""
    torch.manual_seed(11)
    a = torch.randn(M, K, device='cuda').to(BF16)
    w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).to(BF16)
    b = torch.randn(N, device='cuda').to(BF16) if with_bias else None
    c = vbx.ops.gemm_bf16(a, w, b)
    ref = a.float() @ w.float().t() + (b.float() if with_bias else 0)
    assert c.shape == (M, N) and c.dtype == BF16
    assert rel_err(c, ref) < 2 ** -7


@pytest.mark.parametrize('T,K,Fp', [(300, 128, 192), (520, 128, 384), (1040 * 2, 1024, 2752), (129, 64, 32)])
def test_tcgen05_ff1_geglu_fused_epilogue(vbx, T, K, Fp):
    """csrc/gemm.cu, GEGLU modes (vp.py:337-346): h = x W1^T + b1 rounded to bf16, g = gelu_erf(gate) * value on the ROUNDED h --
    the fused kernel must reproduce cuBLASLt GEMM + vbx_geglu_fwd: h to one bf16 ulp of the accumulation-order noise, g = the
    GEGLU of ITS OWN h (checked by re-running the stand-alone GEGLU kernel on it), and the no-h inference variant the same g."""
    torch.manual_seed(12)
    x = torch.randn(T, K, device='cuda').to(BF16)
    w1 = (torch.randn(2 * Fp, K, device='cuda') / math.sqrt(K)).to(BF16)
    b1 = (0.1 * torch.randn(2 * Fp, device='cuda')).to(BF16)
    h, g = vbx.ops.ff1_geglu(x, w1, b1, True)
    h_ref = x.float() @ w1.float().t() + b1.float()
    assert rel_err(h, h_ref) < 2 ** -7
    with torch.no_grad():
        g_from_h = vbx.ops.geglu(h)
    assert rel_err(g, g_from_h) <= 2 ** -8          # same formula on the same rounded h (at most an ulp from FMA contraction)
    _, g2 = vbx.ops.ff1_geglu(x, w1, b1, False)
    assert torch.equal(g2, g)
    val, gate = rbf(h_ref).chunk(2, dim=-1)
    assert rel_err(g, torch.nn.functional.gelu(gate) * val) < 2 ** -6


@pytest.mark.parametrize('T,K,Fp', [(300, 128, 192), (129, 64, 64), (100, 64, 128), (1040 * 2, 1024, 2752), (520, 256, 320)])
def test_tcgen05_ff2_dgrad_with_geglu_backward_epilogue(vbx, T, K, Fp):
    """csrc/gemm.cu:gemm_geglu_bwd_kernel (vp.py:337-348, backward): dg = dy W2 on tcgen05 with dh = GEGLU'(h) * dg and the FF1
    bias gradient formed in the epilogue -- against cuBLASLt dgrad GEMM + vbx_geglu_bwd (the pair it replaces) and against the
    fp32 formula.  dh: 2^-6 of its max (dg carries one bf16 rounding, as in the reference's autocast backward; a differently
    ordered fp32 accumulation can move that rounding by an ulp); db1 (sums of bf16-rounded dh over T rows): 1e-2 of its max."""
    torch.manual_seed(13)
    dy = torch.randn(T, K, device='cuda').to(BF16)
    w2 = (torch.randn(K, Fp, device='cuda') / math.sqrt(Fp)).to(BF16)          # second Linear's weight [D, Fp]
    h = (torch.randn(T, 2 * Fp, device='cuda') * 1.5).to(BF16)
    w2t = w2.t().contiguous()
    dh = torch.empty_like(h)
    db = torch.zeros(2 * Fp, device='cuda')
    vbx._lib.call('vbx_ff2_dgrad_geglu_bwd', dy.data_ptr(), w2t.data_ptr(), h.data_ptr(), dh.data_ptr(), db.data_ptr(), T, Fp, K,
                  vbx._lib.stream())
    # the pair it replaces
    dg = dy @ w2                                                                  # bf16 [T, Fp]
    dh_pair = torch.empty_like(h)
    db_pair = torch.zeros(2 * Fp, device='cuda')
    vbx._lib.call('vbx_geglu_bwd', h.data_ptr(), dg.data_ptr(), dh_pair.data_ptr(), db_pair.data_ptr(), T, Fp, vbx._lib.stream())
    assert rel_err(dh, dh_pair) < 2 ** -6
    assert rel_err(db, db_pair) < 1e-2
    # the formula in fp32
    hr = h.float().requires_grad_()
    val, gate = hr.chunk(2, dim=-1)
    (torch.nn.functional.gelu(gate) * val).backward(rbf(dy.float() @ w2.float()))
    assert rel_err(dh, hr.grad) < 2 ** -6
    assert rel_err(db, rbf(hr.grad).sum(dim=0)) < 1e-2

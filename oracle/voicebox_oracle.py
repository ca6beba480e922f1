"""CPU/GPU-agnostic fp32 ORACLE for the Voicebox conditional-flow-matching hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
leg may import this file, and only as the checker / timed CPU baseline -- never from the product package
(`voicebox_pytorch_b200`), which must fail loudly if its CUDA library is missing.

It is a functional restatement (plain torch ops on a flat {name: tensor} state dict, no nn.Module) of what
lucidrains/voicebox-pytorch @ v0.5.0 computes on the path BASELINE.json's north_star names.  Every function
cites the reference lines it follows (vp.py = voicebox_pytorch/voicebox_pytorch.py).  State-dict key names are the
reference's own, so a reference checkpoint feeds it unchanged.

Pinning: the reference ships no tests / golden vectors.  This restatement is pinned against the reference ITSELF,
imported read-only in the build container (oracle/ref_import.py), by tests/test_oracle_vs_reference.py (runs only
where /root/reference exists) and, everywhere else, against tests/golden/*.npz produced from the reference by
tests/golden/make_golden.py.  `odeint_fixed_grid` restates torchdiffeq's fixed-grid euler / midpoint solvers
(un-vendored, un-pinned dependency: setup.py:27, call site vp.py:1295) -- PARITY UNPINNED for that function.
"""
import math

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------
# mask / index generation (bit-exact contract)
# ---------------------------------------------------------------------------------------------------


def prob_mask_like(shape, prob, device):
    """vp.py:68-74.  p in {0,1} short-circuits without touching the RNG."""
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


def mask_from_start_end(seq_len, start, end):
    """vp.py:121-135.  start/end are FLOAT tensors; truncation to int64 happens here."""
    pos = torch.arange(seq_len, device=start.device, dtype=torch.long)
    pos = pos.reshape(*((1,) * start.ndim), seq_len)
    return (pos >= start[..., None].long()) & (pos < end[..., None].long())


def mask_from_frac_lengths(seq_len, frac_lengths, rand=None):
    """vp.py:137-150.  fp32 op order matters: lengths=(frac*N).long(); start=(max_start*rand).clamp(0) stays float;
    end=start+lengths is a float add; both truncated afterwards.  `rand` may be injected (tests)."""
    lengths = (frac_lengths * seq_len).long()
    max_start = seq_len - lengths
    if rand is None:
        rand = torch.zeros_like(frac_lengths).float().uniform_(0, 1)
    start = (max_start * rand).clamp(min=0)
    end = start + lengths
    return mask_from_start_end(seq_len, start, end)


# ---------------------------------------------------------------------------------------------------
# embeddings
# ---------------------------------------------------------------------------------------------------


def learned_sinusoidal(times, weights):
    """vp.py:154-167: cat(sin(t*w*2pi), cos(t*w*2pi))."""
    f = times[:, None] * weights[None, :] * 2 * math.pi
    return torch.cat((f.sin(), f.cos()), dim=-1)


def time_embedding(sd, times, prefix=''):
    """vp.py:916-920, 1082: SiLU(Linear(sinusoidal(t)))."""
    s = learned_sinusoidal(times, sd[prefix + 'sinu_pos_emb.0.weights'])
    return F.silu(F.linear(s, sd[prefix + 'sinu_pos_emb.1.weight'], sd[prefix + 'sinu_pos_emb.1.bias']))


def rotary_angles(positions, inv_freq):
    """vp.py:172-191: outer(pos, inv_freq) duplicated along the last dim (half-split layout), fp32."""
    fr = positions.to(inv_freq.dtype)[:, None] * inv_freq[None, :]
    return torch.cat((fr, fr), dim=-1)


def apply_rotary(angles, t):
    """vp.py:193-199: t*cos + rotate_half(t)*sin with rotate_half([a,b]) = [-b,a]."""
    a, b = t.chunk(2, dim=-1)
    return t * angles.cos() + torch.cat((-b, a), dim=-1) * angles.sin()


def conv_pos_embed(x, weight, bias, mask=None):
    """vp.py:203-233: mask -> depthwise conv1d (k odd, pad k//2, groups=dim) -> exact GELU -> mask.
    Caller adds the residual (vp.py:826, 1080)."""
    if mask is not None:
        x = x.masked_fill(~mask[..., None], 0.)
    k = weight.shape[-1]
    y = F.conv1d(x.transpose(1, 2), weight, bias, padding=k // 2, groups=weight.shape[0])
    y = F.gelu(y).transpose(1, 2)
    if mask is not None:
        y = y.masked_fill(~mask[..., None], 0.)
    return y


# ---------------------------------------------------------------------------------------------------
# norms
# ---------------------------------------------------------------------------------------------------


def l2_normalize(x):
    """F.normalize(x, dim=-1): x / max(||x||_2, 1e-12)  (vp.py:247, 271, 287)."""
    return x / x.norm(dim=-1, keepdim=True).clamp(min=1e-12)


def rms_norm(x, gamma):
    """vp.py:237-247."""
    return l2_normalize(x) * (x.shape[-1] ** 0.5) * gamma


def adaptive_rms_norm(x, cond, wg, bg, wb, bb):
    """vp.py:249-276: gamma, beta = Linear(cond); normed*gamma[:,None] + beta[:,None]."""
    normed = l2_normalize(x) * (x.shape[-1] ** 0.5)
    gamma = F.linear(cond, wg, bg)[:, None, :]
    beta = F.linear(cond, wb, bb)[:, None, :]
    return normed * gamma + beta


def multihead_rms_norm(x, gamma):
    """vp.py:280-287: x (b,h,n,d), gamma (h,1,d)."""
    return l2_normalize(x) * gamma * (x.shape[-1] ** 0.5)


# ---------------------------------------------------------------------------------------------------
# attention / feed-forward / trunk
# ---------------------------------------------------------------------------------------------------


def attend(q, k, v, scale, key_mask=None):
    """attend.py:119-137 (math path; dropout p=0): softmax(q k^T * scale, masked keys -> -finfo.max) v."""
    sim = torch.einsum('bhid,bhjd->bhij', q, k) * scale
    if key_mask is not None:
        sim = sim.masked_fill(~key_mask[:, None, None, :], -torch.finfo(sim.dtype).max)
    return torch.einsum('bhij,bhjd->bhid', sim.softmax(dim=-1), v)


def attention_block(sd, p, x, heads, qk_norm, key_mask, angles):
    """vp.py:289-333: to_qkv -> split heads -> [qk-norm] -> rotary(q,k) -> Attend -> merge -> to_out."""
    b, n, _ = x.shape
    qkv = F.linear(x, sd[p + 'to_qkv.weight'])
    q, k, v = (t.reshape(b, n, heads, -1).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    dh = q.shape[-1]
    scale = dh ** -0.5
    if qk_norm:
        q = multihead_rms_norm(q, sd[p + 'q_norm.gamma'])
        k = multihead_rms_norm(k, sd[p + 'k_norm.gamma'])
        scale = 10.  # vp.py:304 qk_norm_scale
    q, k = apply_rotary(angles, q), apply_rotary(angles, k)
    o = attend(q, k, v, scale, key_mask)
    return F.linear(o.transpose(1, 2).reshape(b, n, heads * dh), sd[p + 'to_out.weight'])


def feed_forward(sd, p, x):
    """vp.py:337-349: Linear(D,2F) -> chunk -> gelu(second half) * first half -> Linear(F,D)."""
    h = F.linear(x, sd[p + '0.weight'], sd[p + '0.bias'])
    val, gate = h.chunk(2, dim=-1)
    return F.linear(F.gelu(gate) * val, sd[p + '3.weight'], sd[p + '3.bias'])


def transformer(sd, x, *, prefix, depth, heads, qk_norm=False, adaptive=False, cond=None, mask=None,
                num_register_tokens=0, skip_connect_scale=2 ** -0.5):
    """vp.py:412-479.  Register tokens packed on the left with rotary position -10000; per layer
    [unet skip-combine] norm -> attn -> +res -> norm -> ff -> +res; unpack; final RMSNorm."""
    b, n, _ = x.shape
    r = num_register_tokens
    if r > 0:
        reg = sd[prefix + 'register_tokens']
        x = torch.cat((reg[None].expand(b, -1, -1), x), dim=1)
        if mask is not None:
            mask = F.pad(mask, (r, 0), value=True)
        positions = torch.cat((torch.full((r,), -10000, device=x.device, dtype=torch.long),
                               torch.arange(n, device=x.device, dtype=torch.long)))
    else:
        positions = torch.arange(n, device=x.device)
    angles = rotary_angles(positions, sd[prefix + 'rotary_emb.inv_freq'])

    def norm(p, t):
        if adaptive:
            return adaptive_rms_norm(t, cond, sd[p + 'to_gamma.weight'], sd[p + 'to_gamma.bias'],
                                     sd[p + 'to_beta.weight'], sd[p + 'to_beta.bias'])
        return rms_norm(t, sd[p + 'gamma'])

    skips = []
    for i in range(depth):
        lp = f'{prefix}layers.{i}.'
        if (lp + '0.weight') in sd:  # U-Net skip combiner (vp.py:458-463)
            s = skips.pop() * skip_connect_scale
            x = F.linear(torch.cat((x, s), dim=-1), sd[lp + '0.weight'], sd[lp + '0.bias'])
        else:
            skips.append(x)
        x = attention_block(sd, lp + '3.', norm(lp + '2.', x), heads, qk_norm, mask, angles) + x
        x = feed_forward(sd, lp + '5.', norm(lp + '4.', x)) + x
    if r > 0:
        x = x[:, r:]
    return rms_norm(x, sd[prefix + 'final_norm.gamma'])


# ---------------------------------------------------------------------------------------------------
# VoiceBox / DurationPredictor / CFM wrapper
# ---------------------------------------------------------------------------------------------------


def interpolate_1d(t, length):
    """vp.py:89-107 ('bilinear' over a (n,1) image == linear along n, align_corners=False)."""
    return F.interpolate(t.float()[..., None], (length, 1), mode='bilinear')[..., 0].to(t.dtype)


def masked_mse(pred, target, loss_mask):
    """vp.py:1099-1115: per-frame mean over d, zero unmasked, per-sample sum/count (clamped 1e-5), batch mean."""
    if loss_mask is None:
        return F.mse_loss(pred, target)
    per = F.mse_loss(pred, target, reduction='none').mean(dim=-1).masked_fill(~loss_mask, 0.)
    return (per.sum(dim=-1) / loss_mask.sum(dim=-1).clamp(min=1e-5)).mean()


def voicebox_forward(sd, cfg, x, *, times, cond=None, cond_mask=None, target=None, self_attn_mask=None,
                     cond_token_ids=None, cond_drop_prob=0., training=False, prefix=''):
    """vp.py:987-1115.  cfg: dict(depth, heads, num_register_tokens, qk_norm, condition_on_text, frac_lengths_mask).
    `cond = default(cond, target)` quirk preserved (vp.py:1003).  RNG draws (training, cond_mask None):
    uniform_(frac lo,hi) then uniform_(0,1) inside mask_from_frac_lengths, in this order (vp.py:1025, 146)."""
    if cond is None:
        cond = target
    b, n, _ = cond.shape
    if times.ndim == 0 or (times.ndim == 1 and times.shape[0] == 1):
        times = times.reshape(-1).expand(b)
    if cond_mask is None:
        if training:
            lo, hi = cfg.get('frac_lengths_mask', (0.7, 1.))
            frac = torch.zeros((b,), device=x.device).float().uniform_(lo, hi)
            cond_mask = mask_from_frac_lengths(n, frac)
        else:
            cond_mask = torch.ones((b, n), device=x.device, dtype=torch.bool)
    cond = cond * ~cond_mask[..., None]

    cond_ids = cond_token_ids
    if cond_drop_prob > 0.:
        drop = prob_mask_like((b,), cond_drop_prob, x.device)
        cond = torch.where(drop[:, None, None], sd[prefix + 'null_cond'], cond)
        cond_ids = torch.where(drop[:, None], cfg['num_cond_tokens'], cond_token_ids)

    parts = [x]
    if cfg.get('condition_on_text', False):
        emb = F.embedding(cond_ids, sd[prefix + 'to_cond_emb.weight'])
        if emb.shape[-2] != n:
            emb = interpolate_1d(emb.transpose(1, 2), n).transpose(1, 2)
            if self_attn_mask is not None:
                self_attn_mask = interpolate_1d(self_attn_mask, n)
        parts.append(emb)
    parts.append(cond)

    h = F.linear(torch.cat(parts, dim=-1), sd[prefix + 'to_embed.weight'], sd[prefix + 'to_embed.bias'])
    h = conv_pos_embed(h, sd[prefix + 'conv_embed.dw_conv1d.0.weight'], sd[prefix + 'conv_embed.dw_conv1d.0.bias'],
                       self_attn_mask) + h
    temb = time_embedding(sd, times, prefix)
    h = transformer(sd, h, prefix=prefix + 'transformer.', depth=cfg['depth'], heads=cfg['heads'],
                    qk_norm=cfg.get('qk_norm', True), adaptive=True, cond=temb, mask=self_attn_mask,
                    num_register_tokens=cfg.get('num_register_tokens', 16))
    pred = F.linear(h, sd[prefix + 'to_pred.weight'])
    if target is None:
        return pred
    loss_mask = cond_mask if self_attn_mask is None else (cond_mask & self_attn_mask)
    return masked_mse(pred, target, loss_mask)


def voicebox_forward_with_cond_scale(sd, cfg, x, *, cond_scale=1., **kw):
    """vp.py:972-985 (classifier-free guidance): logits at cond_drop_prob = 0; if cond_scale != 1 a second pass at
    cond_drop_prob = 1 (conditioning -> null_cond, token ids -> the null id) and null + (logits - null) * cond_scale."""
    logits = voicebox_forward(sd, cfg, x, cond_drop_prob=0., **kw)
    if cond_scale == 1.:
        return logits
    null = voicebox_forward(sd, cfg, x, cond_drop_prob=1., **kw)
    return null + (logits - null) * cond_scale


def cfm_interpolate(x0, x1, times, sigma=0.):
    """vp.py:1403-1410: w = (1-(1-s)t) x0 + t x1 ; flow = x1 - (1-s) x0."""
    t = times[:, None, None]
    return (1 - (1 - sigma) * t) * x0 + t * x1, x1 - (1 - sigma) * x0


def cfm_loss(sd, cfg, x1, *, sigma=0., cond=None, cond_mask=None, mask=None, x0=None, times=None, prefix='voicebox.'):
    """ConditionalFlowMatcherWrapper.forward core, vp.py:1397-1427.  RNG order: randn_like(x1), rand(B), then the
    mask draws inside voicebox_forward.  x0/times may be injected (golden tests)."""
    if x0 is None:
        x0 = torch.randn_like(x1)
    if times is None:
        times = torch.rand((x1.shape[0],), dtype=x1.dtype, device=x1.device)
    w, flow = cfm_interpolate(x0, x1, times, sigma)
    return voicebox_forward(sd, cfg, w, times=times, cond=cond, cond_mask=cond_mask, target=flow,
                            self_attn_mask=mask, cond_drop_prob=cfg.get('cond_drop_prob', 0.), training=True,
                            prefix=prefix)


def odeint_fixed_grid(fn, y0, t, *, method='midpoint', atol=None, rtol=None, **_):
    """Restatement of torchdiffeq.odeint for its fixed-grid 'euler' and 'midpoint' solvers (un-vendored dependency;
    call site vp.py:1295).  The grid is the user's `t`; t0/t1 are 0-dim tensor slices of t; atol/rtol are accepted and
    ignored by fixed-grid solvers; the solution at every grid point is returned stacked (caller takes [-1])."""
    if method not in ('euler', 'midpoint', 'rk4'):
        raise NotImplementedError(f'only the fixed-grid solvers euler/midpoint/rk4 are on the hot path, got {method!r}')
    ys = [y0]
    y = y0
    for i in range(t.shape[0] - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        f0 = fn(t0, y)
        if method == 'euler':
            dy = dt * f0
        elif method == 'rk4':
            # torchdiffeq RK4._step_func -> rk4_alt_step_func: the 3/8 rule ("smaller error with slightly more compute")
            k2 = fn(t0 + dt / 3, y + dt * f0 / 3)
            k3 = fn(t0 + dt * 2 / 3, y + dt * (k2 - f0 / 3))
            k4 = fn(t1, y + dt * (f0 - k2 + k3))
            dy = dt * (f0 + 3 * (k2 + k3) + k4) / 8
        else:
            half_dt = 0.5 * dt
            dy = dt * fn(t0 + half_dt, y + f0 * half_dt)
        y = y + dy
        ys.append(y)
    return torch.stack(ys)


def cfm_sample(sd, cfg, *, cond, cond_mask=None, steps=3, method='midpoint', y0=None, self_attn_mask=None,
               prefix='voicebox.'):
    """ConditionalFlowMatcherWrapper.sample, unconditional-text branch, cond_scale == 1 (vp.py:1263-1296, 972-978).
    `steps` is the number of GRID POINTS: steps-1 solver intervals."""
    if y0 is None:
        y0 = torch.randn_like(cond)
    t = torch.linspace(0, 1, steps, device=cond.device)

    def fn(tt, x):
        return voicebox_forward(sd, cfg, x, times=tt, cond=cond, cond_mask=cond_mask, self_attn_mask=self_attn_mask,
                                cond_drop_prob=0., training=False, prefix=prefix)

    return odeint_fixed_grid(fn, y0, t, method=method)[-1]


def duration_predictor_forward(sd, cfg, *, cond, phoneme_ids, cond_mask, self_attn_mask=None, prefix=''):
    """DurationPredictor.forward, eval branch with explicit cond_mask and cond_drop_prob=0 (vp.py:757-837)."""
    cond = cond * ~cond_mask[..., None]
    if self_attn_mask is None:
        self_attn_mask = phoneme_ids != -1
    ids = phoneme_ids.clamp(min=0)
    emb = F.embedding(ids, sd[prefix + 'to_phoneme_emb.weight'])
    n = ids.shape[-1]
    if cond.shape[-2] > n:
        cond = cond[..., :n, :]
    elif cond.shape[-2] < n:
        cond = F.pad(cond, (0, 0, 0, n - cond.shape[-2]), value=0.)
    h = F.linear(torch.cat((emb, cond), dim=-1), sd[prefix + 'to_embed.weight'], sd[prefix + 'to_embed.bias'])
    h = conv_pos_embed(h, sd[prefix + 'conv_embed.dw_conv1d.0.weight'], sd[prefix + 'conv_embed.dw_conv1d.0.bias'],
                       self_attn_mask) + h
    h = transformer(sd, h, prefix=prefix + 'transformer.', depth=cfg['depth'], heads=cfg['heads'],
                    qk_norm=cfg.get('qk_norm', True), adaptive=False, mask=self_attn_mask, num_register_tokens=0)
    return F.linear(h, sd[prefix + 'to_pred.0.weight'], sd[prefix + 'to_pred.0.bias'])[..., 0]

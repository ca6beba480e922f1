"""Host-side mirror of the reference's class surface for the CFM hot path.

Same class names, constructor kwargs, attribute paths and state_dict keys as lucidrains/voicebox-pytorch v0.5.0
(`vp.py` = voicebox_pytorch/voicebox_pytorch.py), so checkpoints load both ways and `patch_reference()` can rebind these
`forward`s onto reference-constructed objects.  The modules hold parameters only; every forward routes through the
sm_100a kernels in `ops` (plus library bf16 GEMMs) on a bf16-activation / fp32-residual / fp32-master-weight layout --
the layout the reference itself reaches under `torch.autocast(bfloat16)`.

Not implemented (raise at construction, no silent fallback): GateLoop layers, torchode, dropout p > 0.
"""
import math
import os
from pathlib import Path
from random import random

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from . import pack as _pack
from .ode import odeint_fixed, METHODS

BF16 = torch.bfloat16
# EXPERIMENT (off by default until timed on a B200): batch the adaptive norms' gamma/beta projections, see ops.batched_affine
BATCHED_GAMMA_BETA = os.environ.get('VBX_BATCHED_GB', '0') == '1'
# operand packs (pack.py): bf16 operand copies of all Linear weights refreshed by one launch per optimizer step, fp32 weight
# gradients accumulated straight into the master gradients, all gamma/beta projections as one batched GEMM.  VBX_PACKED=0
# falls back to per-use differentiable casts (the reference's autocast behaviour, launch for launch).
PACKED = os.environ.get('VBX_PACKED', '1') != '0'


def _build_pack(root):
    def builder(pk):
        for lin in (getattr(root, 'to_embed', None), getattr(root, 'to_pred', None)):
            lin = lin[0] if isinstance(lin, nn.Sequential) else lin
            if isinstance(lin, nn.Linear):
                pk.add(lin.weight)
                if lin.bias is not None:
                    pk.add(lin.bias)
        tmlp = getattr(root, 'sinu_pos_emb', None)
        if tmlp is not None:
            pk.add(tmlp[1].weight)
            pk.add(tmlp[1].bias)
        tr = root if type(root).__name__ == 'Transformer' else root.transformer
        _pack.build_for_transformer(pk, tr)
    return builder


def _pack_scope(root):
    """Context manager making `root`'s OperandPack (refreshed if any parameter changed) the active one for the forward inside."""
    if not PACKED or _pack.active() is not None:
        return _pack.use(_pack.active())
    with torch.inference_mode(False), torch.no_grad():   # operand buffers must never be inference tensors
        pk = _pack.for_module(root, _build_pack(root))
        if pk is not None:
            pk.refresh()
    return _pack.use(pk)


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else d


# ---------------------------------------------------------------------------------------------------------------------
# mask / index generation -- bit-exact contract (vp.py:68-74, 121-150): same torch calls, same fp32 op order
# ---------------------------------------------------------------------------------------------------------------------
def prob_mask_like(shape, prob, device):
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


def mask_from_start_end_indices(seq_len, start, end):
    assert start.shape == end.shape
    pos = torch.arange(seq_len, device=start.device, dtype=torch.long)
    pos = pos.reshape(*((1,) * start.ndim), seq_len)
    return (pos >= start[..., None].long()) & (pos < end[..., None].long())


def mask_from_frac_lengths(seq_len, frac_lengths):
    lengths = (frac_lengths * seq_len).long()
    max_start = seq_len - lengths
    rand = torch.zeros_like(frac_lengths).float().uniform_(0, 1)
    start = (max_start * rand).clamp(min=0)
    return mask_from_start_end_indices(seq_len, start, start + lengths)


def reduce_masks_with_and(*masks):
    out = None
    for m in masks:
        if exists(m):
            out = m if out is None else (out & m)
    return out


def interpolate_1d(t, length, mode='bilinear'):
    """vp.py:89-107 (text-conditioned path only)."""
    dtype = t.dtype
    t = t.float()
    squeeze = t.ndim == 2
    if squeeze:
        t = t[:, None]
    t = F.interpolate(t[..., None], (length, 1), mode=mode)[..., 0]
    if squeeze:
        t = t[:, 0]
    return t.to(dtype)


def curtail_or_pad(t, target_length):
    n = t.shape[-2]
    if n > target_length:
        return t[..., :target_length, :]
    if n < target_length:
        return F.pad(t, (0, 0, 0, target_length - n), value=0.)
    return t


# ---------------------------------------------------------------------------------------------------------------------
# parameter holders (attribute names fixed by the state_dict contract)
# ---------------------------------------------------------------------------------------------------------------------
class LearnedSinusoidalPosEmb(nn.Module):
    """vp.py:154-167."""

    def __init__(self, dim):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))

    def forward(self, x):
        f = x.float()[:, None] * self.weights[None, :] * 2 * math.pi
        return torch.cat((f.sin(), f.cos()), dim=-1)


class RotaryEmbedding(nn.Module):
    """vp.py:172-191 (theta 50000, half-split layout).  forward() returns the (n, dim) angle table in fp32."""

    def __init__(self, dim, theta=50000):
        super().__init__()
        self.register_buffer('inv_freq', 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim)))

    @property
    def device(self):
        return self.inv_freq.device

    def forward(self, t):
        if not torch.is_tensor(t):
            t = torch.arange(t, device=self.device)
        fr = t.to(self.inv_freq.dtype)[:, None] * self.inv_freq[None, :]
        return torch.cat((fr, fr), dim=-1)


class ConvPositionEmbed(nn.Module):
    """vp.py:203-233.  Returns the conv branch only (the caller adds the residual), like the reference."""

    def __init__(self, dim, *, kernel_size, groups=None):
        super().__init__()
        assert kernel_size % 2 == 1
        groups = default(groups, dim)
        if groups != dim:
            raise NotImplementedError('only the full depthwise conv (groups == dim, the reference default) is implemented')
        self.dw_conv1d = nn.Sequential(nn.Conv1d(dim, dim, kernel_size, groups=groups, padding=kernel_size // 2), nn.GELU())

    def forward(self, x, mask=None):
        conv = self.dw_conv1d[0]
        xb = x.to(BF16)
        y = ops.convpos_residual_pack(xb, conv.weight, conv.bias, mask, None)
        return (y - xb.float()).to(x.dtype)


class RMSNorm(nn.Module):
    """vp.py:237-247."""

    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        shp = x.shape
        _, h = ops.resid_norm(x.reshape(-1, 1, shp[-1]).float(), None, self.gamma)
        return h.reshape(shp).to(x.dtype)


class AdaptiveRMSNorm(nn.Module):
    """vp.py:249-276 (identity init: gamma = 1, beta = 0)."""

    def __init__(self, dim, cond_dim=None):
        super().__init__()
        cond_dim = default(cond_dim, dim)
        self.scale = dim ** 0.5
        self.to_gamma = nn.Linear(cond_dim, dim)
        self.to_beta = nn.Linear(cond_dim, dim)
        nn.init.zeros_(self.to_gamma.weight)
        nn.init.ones_(self.to_gamma.bias)
        nn.init.zeros_(self.to_beta.weight)
        nn.init.zeros_(self.to_beta.bias)

    def gamma_beta(self, cond_bf16):
        g = ops.linear(cond_bf16, self.to_gamma.weight, self.to_gamma.bias).float()
        b = ops.linear(cond_bf16, self.to_beta.weight, self.to_beta.bias).float()
        return g, b

    def forward(self, x, *, cond):
        g, b = self.gamma_beta(cond.to(BF16))
        _, h = ops.resid_norm(x.float(), None, g, b)
        return h.to(x.dtype)


class MultiheadRMSNorm(nn.Module):
    """vp.py:280-287: parameter holder; applied inside the attention prologue kernel."""

    def __init__(self, dim, heads):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(heads, 1, dim))


class Attention(nn.Module):
    """vp.py:289-333."""

    def __init__(self, dim, dim_head=64, heads=8, dropout=0, flash=False, qk_norm=False, qk_norm_scale=10):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError('the tcgen05 attention kernel is specialised for dim_head = 64 (the reference default)')
        if dropout:
            raise NotImplementedError('attention dropout > 0 is not implemented (no Philox-exact dropout yet)')
        self.heads = heads
        self.qk_norm = qk_norm
        self.scale = qk_norm_scale if qk_norm else dim_head ** -0.5
        self.flash = flash  # accepted for API parity; the fused kernel replaces both reference paths
        if qk_norm:
            self.q_norm = MultiheadRMSNorm(dim_head, heads=heads)
            self.k_norm = MultiheadRMSNorm(dim_head, heads=heads)
        self.to_qkv = nn.Linear(dim, dim_head * heads * 3, bias=False)
        self.to_out = nn.Linear(dim_head * heads, dim, bias=False)

    def _core(self, h, key_mask, cosv, sinv):
        """h bf16 [B,N,D] (normed input) -> bf16 [B,N,D] attention branch (to_out applied)."""
        qkv = ops.linear(h, self.to_qkv.weight)
        gq = self.q_norm.gamma if self.qk_norm else None
        gk = self.k_norm.gamma if self.qk_norm else None
        scale = getattr(self, 'scale', None)
        if scale is None:  # reference-constructed Attention (patch_reference): the scale lives on its Attend module
            scale = default(self.attend.scale, 64 ** -0.5)
        o = ops.attention(qkv, cosv, sinv, gq, gk, key_mask, scale, self.heads)
        return ops.linear(o, self.to_out.weight)

    def forward(self, x, mask=None, rotary_emb=None):
        n = x.shape[1]
        if exists(rotary_emb):
            ang = rotary_emb[:, :32].float()
        else:
            ang = torch.zeros((n, 32), device=x.device)
        return self._core(x.to(BF16), mask, ang.cos().contiguous(), ang.sin().contiguous()).to(x.dtype)


class GEGLU(nn.Module):
    """vp.py:337-340: first half = value, second half = gate."""

    def forward(self, x):
        return ops.geglu(x.to(BF16)).to(x.dtype)


def FeedForward(dim, mult=4, dropout=0.):
    """vp.py:342-349.  Sequential indices (0: Linear, 1: GEGLU, 2: Dropout, 3: Linear) are part of the state_dict contract."""
    if dropout:
        raise NotImplementedError('feed-forward dropout > 0 is not implemented (no Philox-exact dropout yet)')
    inner = int(dim * mult * 2 / 3)
    return nn.Sequential(nn.Linear(dim, inner * 2), GEGLU(), nn.Dropout(dropout), nn.Linear(inner, dim))


def _round_up(v, m):
    return (v + m - 1) // m * m


def _ff_branch(ff, h):
    """FeedForward on bf16 h through the fused GEGLU kernel.  The inner width F = int(dim*8/3) is generally not a
    multiple of 8 (2730 at dim 1024): rows of the (value|gate) GEMM output would be 4-byte aligned only.  The bf16
    operand copies are therefore zero-padded to Fp = roundup(F, 64) -- exact, since gelu(0) * 0 = 0 and the padded
    columns of the second weight are zero."""
    lin1, lin2 = ff[0], ff[3]
    pk = _pack.active()
    if pk is not None:
        ew, eb = pk.lookup(lin1.weight), pk.lookup(lin1.bias)
        if ew is not None and eb is not None:
            y = ops.feed_forward_packed(h, lin1, lin2, pk, ff.__dict__.get('_vbx_w2t'))   # training: one node, fused epilogues
            if y is not None:
                return y
            g = ops.linear_geglu_packed(h, lin1.weight, lin1.bias, ew, eb)
            return ops.linear(g, lin2.weight, lin2.bias)
    f = lin2.in_features
    fp = _round_up(f, 64)
    if fp == f:
        w1, b1, w2 = ops.cast_bf16(lin1.weight, 'w'), ops.cast_bf16(lin1.bias, 'b'), ops.cast_bf16(lin2.weight, 'w')
    else:
        def pad_w1(w):  # [2F, D] -> [2Fp, D]: value rows, zeros, gate rows, zeros
            out = w.new_zeros((2 * fp,) + tuple(w.shape[1:]))
            out[:f] = w[:f]
            out[fp:fp + f] = w[f:]
            return out

        w1 = ops.cast_bf16(lin1.weight, 'w1p', pad_w1)
        b1 = ops.cast_bf16(lin1.bias, 'b1p', pad_w1)
        w2 = ops.cast_bf16(lin2.weight, 'w2p', lambda w: F.pad(w, (0, fp - f)))
    g = ops.linear_geglu(h, w1, b1)
    return F.linear(g, w2, ops.cast_bf16(lin2.bias, 'b'))


# ---------------------------------------------------------------------------------------------------------------------
# Transformer trunk (vp.py:353-479)
# ---------------------------------------------------------------------------------------------------------------------
class Transformer(nn.Module):
    def __init__(self, dim, *, depth, dim_head=64, heads=8, ff_mult=4, attn_dropout=0., ff_dropout=0.,
                 num_register_tokens=0., attn_flash=False, adaptive_rmsnorm=False, adaptive_rmsnorm_cond_dim_in=None,
                 use_unet_skip_connection=False, skip_connect_scale=None, attn_qk_norm=False, use_gateloop_layers=False,
                 gateloop_use_jax=False):
        super().__init__()
        assert depth % 2 == 0
        if use_gateloop_layers:
            raise NotImplementedError('GateLoop layers (third-party gateloop_transformer) are out of scope')
        self.layers = nn.ModuleList([])
        self.rotary_emb = RotaryEmbedding(dim=dim_head)
        self.num_register_tokens = int(num_register_tokens)
        self.has_register_tokens = num_register_tokens > 0
        if self.has_register_tokens:
            self.register_tokens = nn.Parameter(torch.randn(self.num_register_tokens, dim))
        self.adaptive_rmsnorm = adaptive_rmsnorm
        norm = (lambda: AdaptiveRMSNorm(dim, cond_dim=adaptive_rmsnorm_cond_dim_in)) if adaptive_rmsnorm else (lambda: RMSNorm(dim))
        self.skip_connect_scale = default(skip_connect_scale, 2 ** -0.5)
        for ind in range(depth):
            has_skip = use_unet_skip_connection and (ind + 1) > (depth // 2)
            self.layers.append(nn.ModuleList([
                nn.Linear(dim * 2, dim) if has_skip else None,
                None,  # GateLoop slot (index 1) -- keeps the reference's ModuleList numbering
                norm(),
                Attention(dim=dim, dim_head=dim_head, heads=heads, dropout=attn_dropout, flash=attn_flash, qk_norm=attn_qk_norm),
                norm(),
                FeedForward(dim=dim, mult=ff_mult, dropout=ff_dropout),
            ]))
        self.final_norm = RMSNorm(dim)

    @property
    def device(self):
        return next(self.parameters()).device

    forward = None  # bound below (shared with patch_reference)


def _rotary_tables(self, n, device):
    """cos/sin of positions (x) inv_freq, registers at -10000 (vp.py:436-443); torch's range reduction, fp32 [n',32]."""
    r = int(self.num_register_tokens) if self.has_register_tokens else 0
    cache = self.__dict__.setdefault('_vbx_rotary_cache', {})
    key = (n, r, str(device))
    hit = cache.get(key)
    if hit is None:
        # sample() runs under inference_mode: tensors created there cannot be saved for a later backward, so the table is
        # always built as a normal (non-inference) tensor
        with torch.inference_mode(False), torch.no_grad():
            pos = torch.arange(n, device=device, dtype=torch.long)
            if r:
                pos = torch.cat((torch.full((r,), -10000, device=device, dtype=torch.long), pos))
            fr = pos.to(self.rotary_emb.inv_freq.dtype)[:, None] * self.rotary_emb.inv_freq.to(device)[None, :]
            hit = (fr.cos().float().contiguous(), fr.sin().float().contiguous())
        cache[key] = hit
    return hit


def transformer_trunk(self, x, mask, cond, n_out):
    """x: fp32 residual stream [B, R+N, D] with register tokens already packed on the left; mask: bool [B, R+N] or None;
    cond: time embedding [B, cond_dim] or None.  Returns final_norm(x[:, R:]) in bf16 [B, N, D].

    Per layer: ONE fused kernel adds the previous branch to the residual stream and emits the next norm's bf16 output
    (adaptive: gamma/beta from the time embedding); attention prologue + tcgen05 flash attention; fused GEGLU."""
    B, n_all, D = x.shape
    r = n_all - n_out
    cosv, sinv = _rotary_tables(self, n_out, x.device)
    inplace = not torch.is_grad_enabled()
    cond_bf16 = cond.to(BF16) if exists(cond) else None
    if exists(mask):
        mask = mask.contiguous()

    batched = {}
    stack = self.__dict__.get('_vbx_gb_stack') if _pack.active() is not None else None
    if stack is not None and exists(cond_bf16):
        # every adaptive norm's (gamma, beta) from ONE batched GEMM on the packed, stacked operands (pack.py)
        W, bvec, _ = stack
        norms = [n for layer in self.layers for n in (layer[2], layer[4]) if hasattr(n, 'to_gamma')]
        params = [t for n in norms for lin in (n.to_gamma, n.to_beta) for t in (lin.weight, lin.bias)]
        gbs = ops.batched_affine_packed(cond_bf16, W, bvec, params)
        batched = {id(n): (gbs[2 * i], gbs[2 * i + 1]) for i, n in enumerate(norms)}
    elif BATCHED_GAMMA_BETA and exists(cond_bf16):
        # every adaptive norm's (gamma, beta) from ONE batched GEMM over the time embedding instead of 2 tiny GEMMs per norm
        norms = [n for layer in self.layers for n in (layer[2], layer[4]) if hasattr(n, 'to_gamma')]
        if norms:
            gbs = ops.batched_affine(cond_bf16, [w for n in norms for w in (n.to_gamma.weight, n.to_beta.weight)],
                                     [b for n in norms for b in (n.to_gamma.bias, n.to_beta.bias)])
            batched = {id(n): (gbs[2 * i], gbs[2 * i + 1]) for i, n in enumerate(norms)}

    def gb(norm):
        if id(norm) in batched:
            return batched[id(norm)]
        if hasattr(norm, 'to_gamma'):  # AdaptiveRMSNorm: gamma/beta from the time embedding (vp.py:273)
            return AdaptiveRMSNorm.gamma_beta(norm, cond_bf16)
        return norm.gamma, None

    has_skips = any(exists(layer[0]) for layer in self.layers)
    inplace = inplace and not has_skips  # saved skip tensors must survive the later residual updates
    pending = None  # bf16 branch output not yet added to the residual stream (fused into the next norm kernel)
    skips = []
    for skip_combiner, gateloop, attn_norm, attn, ff_norm, ff in self.layers:
        if exists(gateloop):
            raise NotImplementedError('GateLoop layers are out of scope')
        if has_skips:  # U-Net skips (vp.py:458-463): off in VoiceBox / DurationPredictor; plain torch ops
            if exists(pending):
                x, pending = x + pending.float(), None
            if not exists(skip_combiner):
                skips.append(x)
            else:
                s = skips.pop() * self.skip_connect_scale
                x = ops.linear(torch.cat((x, s), dim=-1).to(BF16), skip_combiner.weight, skip_combiner.bias).float()
        g, b = gb(attn_norm)
        x, h = ops.resid_norm(x, pending, g, b, inplace=inplace)
        a = Attention._core(attn, h, mask, cosv, sinv)
        g, b = gb(ff_norm)
        x, h = ops.resid_norm(x, a, g, b, inplace=inplace)
        pending = _ff_branch(ff, h)
    _, h = ops.resid_norm(x, pending, self.final_norm.gamma, None, row0=r, rows=n_out, need_x_out=False)
    return h


def transformer_forward(self, x, mask=None, adaptive_rmsnorm_cond=None):
    """Transformer.forward (vp.py:412-479): public entry, any float dtype in, same dtype out."""
    B, n, _ = x.shape
    x32 = x.float()
    if x32 is x or x32.data_ptr() == x.data_ptr():
        x32 = x32.clone()  # the trunk updates the residual stream in place under no_grad: never alias the caller's tensor
    if self.has_register_tokens:
        x32 = torch.cat((self.register_tokens.float()[None].expand(B, -1, -1), x32), dim=1)
        if exists(mask):
            mask = F.pad(mask, (int(self.num_register_tokens), 0), value=True)
    with _pack_scope(self):
        return transformer_trunk(self, x32.contiguous(), mask, adaptive_rmsnorm_cond, n).to(x.dtype)


Transformer.forward = transformer_forward


# ---------------------------------------------------------------------------------------------------------------------
# audio codec interface (vp.py:483-592) -- duck-typed; pretrained codecs themselves are out of scope
# ---------------------------------------------------------------------------------------------------------------------
class AudioEncoderDecoder(nn.Module):
    """Base class users subclass: latent_dim, sampling_rate, downsample_factor, encode(audio) -> (b, n, d),
    decode(latents), decode_to_codes(latents)."""
    pass


# ---------------------------------------------------------------------------------------------------------------------
# VoiceBox (vp.py:878-1115)
# ---------------------------------------------------------------------------------------------------------------------
class VoiceBox(nn.Module):
    def __init__(self, *, num_cond_tokens=None, audio_enc_dec=None, dim_in=None, dim_cond_emb=1024, dim=1024, depth=24,
                 dim_head=64, heads=16, ff_mult=4, ff_dropout=0., time_hidden_dim=None, conv_pos_embed_kernel_size=31,
                 conv_pos_embed_groups=None, attn_dropout=0, attn_flash=False, attn_qk_norm=True, use_gateloop_layers=False,
                 num_register_tokens=16, p_drop_prob=0.3, frac_lengths_mask=(0.7, 1.), condition_on_text=True):
        super().__init__()
        assert audio_enc_dec is None or isinstance(audio_enc_dec, nn.Module)
        dim_in = default(dim_in, dim)
        time_hidden_dim = default(time_hidden_dim, dim * 4)
        self.audio_enc_dec = audio_enc_dec
        if exists(audio_enc_dec) and dim != audio_enc_dec.latent_dim:
            self.proj_in = nn.Linear(audio_enc_dec.latent_dim, dim)
        else:
            self.proj_in = nn.Identity()
        self.sinu_pos_emb = nn.Sequential(LearnedSinusoidalPosEmb(dim), nn.Linear(dim, time_hidden_dim), nn.SiLU())
        assert not (condition_on_text and not exists(num_cond_tokens)), \
            'number of conditioning tokens must be specified (whether phonemes or semantic token ids) if training conditional voicebox'
        if not condition_on_text:
            dim_cond_emb = 0
        self.dim_cond_emb = dim_cond_emb
        self.condition_on_text = condition_on_text
        self.num_cond_tokens = num_cond_tokens
        if condition_on_text:
            self.null_cond_id = num_cond_tokens
            self.to_cond_emb = nn.Embedding(num_cond_tokens + 1, dim_cond_emb)
        self.p_drop_prob = p_drop_prob
        self.frac_lengths_mask = frac_lengths_mask
        self.to_embed = nn.Linear(dim_in * 2 + dim_cond_emb, dim)
        self.null_cond = nn.Parameter(torch.zeros(dim_in), requires_grad=False)
        self.conv_embed = ConvPositionEmbed(dim=dim, kernel_size=conv_pos_embed_kernel_size, groups=conv_pos_embed_groups)
        self.transformer = Transformer(dim=dim, depth=depth, dim_head=dim_head, heads=heads, ff_mult=ff_mult,
                                       ff_dropout=ff_dropout, attn_dropout=attn_dropout, attn_flash=attn_flash,
                                       attn_qk_norm=attn_qk_norm, num_register_tokens=num_register_tokens,
                                       adaptive_rmsnorm=True, adaptive_rmsnorm_cond_dim_in=time_hidden_dim,
                                       use_gateloop_layers=use_gateloop_layers)
        dim_out = audio_enc_dec.latent_dim if exists(audio_enc_dec) else dim_in
        self.to_pred = nn.Linear(dim, dim_out, bias=False)

    @property
    def device(self):
        return next(self.parameters()).device

    forward = None
    forward_with_cond_scale = None


def _time_embedding(self, times):
    """sinu_pos_emb (vp.py:916-920, 1082): fp32 sinusoid -> bf16 Linear -> SiLU."""
    s = self.sinu_pos_emb[0](times)
    lin = self.sinu_pos_emb[1]
    return F.silu(ops.linear(s.to(BF16), lin.weight, lin.bias).float())


def _voicebox_body(self, emb, times, self_attn_mask):
    """emb: bf16 [B,N,2*dim_in(+dim_cond_emb)] = cat(x, [cond_emb], cond)  ->  prediction bf16 [B,N,dim_out]."""
    with _pack_scope(self):
        return _voicebox_body_inner(self, emb, times, self_attn_mask)


def _voicebox_body_inner(self, emb, times, self_attn_mask):
    B, n, _ = emb.shape
    tr = self.transformer
    times = _fix_times(times, B)                                                                  # vp.py:1012-1016
    h = ops.linear(emb, self.to_embed.weight, self.to_embed.bias)                                  # vp.py:1078
    conv = self.conv_embed.dw_conv1d[0]
    reg = tr.register_tokens if tr.has_register_tokens else None
    x = ops.convpos_residual_pack(h, conv.weight, conv.bias, self_attn_mask, reg)                # vp.py:1080 + :422-425
    mask = self_attn_mask
    if exists(mask) and tr.has_register_tokens:
        mask = F.pad(mask, (int(tr.num_register_tokens), 0), value=True)
    hfin = transformer_trunk(tr, x, mask, _time_embedding(self, times), n)                        # vp.py:1086-1090
    return ops.linear(hfin, self.to_pred.weight)                                                    # vp.py:1092


def _fix_times(times, batch):
    if times.ndim == 0:
        times = times.reshape(1).expand(batch)
    if times.ndim == 1 and times.shape[0] == 1:
        times = times.expand(batch)
    return times


def _assemble_embedding(self, x, cond, cond_mask, cond_token_ids, drop, self_attn_mask):
    """cat(x, [cond_emb], cond) in bf16 for the generic route (vp.py:1035-1076): conditioning masking, classifier-free drop
    (`drop`: bool [B] or None -- rows whose conditioning becomes null_cond / the null token id, vp.py:1041-1054), token
    embedding gather + interpolate_1d to the frame rate (vp.py:1058-1070)."""
    seq_len = cond.shape[1]
    cond = cond * ~cond_mask[..., None]
    cond_ids = cond_token_ids
    if exists(drop):
        cond = torch.where(drop[:, None, None], self.null_cond, cond)
        cond_ids = torch.where(drop[:, None], self.null_cond_id, cond_token_ids)
    parts = [x]
    if self.condition_on_text:
        cond_emb = self.to_cond_emb(cond_ids)
        if cond_emb.shape[-2] != seq_len:
            cond_emb = interpolate_1d(cond_emb.transpose(1, 2), seq_len).transpose(1, 2)
            if exists(self_attn_mask):
                self_attn_mask = interpolate_1d(self_attn_mask, seq_len)
        parts.append(cond_emb)
    parts.append(cond)
    return torch.cat(parts, dim=-1).to(BF16), self_attn_mask


def voicebox_forward(self, x, *, times, cond_token_ids, self_attn_mask=None, cond_drop_prob=0.1, target=None, cond=None,
                     cond_mask=None):
    """VoiceBox.forward (vp.py:987-1115), same argument meaning and RNG draw order."""
    x = self.proj_in(x)
    cond = default(cond, target)
    if exists(cond):
        cond = self.proj_in(cond)
    batch, seq_len, cond_dim = cond.shape
    assert cond_dim == x.shape[-1]
    times = _fix_times(times, batch)

    if not exists(cond_mask):
        if self.training:
            frac_lengths = torch.zeros((batch,), device=self.device).float().uniform_(*self.frac_lengths_mask)
            cond_mask = mask_from_frac_lengths(seq_len, frac_lengths)
        else:
            cond_mask = torch.ones((batch, seq_len), device=cond.device, dtype=torch.bool)

    # The one-pass kernel writes through raw pointers (no grad_fn): only when nothing upstream can need a gradient.  With a
    # trainable proj_in (audio_enc_dec.latent_dim != dim, vp.py:911-914) x / cond carry a graph and take the torch route below.
    upstream_grad = torch.is_grad_enabled() and (x.requires_grad or cond.requires_grad)
    if not self.condition_on_text and not cond_drop_prob > 0. and not upstream_grad:
        emb = ops.embed_concat(x, cond, cond_mask)            # cond * ~mask, cat, bf16 cast: one pass
    else:
        drop = prob_mask_like(cond.shape[:1], cond_drop_prob, self.device) if cond_drop_prob > 0. else None
        emb, self_attn_mask = _assemble_embedding(self, x, cond, cond_mask, cond_token_ids, drop, self_attn_mask)

    pred = _voicebox_body(self, emb, times, self_attn_mask)
    if not exists(target):
        return pred.to(x.dtype)
    loss_mask = reduce_masks_with_and(cond_mask, self_attn_mask)
    return ops.masked_mse(pred, loss_mask, target=target)


CFG_BATCHED = os.environ.get('VBX_CFG_BATCHED', '1') != '0'


def _cfg_batched(self, x, *, times, cond_token_ids, cond, cond_mask=None, self_attn_mask=None, cond_scale=1.):
    """Classifier-free guidance as ONE forward over a 2B batch: rows [0, B) conditioned, rows [B, 2B) with null conditioning
    (what the reference computes in two sequential passes at cond_drop_prob 0 and 1, vp.py:972-985; at p = 1 `prob_mask_like`
    draws nothing, so there is no RNG order to preserve).  Samples are independent through the whole trunk, so the halves
    equal the two separate passes up to GEMM tile selection; every weight is read once instead of twice per evaluation."""
    x = self.proj_in(x)
    cond = self.proj_in(cond)
    B, seq_len, _ = cond.shape
    times = _fix_times(times, B)
    if not exists(cond_mask):
        cond_mask = torch.ones((B, seq_len), device=cond.device, dtype=torch.bool)
    drop = torch.zeros((2 * B,), device=cond.device, dtype=torch.bool)
    drop[B:] = True
    two = lambda t: None if t is None else torch.cat((t, t), dim=0)
    emb, mask2 = _assemble_embedding(self, two(x), two(cond), two(cond_mask), two(cond_token_ids), drop, two(self_attn_mask))
    pred = _voicebox_body(self, emb, two(times), mask2).to(x.dtype)
    logits, null_logits = pred[:B], pred[B:]
    return null_logits + (logits - null_logits) * cond_scale


@torch.inference_mode()
def voicebox_forward_with_cond_scale(self, *args, cond_scale=1., **kwargs):
    """vp.py:972-985."""
    if cond_scale == 1.:
        return self.forward(*args, cond_drop_prob=0., **kwargs)
    simple = (len(args) == 1 and kwargs.get('target') is None and kwargs.get('cond') is not None
              and (kwargs.get('cond_mask') is not None or not self.training)
              and set(kwargs) <= {'times', 'cond_token_ids', 'cond', 'cond_mask', 'self_attn_mask', 'target'})
    if CFG_BATCHED and simple:
        kw = {k: v for k, v in kwargs.items() if k != 'target'}
        return _cfg_batched(self, args[0], cond_scale=cond_scale, **kw)
    logits = self.forward(*args, cond_drop_prob=0., **kwargs)
    null_logits = self.forward(*args, cond_drop_prob=1., **kwargs)
    return null_logits + (logits - null_logits) * cond_scale


VoiceBox.forward = voicebox_forward
VoiceBox.forward_with_cond_scale = voicebox_forward_with_cond_scale


def voicebox_cfm_loss(self, x0, x1, times, *, sigma, cond_mask=None, self_attn_mask=None):
    """Fused training entry used by ConditionalFlowMatcherWrapper.forward when cond is None (the README path):
    w / flow / cond masking / concat in one kernel, flow recomputed inside the loss kernel (never materialised)."""
    batch, seq_len, _ = x1.shape
    if not exists(cond_mask):
        frac_lengths = torch.zeros((batch,), device=self.device).float().uniform_(*self.frac_lengths_mask)
        cond_mask = mask_from_frac_lengths(seq_len, frac_lengths)
    emb = ops.cfm_embed(x0, x1, times, cond_mask, sigma)
    pred = _voicebox_body(self, emb, times, self_attn_mask)
    loss_mask = reduce_masks_with_and(cond_mask, self_attn_mask)
    return ops.masked_mse(pred, loss_mask, x0=x0, x1=x1, sigma=sigma)


# ---------------------------------------------------------------------------------------------------------------------
# DurationPredictor (vp.py:596-876): eval forward only (the training branch needs the absent NS2 aligner)
# ---------------------------------------------------------------------------------------------------------------------
class DurationPredictor(nn.Module):
    def __init__(self, *, audio_enc_dec=None, tokenizer=None, num_phoneme_tokens=None, dim_phoneme_emb=512, dim=512, depth=10,
                 dim_head=64, heads=8, ff_mult=4, ff_dropout=0., conv_pos_embed_kernel_size=31, conv_pos_embed_groups=None,
                 attn_dropout=0, attn_flash=False, attn_qk_norm=True, use_gateloop_layers=False, p_drop_prob=0.2,
                 frac_lengths_mask=(0.1, 1.), aligner_kwargs=None):
        super().__init__()
        self.audio_enc_dec = audio_enc_dec
        if exists(audio_enc_dec) and dim != audio_enc_dec.latent_dim:
            self.proj_in = nn.Linear(audio_enc_dec.latent_dim, dim)
        else:
            self.proj_in = nn.Identity()
        assert not (exists(tokenizer) and exists(num_phoneme_tokens))
        if not exists(tokenizer) and not exists(num_phoneme_tokens):
            raise NotImplementedError('pass num_phoneme_tokens or a tokenizer: the default espeak Tokenizer lives in the '
                                      'absent third-party naturalspeech2_pytorch package')
        if exists(tokenizer):
            num_phoneme_tokens = tokenizer.vocab_size
        self.tokenizer = tokenizer
        self.to_phoneme_emb = nn.Embedding(num_phoneme_tokens, dim_phoneme_emb)
        self.p_drop_prob = p_drop_prob
        self.frac_lengths_mask = frac_lengths_mask
        self.to_embed = nn.Linear(dim + dim_phoneme_emb, dim)
        self.null_cond = nn.Parameter(torch.zeros(dim), requires_grad=False)
        self.conv_embed = ConvPositionEmbed(dim=dim, kernel_size=conv_pos_embed_kernel_size, groups=conv_pos_embed_groups)
        self.transformer = Transformer(dim=dim, depth=depth, dim_head=dim_head, heads=heads, ff_mult=ff_mult,
                                       ff_dropout=ff_dropout, attn_dropout=attn_dropout, attn_flash=attn_flash,
                                       attn_qk_norm=attn_qk_norm, use_gateloop_layers=use_gateloop_layers)
        self.to_pred = nn.Sequential(nn.Linear(dim, 1), nn.Flatten(-2, -1))  # '... 1 -> ...'

    @property
    def device(self):
        return next(self.parameters()).device

    @torch.inference_mode()
    def forward_with_cond_scale(self, *args, texts=None, phoneme_ids=None, cond_scale=1., return_aligned_phoneme_ids=False, **kwargs):
        if return_aligned_phoneme_ids:
            raise NotImplementedError('aligned phoneme ids need naturalspeech2_pytorch.generate_mask_from_repeats (absent)')
        logits = self.forward(*args, texts=texts, phoneme_ids=phoneme_ids, cond_drop_prob=0., **kwargs)
        if cond_scale == 1.:
            return logits
        null_logits = self.forward(*args, texts=texts, phoneme_ids=phoneme_ids, cond_drop_prob=1., **kwargs)
        return null_logits + (logits - null_logits) * cond_scale

    forward = None


def duration_predictor_forward(self, *, cond, texts=None, phoneme_ids=None, cond_drop_prob=0., target=None, cond_mask=None,
                               mel=None, phoneme_len=None, mel_len=None, phoneme_mask=None, mel_mask=None, self_attn_mask=None,
                               return_aligned_phoneme_ids=False):
    """DurationPredictor.forward, vp.py:757-839 (eval branch)."""
    if self.training:
        raise NotImplementedError('DurationPredictor training needs the naturalspeech2_pytorch aligner (absent; the reference '
                                  'branch is itself unfinished, README.md:155)')
    if return_aligned_phoneme_ids:
        raise NotImplementedError('aligned phoneme ids need naturalspeech2_pytorch.generate_mask_from_repeats (absent)')
    batch, seq_len, _ = cond.shape
    cond = self.proj_in(cond)
    if not exists(phoneme_ids):
        assert exists(self.tokenizer)
        phoneme_ids = self.tokenizer.texts_to_tensor_ids(texts)
    if not exists(cond_mask):
        if random() < 0.5:
            frac_lengths = torch.zeros((batch,), device=self.device).float().uniform_(*self.frac_lengths_mask)
            cond_mask = mask_from_frac_lengths(seq_len, frac_lengths)
        else:
            cond_mask = prob_mask_like((batch, seq_len), self.p_drop_prob, self.device)
    cond = cond * ~cond_mask[..., None]
    if cond_drop_prob > 0.:
        drop = prob_mask_like(cond.shape[:1], cond_drop_prob, cond.device)
        cond = torch.where(drop[:, None, None], self.null_cond, cond)
    if not exists(self_attn_mask):
        self_attn_mask = phoneme_ids != -1
    phoneme_ids = phoneme_ids.clamp(min=0)
    phoneme_emb = self.to_phoneme_emb(phoneme_ids)
    cond = curtail_or_pad(cond, phoneme_ids.shape[-1])
    emb = torch.cat((phoneme_emb, cond), dim=-1).to(BF16)
    n = emb.shape[1]
    with _pack_scope(self):
        h = ops.linear(emb, self.to_embed.weight, self.to_embed.bias)
        conv = self.conv_embed.dw_conv1d[0]
        x = ops.convpos_residual_pack(h, conv.weight, conv.bias, self_attn_mask, None)
        hfin = transformer_trunk(self.transformer, x, self_attn_mask, None, n)
        lin = self.to_pred[0]
        return ops.linear(hfin, lin.weight, lin.bias)[..., 0].to(cond.dtype)


DurationPredictor.forward = duration_predictor_forward


# ---------------------------------------------------------------------------------------------------------------------
# ConditionalFlowMatcherWrapper (vp.py:1119-1427)
# ---------------------------------------------------------------------------------------------------------------------
def is_probably_audio_from_shape(t):
    return exists(t) and (t.ndim == 2 or (t.ndim == 3 and t.shape[1] == 1))


class ConditionalFlowMatcherWrapper(nn.Module):
    def __init__(self, voicebox, text_to_semantic=None, duration_predictor=None, sigma=0., ode_atol=1e-5, ode_rtol=1e-5,
                 use_torchode=False, torchdiffeq_ode_method='midpoint', torchode_method_klass=None, cond_drop_prob=0.):
        super().__init__()
        if use_torchode:
            raise NotImplementedError('torchode (adaptive Tsit5 + torch.compile, vp.py:1297-1322) is out of scope')
        if torchdiffeq_ode_method not in METHODS:
            raise NotImplementedError(f'only the fixed-grid solvers {METHODS} are implemented, got {torchdiffeq_ode_method!r}')
        self.sigma = sigma
        self.voicebox = voicebox
        self.condition_on_text = voicebox.condition_on_text
        assert not (not self.condition_on_text and exists(text_to_semantic)), \
            'TextToSemantic should not be passed in if not conditioning on text'
        self.text_to_semantic = text_to_semantic
        self.duration_predictor = duration_predictor
        if self.condition_on_text and (exists(text_to_semantic) or exists(duration_predictor)):
            assert exists(text_to_semantic) ^ exists(duration_predictor), \
                'you should use either TextToSemantic from Spear-TTS, or DurationPredictor for the text / phoneme to audio alignment, but not both'
        self.cond_drop_prob = cond_drop_prob
        self.use_torchode = False
        self.torchode_method_klass = torchode_method_klass
        self.odeint_kwargs = dict(atol=ode_atol, rtol=ode_rtol, method=torchdiffeq_ode_method)

    @property
    def device(self):
        return next(self.parameters()).device

    def load(self, path, strict=True):
        path = Path(path)
        assert path.exists()
        pkg = torch.load(str(path), map_location='cpu')
        # the reference DurationPredictor owns a third-party `aligner` (vp.py:682-683) that this mirror does not build: its
        # keys are the one documented exception to the state_dict contract and are skipped, not failed on
        own = set(self.state_dict().keys())
        model = {k: v for k, v in pkg['model'].items() if k in own or '.aligner.' not in '.' + k}
        self.load_state_dict(model, strict=strict)
        return pkg

    forward = None
    sample = None


def _encode_if_raw(self, t, input_sampling_rate):
    enc = self.voicebox.audio_enc_dec
    assert exists(enc), 'audio_enc_dec must be set on VoiceBox to train directly on raw audio'
    sr = default(input_sampling_rate, enc.sampling_rate)
    with torch.no_grad():
        enc.eval()
        if sr != enc.sampling_rate:
            from torchaudio.functional import resample
            t = resample(t, sr, enc.sampling_rate)
        return enc.encode(t)


def cfm_forward(self, x1, *, mask=None, semantic_token_ids=None, phoneme_ids=None, cond=None, cond_mask=None,
                input_sampling_rate=None):
    """ConditionalFlowMatcherWrapper.forward (vp.py:1332-1427).  RNG draw order on x1's device is the reference's:
    randn_like(x1) -> rand(B) -> [uniform_(frac) -> uniform_(start)] -> [uniform_ CFG]  (vp.py:1399, 1403, 1025, 146, 1042)."""
    sigma = self.sigma
    vb = self.voicebox
    if is_probably_audio_from_shape(x1):
        if self.condition_on_text and exists(self.text_to_semantic) and not exists(semantic_token_ids):
            raise NotImplementedError('deriving semantic ids from raw audio needs the wav2vec of spear_tts_pytorch (absent)')
        x1 = _encode_if_raw(self, x1, input_sampling_rate)
    if is_probably_audio_from_shape(cond):
        cond = _encode_if_raw(self, cond, input_sampling_rate)
    batch, dtype = x1.shape[0], x1.dtype

    assert self.condition_on_text or not (exists(semantic_token_ids) or exists(phoneme_ids)), \
        'semantic or phoneme ids should not be passed in if not conditioning on text'
    cond_token_ids = None
    if self.condition_on_text:
        if exists(self.text_to_semantic) or exists(semantic_token_ids):
            assert not exists(phoneme_ids), 'phoneme ids are not needed for conditioning with spear-tts text-to-semantic'
            assert exists(semantic_token_ids)
            cond_token_ids = semantic_token_ids
        else:
            assert exists(phoneme_ids)
            cond_token_ids = phoneme_ids

    x0 = torch.randn_like(x1)                                                   # vp.py:1399
    times = torch.rand((batch,), dtype=dtype, device=self.device)              # vp.py:1403
    vb.train()                                                                  # vp.py:1414

    fused = (not self.condition_on_text and cond is None and not self.cond_drop_prob > 0.
             and isinstance(vb.proj_in, nn.Identity))
    if fused:
        return voicebox_cfm_loss(vb, x0, x1, times, sigma=sigma, cond_mask=cond_mask, self_attn_mask=mask)
    t = times[:, None, None]
    w = (1 - (1 - sigma) * t) * x0 + t * x1
    flow = x1 - (1 - sigma) * x0
    return vb(w, cond=cond, cond_mask=cond_mask, times=times, target=flow, self_attn_mask=mask,
              cond_token_ids=cond_token_ids, cond_drop_prob=self.cond_drop_prob)


@torch.inference_mode()
def cfm_sample(self, *, cond=None, texts=None, text_token_ids=None, semantic_token_ids=None, phoneme_ids=None, cond_mask=None,
               steps=3, cond_scale=1., decode_to_audio=True, decode_to_codes=False, max_semantic_token_ids=2048,
               spec_decode=False, spec_decode_gamma=5):
    """ConditionalFlowMatcherWrapper.sample (vp.py:1175-1330) on the fixed-grid euler / midpoint solvers.
    `steps` = number of grid points (steps-1 solver intervals), as in the reference."""
    vb = self.voicebox
    if is_probably_audio_from_shape(cond):
        assert exists(vb.audio_enc_dec)
        vb.audio_enc_dec.eval()
        cond = vb.audio_enc_dec.encode(cond)

    num_cond_inputs = sum(map(exists, (texts, text_token_ids, semantic_token_ids, phoneme_ids)))
    assert num_cond_inputs <= 1
    self_attn_mask = None
    cond_token_ids = None
    if self.condition_on_text:
        if exists(semantic_token_ids):
            cond_token_ids = semantic_token_ids
        elif exists(phoneme_ids) and not exists(self.duration_predictor):
            cond_token_ids = phoneme_ids
        else:
            raise NotImplementedError('text front-ends (TextToSemantic.generate / DurationPredictor alignment) depend on absent '
                                      'third-party packages; pass semantic_token_ids (or aligned phoneme_ids) directly')
        target_len = cond_token_ids.shape[-1]
        if exists(cond):
            cond = curtail_or_pad(cond, target_len)
        else:
            cond = torch.zeros((cond_token_ids.shape[0], target_len, vb.audio_enc_dec.latent_dim), device=self.device)
    else:
        assert num_cond_inputs == 0, 'no conditioning inputs should be given if not conditioning on text'

    vb.eval()
    sampled = odeint_fixed(vb, cond=cond, cond_mask=cond_mask, cond_token_ids=cond_token_ids, self_attn_mask=self_attn_mask,
                           steps=steps, cond_scale=cond_scale, method=self.odeint_kwargs['method'])

    if decode_to_codes and exists(vb.audio_enc_dec):
        return vb.audio_enc_dec.decode_to_codes(sampled)
    if not decode_to_audio or not exists(vb.audio_enc_dec):
        return sampled
    return vb.audio_enc_dec.decode(sampled)


ConditionalFlowMatcherWrapper.forward = cfm_forward
ConditionalFlowMatcherWrapper.sample = cfm_sample

"""torch.autograd bindings of the vbx C ABI (include/vbx.h).  Host code only: shapes, allocation, saved tensors.

Every op here runs hand-written sm_100a CUDA through `_lib.call`; GEMMs (`linear`) are plain library GEMMs through
torch (cuBLASLt) on bf16 operands, exactly where the reference's autocast would run them.  Nothing falls back to CPU.
"""
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import call, ptr, stream

BF16 = torch.bfloat16


def _c(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


def _m(mask):
    """Masks cross the C ABI as one byte per element (0 / 1): anything that is not torch.bool is converted first (a float or
    long mask read as bytes would be silently wrong)."""
    if mask is None:
        return None
    if mask.dtype != torch.bool:
        mask = mask.to(torch.bool)
    return _c(mask)


# ---------------------------------------------------------------------------------------------------------------------
# bf16 operand cache: fp32 master parameters are cast to bf16 at use (what autocast does every forward, trainer.py:267);
# under no_grad (sampling: 2 evaluations per solver step with constant weights) the cast is cached per parameter version.
# ---------------------------------------------------------------------------------------------------------------------
_cast_cache = {}


def cast_bf16(p, tag='', transform=None):
    """bf16 copy of parameter `p` (optionally through `transform`, e.g. the GEGLU zero-padding).  Differentiable when grad
    is enabled; cached per tensor object (weakly referenced) and (version, data_ptr) otherwise."""
    if p is None:
        return None
    if torch.is_grad_enabled() and p.requires_grad:
        q = p.to(BF16)
        return transform(q) if transform is not None else q
    key = (id(p), tag)
    hit = _cast_cache.get(key)
    # the weak reference pins the entry to THIS tensor object: id() values (and, through the caching allocator, device
    # addresses and version counters) are reused once a model is freed and another one is built
    if hit is not None and hit[0]() is p and hit[1] == p._version and hit[2] == p.data_ptr():
        return hit[3]
    with torch.inference_mode(False), torch.no_grad():  # never cache an inference tensor (sample-then-train)
        q = p.detach().to(BF16)
        q = transform(q) if transform is not None else q
    if len(_cast_cache) > 4096:  # entries of freed models are only dropped here
        for k in [k for k, v in _cast_cache.items() if v[0]() is None]:
            del _cast_cache[k]
    _cast_cache[key] = (weakref.ref(p), p._version, p.data_ptr(), q)
    return q


def clear_cast_cache():
    _cast_cache.clear()
    _stack_cache.clear()


class _StackCastBf16(torch.autograd.Function):
    """bf16 stack [K, *shape] of K same-shape fp32 parameters with ONE multi-tensor cast (instead of K casts + a cat);
    the backward hands every parameter the fp32, contiguous slice of the stacked gradient."""

    @staticmethod
    def forward(ctx, *ps):
        out = torch.empty((len(ps),) + tuple(ps[0].shape), device=ps[0].device, dtype=BF16)
        torch._foreach_copy_(list(out.unbind(0)), [p.detach() for p in ps])
        return out

    @staticmethod
    def backward(ctx, g):
        return tuple(g.to(torch.float32, memory_format=torch.contiguous_format).unbind(0))


_stack_cache = {}


def stack_cast_bf16(ps, tag=''):
    """Differentiable when grad is enabled; otherwise (sampling: constant weights, 2 evaluations per solver step) cached on
    the parameter objects and their (version, data_ptr) like `cast_bf16`."""
    ps = list(ps)
    if torch.is_grad_enabled() and any(p.requires_grad for p in ps):
        return _StackCastBf16.apply(*ps)
    key = (tag, id(ps[0]), len(ps))
    sig = tuple((p._version, p.data_ptr()) for p in ps)
    hit = _stack_cache.get(key)
    if hit is not None and hit[0]() is ps[0] and hit[1] == sig:
        return hit[2]
    with torch.inference_mode(False), torch.no_grad():
        q = _StackCastBf16.forward(None, *ps)
    _stack_cache[key] = (weakref.ref(ps[0]), sig, q)
    return q


def batched_affine(cond_bf16, weights, biases):
    """y[k] = cond @ weights[k]^T + biases[k] for K same-shape Linear layers as ONE batched library GEMM (bf16 operands,
    fp32 accumulate, bf16 output -- the same rounding points as K separate `linear` calls).  cond bf16 [B, C];
    weights K x f32 [D, C]; biases K x f32 [D].  Returns a tuple of K contiguous f32 [B, D] tensors (views of one buffer).

    The 48 adaptive norms of the Voicebox trunk need 96 such [B x C] x [C x D] products per forward (vp.py:273-274): done
    one by one they cost ~1100 tiny launches per training step (casts, GEMMs, bias reductions, gradient casts)."""
    K = len(weights)
    W = stack_cast_bf16(weights, 'w')            # [K, D, C]
    bvec = stack_cast_bf16(biases, 'b')          # [K, D]
    B, C = cond_bf16.shape
    out = torch.baddbmm(bvec[:, None, :], cond_bf16.unsqueeze(0).expand(K, B, C), W.transpose(1, 2))  # [K, B, D] bf16
    return out.float().unbind(0)


def linear(x, weight, bias=None, tag=''):
    """bf16 library GEMM (cuBLASLt via torch), fp32 accumulate: vp.py:320, 333, 345, 348, 1078, 1092 under autocast."""
    return F.linear(x, cast_bf16(weight, tag + 'w'), cast_bf16(bias, tag + 'b'))


# ---------------------------------------------------------------------------------------------------------------------
# fused residual add + (adaptive) RMSNorm
# ---------------------------------------------------------------------------------------------------------------------
def _adarms_launch(x_in, branch, gamma, beta, per_batch, x_out, h, rstd, B, rows, D, row0, xbs):
    call('vbx_adarms_fwd', ptr(x_in), xbs, row0, ptr(branch), ptr(gamma), ptr(beta), int(per_batch), ptr(x_out), ptr(h),
         ptr(rstd), B, rows, D, stream())


class _ResidNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_in, branch, gamma, beta, row0, rows):
        B, n_in, D = x_in.shape
        per_batch = gamma.dim() == 2
        x_in, branch, gamma, beta = _c(x_in), _c(branch), _c(gamma), _c(beta)
        dense = (row0 == 0 and rows == n_in)
        h = torch.empty((B, rows, D), device=x_in.device, dtype=BF16)
        rstd = torch.empty((B * rows,), device=x_in.device, dtype=torch.float32)
        if branch is None and dense:
            x_out = x_in  # nothing added: the normalised input is x_in itself
            _adarms_launch(x_in, None, gamma, beta, per_batch, None, h, rstd, B, rows, D, row0, n_in * D)
        else:
            x_out = torch.empty((B, rows, D), device=x_in.device, dtype=torch.float32)
            _adarms_launch(x_in, branch, gamma, beta, per_batch, x_out, h, rstd, B, rows, D, row0, n_in * D)
        ctx.save_for_backward(x_out, rstd, gamma)
        ctx.geom = (B, n_in, D, row0, rows, per_batch, branch is not None, beta is not None, dense)
        # nothing was added: hand back no new residual tensor (the caller keeps using x_in; autograd sums the two uses)
        return (None if x_out is x_in else x_out), h

    @staticmethod
    def backward(ctx, dx_out, dh):
        x, rstd, gamma = ctx.saved_tensors
        B, n_in, D, row0, rows, per_batch, has_branch, has_beta, dense = ctx.geom
        dev = x.device
        if dh is None:
            dh = torch.zeros((B, rows, D), device=dev, dtype=BF16)
        dh, dx_out = _c(dh), _c(dx_out)
        dx = torch.empty((B, rows, D), device=dev, dtype=torch.float32)
        dbranch = torch.empty((B, rows, D), device=dev, dtype=BF16) if has_branch else None
        dgamma = torch.zeros_like(gamma, dtype=torch.float32)
        dbeta = torch.zeros_like(gamma, dtype=torch.float32) if has_beta else None
        call('vbx_adarms_bwd', ptr(x), rows * D, 0, ptr(rstd), ptr(gamma), int(per_batch), ptr(dh), ptr(dx_out), ptr(dx),
             ptr(dbranch), ptr(dgamma), ptr(dbeta), B, rows, D, stream())
        if not dense:  # rows skipped by the forward (register tokens) receive zero gradient
            dx = F.pad(dx, (0, 0, row0, n_in - rows - row0))
            if dbranch is not None:
                dbranch = F.pad(dbranch, (0, 0, row0, n_in - rows - row0))
        return dx, dbranch, dgamma, dbeta, None, None


def resid_norm(x_in, branch, gamma, beta=None, *, row0=0, rows=None, inplace=False, need_x_out=True):
    """x_out = x_in + branch ; h = (adaptive) RMSNorm(x_out) in bf16.  Returns (x_out, h).

    gamma: f32 [B,D] (adaptive, with beta [B,D]) or f32 [D] (plain RMSNorm).  x_in/branch: [B, n_in, D]; only rows
    [row0, row0+rows) are processed (final norm after dropping the register tokens, vp.py:476-479)."""
    B, n_in, D = x_in.shape
    rows = n_in if rows is None else rows
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (x_in, branch, gamma, beta))
    if needs_grad:
        x_out, h = _ResidNorm.apply(x_in, branch, gamma, beta, row0, rows)
        return (x_in if x_out is None else x_out), h
    x_in, branch, gamma, beta = _c(x_in), _c(branch), _c(gamma), _c(beta)
    dense = (row0 == 0 and rows == n_in)
    h = torch.empty((B, rows, D), device=x_in.device, dtype=BF16)
    if branch is None or not need_x_out:
        x_out = None
    elif inplace and dense:
        x_out = x_in
    else:
        x_out = torch.empty((B, rows, D), device=x_in.device, dtype=torch.float32)
    _adarms_launch(x_in, branch, gamma, beta, gamma.dim() == 2, x_out, h, None, B, rows, D, row0, n_in * D)
    return (x_in if branch is None else x_out), h


# ---------------------------------------------------------------------------------------------------------------------
# GEGLU
# ---------------------------------------------------------------------------------------------------------------------
class _Geglu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        h = _c(h)
        T, two_f = h.numel() // h.shape[-1], h.shape[-1]
        out = torch.empty(h.shape[:-1] + (two_f // 2,), device=h.device, dtype=BF16)
        call('vbx_geglu_fwd', ptr(h), ptr(out), T, two_f // 2, stream())
        ctx.save_for_backward(h)
        return out

    @staticmethod
    def backward(ctx, dout):
        (h,) = ctx.saved_tensors
        dout = _c(dout)
        T, two_f = h.numel() // h.shape[-1], h.shape[-1]
        dh = torch.empty_like(h)
        call('vbx_geglu_bwd', ptr(h), ptr(dout), ptr(dh), None, T, two_f // 2, stream())
        return dh


def geglu(h):
    """h bf16 [..., 2*Fp] (value | gate) -> gelu_erf(gate) * value, bf16 [..., Fp]   (vp.py:337-340)."""
    return _Geglu.apply(h)


class _LinearGeglu(torch.autograd.Function):
    """g = GEGLU(x @ w^T + b)  (vp.py:345-346) as ONE autograd node: the backward gets the Linear's bias gradient as a by-product
    of the GEGLU backward kernel (column sums of dh held in registers) instead of a separate reduction pass over dh."""

    @staticmethod
    def forward(ctx, x, w, b):
        shp = x.shape
        x2 = _c(x.reshape(-1, shp[-1]))
        h = F.linear(x2, w, b)                     # bf16 library GEMM [T, 2Fp]
        T, two_f = h.shape
        out = torch.empty((T, two_f // 2), device=h.device, dtype=BF16)
        call('vbx_geglu_fwd', ptr(h), ptr(out), T, two_f // 2, stream())
        ctx.save_for_backward(x2, w, h)
        ctx.shp = shp
        return out.reshape(shp[:-1] + (two_f // 2,))

    @staticmethod
    def backward(ctx, dout):
        x2, w, h = ctx.saved_tensors
        T, two_f = h.shape
        dout = _c(dout.reshape(T, two_f // 2))
        dh = torch.empty_like(h)
        db = torch.zeros((two_f,), device=h.device, dtype=torch.float32)
        call('vbx_geglu_bwd', ptr(h), ptr(dout), ptr(dh), ptr(db), T, two_f // 2, stream())
        dx = (dh @ w).reshape(ctx.shp) if ctx.needs_input_grad[0] else None
        dw = dh.t() @ x2 if ctx.needs_input_grad[1] else None
        return dx, dw, db.to(BF16) if ctx.needs_input_grad[2] else None


def linear_geglu(x, w_bf16, b_bf16):
    """x bf16 [..., D], w bf16 [2Fp, D], b bf16 [2Fp] -> bf16 [..., Fp]."""
    if torch.is_grad_enabled() and (x.requires_grad or w_bf16.requires_grad or b_bf16.requires_grad):
        return _LinearGeglu.apply(x, w_bf16, b_bf16)
    return geglu(F.linear(x, w_bf16, b_bf16))


# ---------------------------------------------------------------------------------------------------------------------
# conv positional embedding + residual + register-token pack
# ---------------------------------------------------------------------------------------------------------------------
class _ConvPos(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mask, reg):
        x, mask = _c(x), _m(mask)
        B, N, C = x.shape
        K = weight.shape[-1]
        w2 = _c(weight.reshape(C, K).float())
        b2 = _c(bias.float())
        R = 0 if reg is None else reg.shape[0]
        regc = None if reg is None else _c(reg.float())
        y = torch.empty((B, R + N, C), device=x.device, dtype=torch.float32)
        need_bwd = any(ctx.needs_input_grad)
        pre = torch.empty_like(x) if need_bwd else None
        call('vbx_convpos_fwd', ptr(x), ptr(w2), ptr(b2), ptr(mask), ptr(regc), ptr(y), ptr(pre), B, N, C, K, R, stream())
        if need_bwd:
            ctx.save_for_backward(x, pre, w2, mask)
        ctx.geom = (B, N, C, K, R, weight.shape, reg is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, w2, mask = ctx.saved_tensors
        B, N, C, K, R, wshape, has_reg = ctx.geom
        dy = _c(dy.float())
        dx = torch.empty_like(x)
        dw = torch.zeros((C, K), device=x.device, dtype=torch.float32)
        db = torch.zeros((C,), device=x.device, dtype=torch.float32)
        dreg = torch.zeros((R, C), device=x.device, dtype=torch.float32) if has_reg else None
        call('vbx_convpos_bwd', ptr(x), ptr(pre), ptr(w2), ptr(mask), ptr(dy), ptr(dx), ptr(dw), ptr(db), ptr(dreg), B, N, C, K,
             R, stream())
        return dx, dw.reshape(wshape), db, None, dreg


def convpos_residual_pack(x, weight, bias, mask=None, register_tokens=None):
    """y[:, R:] = gelu(dwconv1d(x*m))*m + x ; y[:, :R] = register tokens.  x bf16 [B,N,C] -> f32 [B,R+N,C]
    (vp.py:203-233 + the `+ x` at vp.py:826/1080 + the register pack at vp.py:422-425)."""
    return _ConvPos.apply(x, weight, bias, mask, register_tokens)


# ---------------------------------------------------------------------------------------------------------------------
# attention: qk-norm + rotary prologue, tcgen05 flash attention
# ---------------------------------------------------------------------------------------------------------------------
class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cosv, sinv, gq, gk, key_mask, scale, heads):
        qkv, key_mask = _c(qkv), _m(key_mask)
        B, N, three_hd = qkv.shape
        H = heads
        hd = three_hd // 3
        assert hd == H * 64, 'dim_head must be 64'
        dev = qkv.device
        gq2 = None if gq is None else _c(gq.reshape(H, 64).float())
        gk2 = None if gk is None else _c(gk.reshape(H, 64).float())
        qh = torch.empty((B, H, N, 64), device=dev, dtype=BF16)
        kh = torch.empty((B, H, N, 64), device=dev, dtype=BF16)
        call('vbx_qkrope_fwd', ptr(qkv), ptr(cosv), ptr(sinv), ptr(gq2), ptr(gk2), ptr(qh), ptr(kh), B, N, H, stream())
        o = torch.empty((B, N, hd), device=dev, dtype=BF16)
        need_bwd = any(ctx.needs_input_grad)
        lse = torch.empty((B, H, N), device=dev, dtype=torch.float32) if need_bwd else None
        v_ptr = qkv.data_ptr() + 2 * hd * 2
        call('vbx_attn_fwd', ptr(qh), ptr(kh), v_ptr, N * three_hd, three_hd, ptr(key_mask), float(scale), ptr(o), ptr(lse), B, H,
             N, stream())
        if need_bwd:
            ctx.save_for_backward(qkv, cosv, sinv, gq2, gk2, key_mask, qh, kh, o, lse)
        ctx.geom = (B, N, H, hd, float(scale), None if gq is None else gq.shape)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, cosv, sinv, gq2, gk2, key_mask, qh, kh, o, lse = ctx.saved_tensors
        B, N, H, hd, scale, gshape = ctx.geom
        dev = qkv.device
        do = _c(do)
        dq = torch.zeros((B, H, N, 64), device=dev, dtype=torch.float32)
        dk = torch.empty((B, H, N, 64), device=dev, dtype=BF16)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty((B, H, N), device=dev, dtype=torch.float32)
        v_ptr = qkv.data_ptr() + 2 * hd * 2
        dv_ptr = dqkv.data_ptr() + 2 * hd * 2
        call('vbx_attn_bwd', ptr(qh), ptr(kh), v_ptr, N * 3 * hd, 3 * hd, ptr(key_mask), scale, ptr(o), ptr(do), ptr(lse),
             ptr(delta), ptr(dq), ptr(dk), dv_ptr, N * 3 * hd, 3 * hd, B, H, N, stream())
        dgq = dgk = None
        if gq2 is not None:
            dgq = torch.zeros((H, 64), device=dev, dtype=torch.float32)
            dgk = torch.zeros((H, 64), device=dev, dtype=torch.float32)
        call('vbx_qkrope_bwd', ptr(qkv), ptr(cosv), ptr(sinv), ptr(gq2), ptr(gk2), ptr(dq), ptr(dk), ptr(dqkv), ptr(dgq), ptr(dgk),
             B, N, H, stream())
        if dgq is not None:
            dgq, dgk = dgq.reshape(gshape), dgk.reshape(gshape)
        return dqkv, None, None, dgq, dgk, None, None, None


def attention(qkv, cosv, sinv, q_gamma, k_gamma, key_mask, scale, heads):
    """qkv bf16 [B,N,3*H*64] (to_qkv output) -> attention output bf16 [B,N,H*64]   (vp.py:320-332 + attend.py:100-137)."""
    return _Attention.apply(qkv, cosv, sinv, q_gamma, k_gamma, key_mask, scale, heads)


# ---------------------------------------------------------------------------------------------------------------------
# CFM passes
# ---------------------------------------------------------------------------------------------------------------------
def cfm_embed(x0, x1, times, cond_mask, sigma):
    """-> bf16 [B,N,2D] = [ w | flow * ~cond_mask ]   (vp.py:1408-1410, 1003, 1035, 1075-1076).  Not differentiable
    (the data does not require grad)."""
    x0, x1, cond_mask = _c(x0.float()), _c(x1.float()), _m(cond_mask)
    B, N, D = x1.shape
    emb = torch.empty((B, N, 2 * D), device=x1.device, dtype=BF16)
    call('vbx_cfm_embed', ptr(x0), ptr(x1), ptr(_c(times.float())), ptr(cond_mask), float(sigma), ptr(emb), B, N, D, stream())
    return emb


def embed_concat(x, cond, cond_mask, out=None):
    """-> bf16 [B,N,2D] = [ x | cond * ~cond_mask ]; either half may be skipped (None) when `out` is given."""
    ref = x if x is not None else cond
    B, N, D = ref.shape
    if out is None:
        out = torch.empty((B, N, 2 * D), device=ref.device, dtype=BF16)
    x = None if x is None else _c(x.float())
    cond = None if cond is None else _c(cond.float())
    call('vbx_embed_concat', ptr(x), ptr(cond), ptr(_m(cond_mask)), ptr(out), B, N, D, stream())
    return out


class _MaskedMse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tgt, x0, x1, sigma, loss_mask):
        pred, loss_mask = _c(pred), _m(loss_mask)
        B, N, D = pred.shape
        num = torch.zeros((B,), device=pred.device, dtype=torch.float32)
        call('vbx_masked_mse_fwd', ptr(pred), ptr(tgt), ptr(x0), ptr(x1), float(sigma), ptr(loss_mask), ptr(num), B, N, D, stream())
        if loss_mask is not None:
            den = loss_mask.sum(dim=-1).clamp(min=1e-5).float()
        else:
            den = torch.full((B,), float(N), device=pred.device)
        ctx.save_for_backward(pred, tgt, x0, x1, loss_mask, den)
        ctx.sigma = float(sigma)
        return (num / den).mean()

    @staticmethod
    def backward(ctx, gout):
        pred, tgt, x0, x1, loss_mask, den = ctx.saved_tensors
        B, N, D = pred.shape
        coef = (2.0 / (D * B)) * gout.float() / den
        dpred = torch.empty_like(pred)
        call('vbx_masked_mse_bwd', ptr(pred), ptr(tgt), ptr(x0), ptr(x1), ctx.sigma, ptr(loss_mask), ptr(_c(coef)), ptr(dpred), B, N,
             D, stream())
        return dpred, None, None, None, None, None


def masked_mse(pred, loss_mask, *, target=None, x0=None, x1=None, sigma=0.):
    """masked-mean MSE (vp.py:1099-1115).  target f32 [B,N,D], or recomputed as x1 - (1-sigma) x0 inside the kernel."""
    if target is not None:
        target = _c(target.float())
    else:
        x0, x1 = _c(x0.float()), _c(x1.float())
    return _MaskedMse.apply(pred.to(BF16), target, x0, x1, sigma, loss_mask)


def ode_axpy(y, f, t, i0, i1, *, half, y_out, emb=None, t_out=None):
    """y_out = y + (half ? 0.5 : 1) * (t[i1]-t[i0]) * f ; optionally refreshes emb[..., :D] and writes t[i0]+dt/2."""
    B, N, D = y.shape
    call('vbx_ode_axpy', ptr(y), ptr(_c(f)), ptr(t), i0, i1, int(half), ptr(y_out), ptr(emb), ptr(t_out), B, N, D, stream())
    return y_out


def umma_selftest(a, b, variant):
    c = torch.empty((128, 128), device=a.device, dtype=torch.float32)
    call('vbx_umma_selftest', ptr(_c(a)), ptr(_c(b)), ptr(c), int(variant), stream())
    return c

"""torch.autograd bindings of the vbx C ABI (include/vbx.h).  Host code only: shapes, allocation, saved tensors.

Every op here runs hand-written sm_100a CUDA through `_lib.call`; GEMMs (`linear`) are plain library GEMMs through
torch (cuBLASLt) on bf16 operands, exactly where the reference's autocast would run them.  Nothing falls back to CPU.
"""
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from . import pack as _pack
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


def _route_wgrad(w, ew, dy2, x2):
    """Weight gradient of y = x W_op^T for the fp32 master parameter `w` behind the (possibly padded / row-remapped) operand:
    accumulated in fp32 straight into w.grad when the flat bucket owns it (returns None), else returned as a new fp32 tensor."""
    cols = w.numel() // w.shape[0]
    sink = _pack.grad_sink(w)
    full = len(ew.maps) == 1 and ew.maps[0][2] == 0 and ew.maps[0][1] == dy2.shape[1]
    if _pack.accum_mode(dy2.device) == 'kernel':
        # ONE library GEMM in the operand's own (zero-padded, 16-byte aligned) shape, bf16 out: cuBLASLt's fast path for every
        # layer.  Then the row blocks / leading columns that exist in the parameter are added into the fp32 gradient in one pass.
        dw_op = torch.mm(dy2.t(), x2)                                       # [op rows, op cols]
        if sink is not None:
            s2 = sink.reshape(w.shape[0], cols)
            for (p0, cnt, o0) in ew.maps:
                call('vbx_accum_bf16_2d', s2.data_ptr() + 4 * p0 * cols, cols, dw_op.data_ptr() + 2 * o0 * dw_op.stride(0), dw_op.stride(0),
                     cnt, cols, stream())
            _pack.sink_done(w)
            return None
        parts = [dw_op[o0:o0 + cnt, :cols] for (p0, cnt, o0) in sorted(ew.maps)]
        return (parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)).float().reshape(w.shape)
    xs = x2 if x2.shape[1] == cols else x2[:, :cols]
    if sink is not None:
        s2 = sink.reshape(w.shape[0], cols)
        for (p0, cnt, o0) in ew.maps:
            _pack.accumulate_wgrad(s2[p0:p0 + cnt], dy2.t() if full else dy2[:, o0:o0 + cnt].t(), xs)
        _pack.sink_done(w)
        return None
    parts = [_pack.wgrad_fp32(dy2.t() if full else dy2[:, o0:o0 + cnt].t(), xs) for (p0, cnt, o0) in sorted(ew.maps)]
    return (parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)).reshape(w.shape)


def _route_bgrad(b, eb, db32):
    """db32: fp32 [op width] column sums of dy.  Same routing for the bias parameter."""
    maps = eb.maps if eb is not None else [(0, b.numel(), 0)]
    sink = _pack.grad_sink(b)
    if sink is not None:
        for (p0, cnt, o0) in maps:
            sink[p0:p0 + cnt].add_(db32[o0:o0 + cnt])
        _pack.sink_done(b)
        return None
    parts = [db32[o0:o0 + cnt] for (p0, cnt, o0) in sorted(maps)]
    return parts[0].clone() if len(parts) == 1 else torch.cat(parts)


class _LinearW(torch.autograd.Function):
    """y = x W^T + b on the packed bf16 operands of the fp32 master parameters (w, b): the forward is the library GEMM, the
    backward routes dW / db to the masters in fp32 (pack.py) instead of through a differentiable cast."""

    @staticmethod
    def forward(ctx, x, w, b, ew, eb):
        shp = x.shape
        x2 = _c(x.reshape(-1, shp[-1]))
        y = F.linear(x2, ew.op, None if b is None else (eb.op if eb is not None else cast_bf16(b)))
        ctx.save_for_backward(x2)
        ctx.misc = (w, b, ew, eb, shp)
        return y.reshape(shp[:-1] + (y.shape[-1],))

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        w, b, ew, eb, shp = ctx.misc
        dy2 = _c(dy.reshape(-1, dy.shape[-1]))
        dx = (dy2 @ ew.op).reshape(shp) if ctx.needs_input_grad[0] else None
        dw = _route_wgrad(w, ew, dy2, x2) if ctx.needs_input_grad[1] else None
        db = None
        if b is not None and ctx.needs_input_grad[2]:
            db = _route_bgrad(b, eb, dy2.sum(dim=0, dtype=torch.float32))
        return dx, dw, db, None, None


_wgrad_buffers = {}   # persistent [K, D, C] bf16 scratch of the batched gamma/beta weight gradient (stable address -> cached table)


class _BatchedAffineW(torch.autograd.Function):
    """All K adaptive-norm projections y[k] = cond W[k]^T + b[k] (vp.py:259-260, 273) as ONE batched GEMM on the packed, stacked
    operands W [K, D, C] / bvec [K, D]; returns fp32 [K, B, D].  Backward: dcond as ONE GEMM over the stacked weights, the K
    weight gradients as fp32-accumulating GEMMs straight into the master gradients, the bias gradients with one reduction."""

    @staticmethod
    def forward(ctx, cond, W, bvec, *params):
        K, D, C = W.shape
        B = cond.shape[0]
        out = torch.baddbmm(bvec[:, None, :], cond.unsqueeze(0).expand(K, B, C), W.transpose(1, 2))   # [K, B, D] bf16
        ctx.save_for_backward(cond)
        ctx.misc = (W, params)
        return out.float()

    @staticmethod
    def backward(ctx, dout):
        (cond,) = ctx.saved_tensors
        W, params = ctx.misc
        K, B, D = dout.shape
        d16 = dout.to(BF16)
        dcond = None
        if ctx.needs_input_grad[0]:
            dcond = d16.permute(1, 0, 2).reshape(B, K * D) @ W.reshape(K * D, -1)
        db_all = dout.sum(dim=1)                                                         # fp32 [K, D]
        grads, b_sinks, b_vals = [], [], []
        w_sinks = [_pack.grad_sink(params[2 * k]) if ctx.needs_input_grad[3 + 2 * k] else None for k in range(K)]
        table_mode = _pack.accum_mode(dout.device) == 'kernel' and all(s_ is not None for s_ in w_sinks)
        if table_mode:
            # all K weight gradients: ONE batched bf16 GEMM into a persistent buffer + ONE table-driven accumulate launch
            C = cond.shape[1]
            buf = _wgrad_buffers.get((K, D, C, str(dout.device)))
            if buf is None:
                buf = torch.empty((K, D, C), device=dout.device, dtype=BF16)
                _wgrad_buffers[(K, D, C, str(dout.device))] = buf
            torch.bmm(d16.transpose(1, 2), cond.unsqueeze(0).expand(K, B, C), out=buf)
            _pack.accumulate_table(w_sinks, buf, D, C)
        for k in range(K):
            w, b = params[2 * k], params[2 * k + 1]
            gw = gb = None
            if ctx.needs_input_grad[3 + 2 * k]:
                if table_mode:
                    _pack.sink_done(w)
                elif w_sinks[k] is not None:
                    _pack.accumulate_wgrad(w_sinks[k], d16[k].t(), cond)
                    _pack.sink_done(w)
                else:
                    gw = _pack.wgrad_fp32(d16[k].t(), cond)
            if ctx.needs_input_grad[4 + 2 * k]:
                sb = _pack.grad_sink(b)
                if sb is not None:
                    b_sinks.append(sb)
                    b_vals.append(db_all[k])
                else:
                    gb = db_all[k].clone()
            grads += [gw, gb]
        if b_sinks:
            torch._foreach_add_(b_sinks, b_vals)
            for k in range(K):
                if _pack.grad_sink(params[2 * k + 1]) is not None:
                    _pack.sink_done(params[2 * k + 1])
        return (dcond, None, None, *grads)


def batched_affine_packed(cond_bf16, W, bvec, params):
    """-> tuple of K contiguous fp32 [B, D] tensors (gamma_0, beta_0, gamma_1, ...) from the packed stack (pack.py)."""
    if torch.is_grad_enabled() and (cond_bf16.requires_grad or any(p.requires_grad for p in params)):
        out = _BatchedAffineW.apply(cond_bf16, W, bvec, *params)
    else:
        K, D, C = W.shape
        out = torch.baddbmm(bvec[:, None, :], cond_bf16.unsqueeze(0).expand(K, cond_bf16.shape[0], C), W.transpose(1, 2)).float()
    return out.unbind(0)


def linear(x, weight, bias=None, tag=''):
    """bf16 library GEMM (cuBLASLt via torch), fp32 accumulate: vp.py:320, 333, 345, 348, 1078, 1092 under autocast.
    With an active OperandPack (pack.py) the bf16 operands are the pre-packed copies and the gradients go to the fp32 masters."""
    pk = _pack.active()
    ew = pk.lookup(weight) if pk is not None else None
    eb = pk.lookup(bias) if (pk is not None and bias is not None) else None
    if ew is None or ew.op.dim() != weight.dim() or (bias is not None and (eb is None or eb.op.dim() != 1)):
        return F.linear(x, cast_bf16(weight, tag + 'w'), cast_bf16(bias, tag + 'b'))   # not packed (or a stacked operand)
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
        return _LinearW.apply(x, weight, bias, ew, eb)
    return F.linear(x, ew.op, None if bias is None else eb.op)


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


# ---------------------------------------------------------------------------------------------------------------------
# tcgen05 GEMM with fused epilogues (csrc/gemm.cu)
# ---------------------------------------------------------------------------------------------------------------------
import os as _os

# FF1 + bias + GEGLU as ONE tcgen05 kernel (csrc/gemm.cu) instead of cuBLASLt GEMM + geglu_fwd: measured 8-16 % faster than the
# pair on the B200 at both the training (T = 66,560) and the sampling (T = 33,024) geometry; VBX_FUSED_FF1=0 restores the pair.
FUSED_FF1 = _os.environ.get('VBX_FUSED_FF1', '1')


def gemm_bf16(a, w, bias=None):
    """C = a w^T (+ bias) on the hand-written tcgen05 GEMM: a bf16 [..., K], w bf16 [N, K], bias bf16 [N] -> bf16 [..., N]."""
    a2 = _c(a.reshape(-1, a.shape[-1]))
    M, K = a2.shape
    N = w.shape[0]
    c = torch.empty((M, N), device=a.device, dtype=BF16)
    call('vbx_gemm_bf16', ptr(a2), ptr(_c(w)), ptr(bias), ptr(c), M, N, K, stream())
    return c.reshape(a.shape[:-1] + (N,))


def ff1_geglu(x2, w_op, b_op, need_h):
    """x2 bf16 [T, K], w_op bf16 [2Fp, K], b_op bf16 [2Fp] -> (h bf16 [T, 2Fp] or None, g bf16 [T, Fp]) in one launch."""
    T, K = x2.shape
    fp = w_op.shape[0] // 2
    h = torch.empty((T, 2 * fp), device=x2.device, dtype=BF16) if need_h else None
    g = torch.empty((T, fp), device=x2.device, dtype=BF16)
    call('vbx_ff1_geglu', ptr(x2), ptr(w_op), ptr(b_op), ptr(h), ptr(g), T, fp, K, stream())
    return h, g


def _use_fused_ff1(fp, k):
    return FUSED_FF1 == '1' and fp % 32 == 0 and k % 8 == 0


class _LinearGeglu(torch.autograd.Function):
    """g = GEGLU(x @ w^T + b)  (vp.py:345-346) as ONE autograd node: the backward gets the Linear's bias gradient as a by-product
    of the GEGLU backward kernel (column sums of dh held in registers) instead of a separate reduction pass over dh."""

    @staticmethod
    def forward(ctx, x, w, b, ew=None, eb=None):
        """w, b: the bf16 operands themselves (ew is None), or the fp32 master parameters behind the packed operands ew / eb."""
        shp = x.shape
        x2 = _c(x.reshape(-1, shp[-1]))
        w_op, b_op = (w, b) if ew is None else (ew.op, eb.op)
        if _use_fused_ff1(w_op.shape[0] // 2, x2.shape[1]):
            h, out = ff1_geglu(x2, _c(w_op), _c(b_op), True)     # tcgen05 GEMM + bias + GEGLU epilogue, h written once
            T, two_f = h.shape
        else:
            h = F.linear(x2, w_op, b_op)           # bf16 library GEMM [T, 2Fp]
            T, two_f = h.shape
            out = torch.empty((T, two_f // 2), device=h.device, dtype=BF16)
            call('vbx_geglu_fwd', ptr(h), ptr(out), T, two_f // 2, stream())
        ctx.save_for_backward(x2, h)
        ctx.misc = (shp, w, b, ew, eb)
        return out.reshape(shp[:-1] + (two_f // 2,))

    @staticmethod
    def backward(ctx, dout):
        x2, h = ctx.saved_tensors
        shp, w, b, ew, eb = ctx.misc
        T, two_f = h.shape
        dout = _c(dout.reshape(T, two_f // 2))
        dh = torch.empty_like(h)
        db = torch.zeros((two_f,), device=h.device, dtype=torch.float32)
        call('vbx_geglu_bwd', ptr(h), ptr(dout), ptr(dh), ptr(db), T, two_f // 2, stream())
        w_op = w if ew is None else ew.op
        dx = (dh @ w_op).reshape(shp) if ctx.needs_input_grad[0] else None
        if ew is None:
            dw = dh.t() @ x2 if ctx.needs_input_grad[1] else None
            return dx, dw, (db.to(BF16) if ctx.needs_input_grad[2] else None), None, None
        dw = _route_wgrad(w, ew, dh, x2) if ctx.needs_input_grad[1] else None
        dbg = _route_bgrad(b, eb, db) if ctx.needs_input_grad[2] else None
        return dx, dw, dbg, None, None


def linear_geglu(x, w_bf16, b_bf16):
    """x bf16 [..., D], w bf16 [2Fp, D], b bf16 [2Fp] -> bf16 [..., Fp]."""
    if torch.is_grad_enabled() and (x.requires_grad or w_bf16.requires_grad or b_bf16.requires_grad):
        return _LinearGeglu.apply(x, w_bf16, b_bf16)
    return _ff1_nograd(x, w_bf16, b_bf16)


def _ff1_nograd(x, w_op, b_op):
    if _use_fused_ff1(w_op.shape[0] // 2, x.shape[-1]):
        _, g = ff1_geglu(_c(x.reshape(-1, x.shape[-1])), _c(w_op), _c(b_op), False)   # inference: h is never written
        return g.reshape(x.shape[:-1] + (g.shape[-1],))
    return geglu(F.linear(x, w_op, b_op))


# FF2 dgrad + GEGLU backward as ONE tcgen05 kernel (csrc/gemm.cu) instead of cuBLASLt dgrad GEMM + geglu_bwd; VBX_FUSED_FF_BWD=0
# restores the pair
FUSED_FF_BWD = _os.environ.get('VBX_FUSED_FF_BWD', '1')


class _FeedForwardPacked(torch.autograd.Function):
    """The whole feed-forward block y = W2 GEGLU(W1 x + b1) + b2 (vp.py:342-349) on packed operands, as one autograd node:
    forward = tcgen05 FF1 GEMM with the bias + GEGLU epilogue (h and g written once) + library FF2 GEMM;
    backward = library wgrad GEMMs (accumulated into the fp32 masters, pack.py) + ONE tcgen05 kernel for the FF2 data gradient
    with the GEGLU backward and the FF1 bias gradient as its epilogue (dg is never written, h is read once) + library FF1 dgrad."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, e1, eb1, e2, eb2, w2t):
        shp = x.shape
        x2 = _c(x.reshape(-1, shp[-1]))
        h, g = ff1_geglu(x2, e1.op, eb1.op, True)
        y = F.linear(g, e2.op, eb2.op)
        ctx.save_for_backward(x2, h, g)
        ctx.misc = (shp, w1, b1, w2, b2, e1, eb1, e2, eb2, w2t)
        return y.reshape(shp[:-1] + (y.shape[-1],))

    @staticmethod
    def backward(ctx, dy):
        x2, h, g = ctx.saved_tensors
        shp, w1, b1, w2, b2, e1, eb1, e2, eb2, w2t = ctx.misc
        dy2 = _c(dy.reshape(-1, dy.shape[-1]))
        T, two_fp = h.shape
        fp = two_fp // 2
        dw2 = _route_wgrad(w2, e2, dy2, g) if ctx.needs_input_grad[3] else None
        db2 = _route_bgrad(b2, eb2, dy2.sum(dim=0, dtype=torch.float32)) if ctx.needs_input_grad[4] else None
        dh = torch.empty_like(h)
        db1_op = torch.zeros((two_fp,), device=h.device, dtype=torch.float32)
        call('vbx_ff2_dgrad_geglu_bwd', ptr(dy2), ptr(w2t), ptr(h), ptr(dh), ptr(db1_op), T, fp, dy2.shape[1], stream())
        dx = (dh @ e1.op).reshape(shp) if ctx.needs_input_grad[0] else None
        dw1 = _route_wgrad(w1, e1, dh, x2) if ctx.needs_input_grad[1] else None
        db1 = _route_bgrad(b1, eb1, db1_op) if ctx.needs_input_grad[2] else None
        return dx, dw1, db1, dw2, db2, None, None, None, None, None


def feed_forward_packed(x, lin1, lin2, pk, w2t):
    """-> FF(x) through _FeedForwardPacked when everything it needs is packed and enabled, else None (caller takes the two-node path)."""
    if FUSED_FF_BWD != '1' or FUSED_FF1 != '1' or w2t is None or not torch.is_grad_enabled():
        return None
    e1, eb1, e2, eb2 = pk.lookup(lin1.weight), pk.lookup(lin1.bias), pk.lookup(lin2.weight), pk.lookup(lin2.bias)
    if None in (e1, eb1, e2, eb2) or (e1.op.shape[0] // 2) % 64 != 0 or x.shape[-1] % 8 != 0 or lin2.out_features % 8 != 0:
        return None
    if not (x.requires_grad or lin1.weight.requires_grad or lin2.weight.requires_grad):
        return None
    return _FeedForwardPacked.apply(x, lin1.weight, lin1.bias, lin2.weight, lin2.bias, e1, eb1, e2, eb2, w2t)


def linear_geglu_packed(x, w, b, ew, eb):
    """Same on the packed operands ew / eb of the fp32 master parameters w [2F, D], b [2F] (pack.py)."""
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or b.requires_grad):
        return _LinearGeglu.apply(x, w, b, ew, eb)
    return _ff1_nograd(x, ew.op, eb.op)


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
        qh = torch.empty((B, N, H, 64), device=dev, dtype=BF16)      # token-major (include/vbx.h)
        kh = torch.empty((B, N, H, 64), device=dev, dtype=BF16)
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
        dq = torch.zeros((B, N, H, 64), device=dev, dtype=torch.float32)
        dk = torch.empty((B, N, H, 64), device=dev, dtype=BF16)
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

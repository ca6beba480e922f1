"""Operand packs: the bf16 tensor-core operand copies of a model's Linear weights, kept fresh by ONE launch per optimizer step.

What the reference does instead: under `torch.autocast(bfloat16)` (trainer.py:267) every `F.linear` casts its fp32 weight and
bias to bf16 at use -- ~14 cast kernels per layer per forward -- and autograd undoes each cast in the backward (bf16 weight
gradient -> fp32 copy -> add into `.grad`): at depth 24 that is ~1,400 of the ~3,000 launches of a training step.

Here a model gets an `OperandPack`:
  * `refresh()` walks a device table of segments with `vbx_pack_bf16` (csrc/pack.cu): every weight / bias -> its operand
    buffer, including the zero-padded GEGLU layout ([2F,D] -> [2Fp,D], [D,F] -> [D,Fp]) and the stacked [K,D,C] layout of the
    adaptive norms' gamma/beta projections.  It runs when any parameter's version counter or address changed (cheap host check).
  * `linear()` / `batched_affine()` in ops.py pick the operand up through `lookup()`; their custom autograd nodes hand the
    weight gradient to the fp32 master parameter WITHOUT the bf16 detour: `grad += dy^T x` as one GEMM with fp32 output
    accumulated in place into the existing `.grad` (a view of the flat gradient bucket, dist.py) when there is one.
Nothing here changes parameter objects, names or values: `state_dict()` is untouched.
"""
import weakref

import torch

from ._lib import call, ptr, stream

BF16 = torch.bfloat16


class _Entry:
    __slots__ = ('param', 'op', 'segs', 'maps')

    def __init__(self, param, op, segs, maps):
        self.param, self.op, self.segs, self.maps = param, op, segs, maps


class OperandPack:
    def __init__(self, device):
        self.device = device
        self.entries = []
        self._by_param = {}      # id(param) -> _Entry
        self._table = None
        self._row_start = None
        self._sig = None
        self._total_rows = 0
        self.post_copies = []    # (dst, src_view): derived operands refreshed right after the pack launch (e.g. transposes)

    # ---- registration ---------------------------------------------------------------------------------------------
    def add(self, param, op_shape=None, row_map=None, op=None, op_offset=0):
        """Register `param` (fp32 [R, C] or [C]).  Default: a bf16 operand of the same shape.
        op_shape: a larger operand (zero padding: extra rows and / or a wider row pitch).
        row_map: list of (param_row0, rows, op_row0) -- row blocks of the parameter land at other rows of the operand
                 (for a 1-D parameter the 'rows' are element ranges).
        op / op_offset: write into an existing operand buffer at an element offset (stacked operands [K, R, C])."""
        if op is None:
            op = torch.zeros(tuple(param.shape) if op_shape is None else tuple(op_shape), device=self.device, dtype=BF16)
        segs = []
        if param.dim() == 1:
            n = param.numel()
            row_map = row_map if row_map is not None else [(0, n, 0)]
            for (p0, cnt, o0) in row_map:
                segs.append((p0, 1, cnt, 0, op_offset + o0, 0))
        else:
            rows = param.shape[0]
            cols = param.numel() // rows
            pitch = op.shape[-1]
            assert pitch >= cols
            row_map = row_map if row_map is not None else [(0, rows, 0)]
            for (p0, cnt, o0) in row_map:
                segs.append((p0 * cols, cnt, cols, cols, op_offset + o0 * pitch, pitch))
        e = _Entry(param, op, segs, list(row_map))
        self.entries.append(e)
        self._by_param[id(param)] = e
        return op

    # ---- refresh ---------------------------------------------------------------------------------------------------
    def _signature(self):
        return tuple((e.param._version, e.param.data_ptr()) for e in self.entries)

    def _build_table(self):
        rows, tab, start = [], [], [0]
        for e in self.entries:
            base_s, base_d = e.param.data_ptr(), e.op.data_ptr()
            for (s_off, nr, nc, s_pitch, d_off, d_pitch) in e.segs:
                tab.append((base_s + 4 * s_off, base_d + 2 * d_off, nr, nc, s_pitch, d_pitch))
                start.append(start[-1] + nr)
        self._table = torch.tensor(tab, dtype=torch.int64).to(self.device)
        self._row_start = torch.tensor(start, dtype=torch.int64).to(self.device)
        self._total_rows = start[-1]
        self._ptrs = tuple(e.param.data_ptr() for e in self.entries)

    def refresh(self, force=False):
        """Re-pack if any registered parameter changed (version counter) or moved (FlatAdam re-points storage).  One launch."""
        sig = self._signature()
        if not force and sig == self._sig:
            return False
        if self._table is None or self._ptrs != tuple(s[1] for s in sig):
            self._build_table()
        with torch.cuda.device(self.device):
            call('vbx_pack_bf16', ptr(self._table), ptr(self._row_start), self._table.shape[0], self._total_rows, stream())
            for dst, src in self.post_copies:
                dst.copy_(src)
        self._sig = sig
        return True

    def lookup(self, param):
        e = self._by_param.get(id(param))
        return None if e is None or e.param is not param else e


# one pack per model root (weakly keyed by the module object): built lazily by the first forward on a CUDA device
_packs = weakref.WeakKeyDictionary()
_active = None   # the pack the current forward runs under (set by modules._with_pack); ops.linear consults it


def active():
    return _active


class use:
    """Context manager: forwards inside run against `pack` (or nothing when pack is None)."""

    def __init__(self, pack):
        self.pack, self.prev = pack, None

    def __enter__(self):
        global _active
        self.prev, _active = _active, self.pack
        return self.pack

    def __exit__(self, *a):
        global _active
        _active = self.prev
        return False


def round_up(v, m):
    return (v + m - 1) // m * m


def build_for_transformer(pack, tr):
    """Register every Linear of a Transformer trunk (vp.py:353-406): attention projections, GEGLU feed-forward (padded
    operands), adaptive-norm gamma/beta projections (stacked [K, D, C] / [K, D]) and U-Net skip combiners."""
    norms = []
    for layer in tr.layers:
        skip, _, attn_norm, attn, ff_norm, ff = layer
        if skip is not None:
            pack.add(skip.weight)
            pack.add(skip.bias)
        pack.add(attn.to_qkv.weight)
        pack.add(attn.to_out.weight)
        lin1, lin2 = ff[0], ff[3]
        f = lin2.in_features
        fp = round_up(f, 64)
        d_in = lin1.in_features
        if fp == f:
            pack.add(lin1.weight)
            pack.add(lin1.bias)
            pack.add(lin2.weight)
        else:
            pack.add(lin1.weight, op_shape=(2 * fp, d_in), row_map=[(0, f, 0), (f, f, fp)])
            pack.add(lin1.bias, op_shape=(2 * fp,), row_map=[(0, f, 0), (f, f, fp)])
            pack.add(lin2.weight, op_shape=(lin2.out_features, fp))
        pack.add(lin2.bias)
        # the transposed FF2 operand [Fp, D] for the fused dgrad + GEGLU-backward GEMM (K-major B operand); 5.6 MB per layer
        e2 = pack.lookup(lin2.weight)
        w2t = torch.zeros((e2.op.shape[1], e2.op.shape[0]), device=pack.device, dtype=BF16)
        pack.post_copies.append((w2t, e2.op.t()))
        ff.__dict__['_vbx_w2t'] = w2t
        for n in (attn_norm, ff_norm):
            if hasattr(n, 'to_gamma'):
                norms.append(n)
    if norms:
        K = 2 * len(norms)
        D, C = norms[0].to_gamma.weight.shape
        W = torch.zeros((K, D, C), device=pack.device, dtype=BF16)
        bvec = torch.zeros((K, D), device=pack.device, dtype=BF16)
        k = 0
        for n in norms:
            for lin in (n.to_gamma, n.to_beta):
                pack.add(lin.weight, op=W, op_offset=k * D * C)
                pack.add(lin.bias, op=bvec, op_offset=k * D)
                k += 1
        tr.__dict__['_vbx_gb_stack'] = (W, bvec, [id(n) for n in norms])


def for_module(root, builder):
    """The pack of `root` (built on first use with `builder(pack)`), or None when root has no CUDA parameters."""
    pk = _packs.get(root)
    if pk is None:
        p0 = next(root.parameters(), None)
        if p0 is None or not p0.is_cuda:
            return None
        pk = OperandPack(p0.device)
        builder(pk)
        _packs[root] = pk
    elif any(e.param.device != pk.device for e in pk.entries[:1]):
        del _packs[root]
        return for_module(root, builder)
    return pk


# ---------------------------------------------------------------------------------------------------------------------
# fp32 weight-gradient accumulation: grad (fp32) += a^T b with bf16 operands, ONE GEMM, no bf16 intermediate
# ---------------------------------------------------------------------------------------------------------------------
_ACCUM_MODE = None


def _probe_accum_mode(device):
    """Which spelling of 'fp32 C += bf16 A x bf16 B' this torch build accepts on CUDA (checked once, numerically)."""
    g = torch.Generator(device='cpu').manual_seed(0)
    a = torch.randn(24, 16, generator=g).to(device=device, dtype=BF16)
    b = torch.randn(24, 40, generator=g).to(device=device, dtype=BF16)
    c0 = torch.randn(16, 40, generator=g).to(device)
    want = c0 + a.float().t() @ b.float()
    for mode in ('addmm_out', 'mm32'):
        try:
            c = c0.clone()
            _accum(c, a.t(), b, mode)
            if torch.allclose(c, want, rtol=1e-3, atol=1e-3):
                return mode
        except Exception:
            pass
    return 'bf16'


def _accum(gview, a_t, b, mode):
    if mode == 'kernel':
        # bf16-output library GEMM (the fast path on every GPU) + ONE fused cast-and-accumulate pass (csrc/pack.cu)
        tmp = torch.mm(a_t, b)
        assert gview.stride(-1) == 1
        call('vbx_accum_bf16_2d', ptr_strided(gview), gview.stride(0), ptr(tmp), tmp.stride(0), tmp.shape[0], tmp.shape[1], stream())
        return
    if mode == 'addmm_out':
        torch.addmm(gview, a_t, b, out_dtype=torch.float32, out=gview)
    elif mode == 'mm32':
        gview.add_(torch.mm(a_t, b, out_dtype=torch.float32))
    else:
        gview.add_(torch.mm(a_t, b))


def accum_mode(device):
    """'kernel' (default): bf16-output library GEMM + one fused cast-and-accumulate pass (csrc/pack.cu) -- every operand shape and
    alignment stays on cuBLASLt's fast path; 'addmm_out' / 'mm32': fp32-output GEMMs (measured slower on the B200 for the
    zero-padded / odd-pitch feed-forward weights, which fall onto a legacy kernel); override with VBX_WGRAD_ACCUM."""
    global _ACCUM_MODE
    if _ACCUM_MODE is None:
        import os
        _ACCUM_MODE = os.environ.get('VBX_WGRAD_ACCUM') or 'kernel'
        if _ACCUM_MODE == 'probe':
            _ACCUM_MODE = _probe_accum_mode(device)
    return _ACCUM_MODE


_tables = {}


def accumulate_table(dsts, src, rows, cols):
    """dsts[k] (f32 [rows, cols] contiguous views) += src[k] (bf16 [K, rows, cols] contiguous) for all k in ONE launch.  The
    device table is cached on (destination addresses, source address): callers keep `src` in a persistent buffer."""
    key = (tuple(d.data_ptr() for d in dsts), src.data_ptr(), rows, cols)
    tab = _tables.get(key)
    if tab is None:
        if len(_tables) > 64:
            _tables.clear()
        base = src.data_ptr()
        seg = [(base + 2 * k * rows * cols, d.data_ptr(), rows, cols, cols, cols) for k, d in enumerate(dsts)]
        start = [k * rows for k in range(len(dsts) + 1)]
        tab = (torch.tensor(seg, dtype=torch.int64).to(src.device), torch.tensor(start, dtype=torch.int64).to(src.device), start[-1])
        _tables[key] = tab
    call('vbx_accum_bf16_table', ptr(tab[0]), ptr(tab[1]), len(dsts), tab[2], stream())


def ptr_strided(t):
    """Device pointer of a row-strided 2-D view (rows need not be contiguous; the last dimension is)."""
    if not t.is_cuda or t.stride(-1) != 1:
        raise RuntimeError('expected a CUDA tensor with a contiguous last dimension')
    return t.data_ptr()


def accumulate_wgrad(gview, a_t, b):
    """gview (fp32 [n, k], may be a strided view) += a_t (bf16 [n, m]) @ b (bf16 [m, k])."""
    _accum(gview, a_t, b, accum_mode(gview.device))


def wgrad_fp32(a_t, b):
    """-> fp32 [n, k] = a_t @ b without a bf16 rounding of the result (used when the parameter has no .grad buffer yet)."""
    if accum_mode(a_t.device) in ('addmm_out', 'mm32'):
        return torch.mm(a_t, b, out_dtype=torch.float32)
    return torch.mm(a_t, b).float()


def grad_sink(param):
    """The fp32 buffer gradients of `param` may be accumulated into in place, or None (autograd then accumulates the returned
    gradient itself).  In-place accumulation is what autograd's AccumulateGrad would do; post-accumulate hooks registered by
    dist.FlatGradBucket are invoked by `sink_done`."""
    g = param.grad
    if g is not None and g.dtype == torch.float32 and g.is_cuda and getattr(param, '_vbx_inplace_grad', False):
        return g
    return None


def sink_done(param):
    hook = getattr(param, '_vbx_post_accum', None)
    if hook is not None:
        hook(param)

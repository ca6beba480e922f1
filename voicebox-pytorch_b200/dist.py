"""Data-parallel gradient exchange for the CFM training step: ONE flat gradient bucket, reduced in a few large chunks.

Replaces, for this path, what the reference gets from accelerate -> DistributedDataParallel(find_unused_parameters=True)
(trainer.py:89-95, 159-164, 263, 270): ~114 x 25 MB NCCL buckets plus an unused-parameter bitmap all-reduce per step.

Here every trainable parameter's `.grad` is a VIEW into one contiguous fp32 buffer laid out in reverse registration
order (= the order backward produces them: to_pred, final norm, layers L-1 .. 0, conv, to_embed, time MLP).  The buffer
is cut into chunks of >= `chunk_bytes`; a post-accumulate-grad hook counts parameters per chunk and, when a chunk is
complete, issues `all_reduce(chunk, async_op=True)` -- NCCL runs it on its own stream over NVLink/NVSwitch while backward
continues.  Parameters that receive no gradient (duration_predictor / text_to_semantic submodules, vp.py:1146-1147) stay
zero, so no unused-parameter discovery is needed; their chunks are flushed by `finish()`.
Sampling needs no collective at all (batch shards are independent).
"""
import contextlib

import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, module, process_group=None, chunk_bytes=64 << 20, overlap=True):
        self.overlap = overlap  # False: one all-reduce over the whole bucket in finish() (no SM sharing with the backward)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in reversed(list(module.parameters())) if p.requires_grad]
        assert self.params, 'no trainable parameters'
        dev = self.params[0].device
        # every parameter starts on a 16-byte boundary (4 floats): the kernels take raw fp32 pointers into this buffer (and into
        # FlatAdam's parameter buffer, same layout) and use float4 accesses; a 1-element bias must not shift everything after it.
        # The padding stays zero: it adds nothing to the gradient norm and Adam leaves (p = 0, g = 0) at 0.
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        total = off
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.chunks = []        # (start, end) element offsets
        self._chunk_of = {}     # id(param) -> chunk index
        self._need = []         # parameters per chunk
        start = 0
        count = 0
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            p.grad = self.flat[o:o + n].view_as(p)
            self._chunk_of[id(p)] = len(self.chunks)
            off = o + (n + 3) // 4 * 4
            count += 1
            if (off - start) * 4 >= chunk_bytes:
                self.chunks.append((start, off))
                self._need.append(count)
                start, count = off, 0
        if off > start:
            self.chunks.append((start, off))
            self._need.append(count)
        self._seen = [0] * len(self.chunks)
        self._sent = [False] * len(self.chunks)
        self._handles = []
        self._sync = True
        backend = dist.get_backend(process_group) if dist.is_initialized() else None
        self._avg = backend == 'nccl'
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)
            # the packed linear layers (pack.py / ops._LinearW) accumulate fp32 weight gradients straight into these views and
            # then call the hook themselves (autograd's AccumulateGrad never sees those gradients)
            p._vbx_inplace_grad = True
            p._vbx_post_accum = self._hook

    # ---- hooks ---------------------------------------------------------------------------------------------------
    def _hook(self, p):
        if not self._sync or self.world == 1 or not self.overlap:
            return
        c = self._chunk_of[id(p)]
        self._seen[c] += 1
        if self._seen[c] == self._need[c] and not self._sent[c]:
            self._launch(c)

    def _launch(self, c):
        s, e = self.chunks[c]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._handles.append(dist.all_reduce(self.flat[s:e], op=op, group=self.group, async_op=True))
        self._sent[c] = True

    # ---- API -----------------------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation micro-steps: accumulate locally, exchange nothing (trainer.py:263)."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def finish(self):
        """Call after backward of the last micro-step: flushes chunks whose parameters got no gradient, waits for the
        collectives (the current stream waits; no host sync with NCCL), and leaves the MEAN gradient in every .grad."""
        if self.world > 1 and self._sync:
            if not self.overlap:
                op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
                self._handles.append(dist.all_reduce(self.flat, op=op, group=self.group, async_op=True))
                self._sent = [True] * len(self.chunks)
            for c in range(len(self.chunks)):
                if not self._sent[c]:
                    self._launch(c)
            for h in self._handles:
                h.wait()
            if not self._avg:
                self.flat.mul_(1.0 / self.world)
        self._handles.clear()
        self._seen = [0] * len(self.chunks)
        self._sent = [False] * len(self.chunks)

    def zero_grad(self):
        """Use instead of optimizer.zero_grad(set_to_none=True), which would detach the .grad views."""
        self.flat.zero_()
        base = self.flat.data_ptr()
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != base + off * 4:
                p.grad = self.flat[off:off + p.numel()].view_as(p)

    def broadcast_parameters(self, module, src=0):
        """Initial replica sync (DDP does this at construction)."""
        if self.world > 1:
            # one collective per dtype instead of one per tensor (~400 at depth 24)
            by_dtype = {}
            for t in list(module.parameters()) + list(module.buffers()):
                by_dtype.setdefault(t.dtype, []).append(t.data)
            for dtype, ts in by_dtype.items():
                flat = torch.cat([t.reshape(-1) for t in ts])
                dist.broadcast(flat, src=src, group=self.group)
                off = 0
                for t in ts:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()

"""Fused gradient-clip + Adam over flat buffers: the optimizer half of the reference's training step
(trainer.py:106-110 `get_optimizer(..., lr, wd)` -> torch Adam(betas=(0.9, 0.99), eps=1e-8) when wd == 0, trainer.py:274-278
`clip_grad_norm_(max_grad_norm)` + `optim.step()`), as ONE launch of `vbx_adam_step` over every parameter.

`FlatAdam` sits on top of `dist.FlatGradBucket`: the bucket already keeps every `.grad` as a view of one fp32 buffer; here
every parameter's storage is re-pointed the same way into one flat fp32 buffer (same order), and the two Adam moments are flat
too.  Module attribute paths, `state_dict()` keys, shapes and values are unchanged -- only the storage moves.
"""
import torch

from ._lib import call, ptr, stream


class FlatAdam:
    def __init__(self, bucket, lr=3e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0., decoupled=False, max_grad_norm=None,
                 bf16_shadow=False):
        # weight_decay > 0 with decoupled=False is torch.optim.Adam's L2 term on every parameter: expressible in one launch
        if weight_decay > 0 and decoupled:
            # the reference's AdamW branch exempts ndim < 2 parameters from decay (optimizer.py:3-8, 24-30): one flat launch
            # cannot express per-parameter decay
            raise NotImplementedError('grouped weight decay (get_optimizer with wd > 0) is not on the flat path; wd == 0 is '
                                      'the reference trainer default (trainer.py:74)')
        self.bucket = bucket
        self.lr, self.betas, self.eps, self.weight_decay, self.decoupled = lr, betas, eps, weight_decay, decoupled
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        params = bucket.params
        total = bucket.flat.numel()
        dev = bucket.flat.device
        self.flat_p = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_p.zero_()             # alignment padding between parameters (bucket.offsets) stays zero forever
        with torch.no_grad():
            for p, off in zip(params, bucket.offsets):
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view           # same nn.Parameter object (module attributes, optimizer-free), new storage
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.flat_p_bf16 = torch.empty(total, dtype=torch.bfloat16, device=dev) if bf16_shadow else None
        self.found_inf = None           # optional device scalar: non-zero skips the step (GradScaler-style contract)
        self.last_grad_norm = None

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        g = self.bucket.flat
        scale = None
        if self.max_grad_norm is not None:
            # clip_grad_norm_: coefficient min(1, max_norm / (norm + 1e-6)); the kernel divides by its reciprocal
            self.last_grad_norm = g.norm()
            scale = torch.clamp((self.last_grad_norm + 1e-6) / self.max_grad_norm, min=1.0).reshape(1)
        call('vbx_adam_step', ptr(self.flat_p), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.flat_p_bf16),
             self.flat_p.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
             float(self.weight_decay), int(self.decoupled), self.step_count, ptr(scale), ptr(self.found_inf), stream())
        # the kernel wrote through raw pointers: tell autograd / the bf16 operand cache that the parameters changed
        torch.autograd.graph.increment_version(self.bucket.params)

    def zero_grad(self):
        self.bucket.zero_grad()

    def state_dict(self):
        return {'step': self.step_count, 'exp_avg': self.exp_avg, 'exp_avg_sq': self.exp_avg_sq,
                'lr': self.lr, 'betas': self.betas, 'eps': self.eps, 'weight_decay': self.weight_decay}

    def load_state_dict(self, sd):
        self.step_count = int(sd['step'])
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])

#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: CFM train-step frames/sec (+ sample ODE-steps/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (N>1: launched by torchrun)
  python bench.py --impl reference [...]                          # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; fits one GPU):
  VoiceBox dim=1024 depth=24 heads=16 dim_head=64, 16 register tokens, seq=1024, batch 64 PER GPU (the reference trainer's
  batch_size is per process, trainer.py:83,149 -> weak scaling), ConditionalFlowMatcherWrapper loss, synthetic fp32 latents,
  random-init weights with the zero-init adaptive-norm weights perturbed.  One "step" = zero_grad + forward + backward +
  gradient all-reduce (N>1) + clip_grad_norm(0.5) + Adam step (trainer.py:261-278).  bf16 tensor-core math / fp32 master
  weights, residual stream and optimizer, i.e. what the reference runs under accelerate's bf16 autocast.
`value`  : frames/s (B*N*gpus*K / time) with the batch resident in HBM.
`e2e`    : same, through the public API `cfm_wrapper(x)` with x copied from pinned host memory every step and loss.item() read back.
`sample` : ODE-steps/sec of ConditionalFlowMatcherWrapper.sample (configs[3]: seq 2048, batch 16/GPU, midpoint), bounded to a
           few solver steps (the per-step cost is constant); reported beside the headline value, not mixed into it.
Inputs (268 MB/step) are larger than L2 (126 MB): no explicit L2 flush is needed between timed iterations.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM, DEPTH, HEADS, SEQ, REG = 1024, 24, 16, 1024, 16
_REAL_STDOUT = sys.stdout
METRIC = 'cfm_train_frames_per_sec'


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='MEASURED_PEAKS.json')
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback (B200_PROFILING.md)')


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (the reference is pure Python; oracle == reference bit-for-bit in fp32, see
# tests/test_oracle_vs_reference.py) on the host cores, bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------------
def oracle_state(depth=DEPTH, seed=0):
    """Random-init state dict with the reference's key names/shapes (no reference import: it is absent on the GPU box)."""
    import voicebox_pytorch_b200 as vbx
    torch.manual_seed(seed)
    vb = vbx.VoiceBox(dim=DIM, depth=depth, heads=HEADS, condition_on_text=False)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in vb.named_parameters():
            if 'to_gamma.weight' in n or 'to_beta.weight' in n:
                p.normal_(0, 0.02, generator=g)
    w = vbx.ConditionalFlowMatcherWrapper(voicebox=vb)
    return {k: v.detach() for k, v in w.state_dict().items()}


CPU_DEPTHS = (2, 6)    # layer counts timed on the CPU; the full model (DEPTH layers) is the two-point linear extrapolation
CPU_BATCH = 2          # cfg3 at reduced batch B=2, per-sample cost scaled linearly


def _host_cores():
    """CPUs this process may actually use: the affinity mask, capped by the cgroup CPU quota when there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


_THREADS = {}


def _calibrate_threads(cores):
    """torch's intra-op thread count for the CPU arm.  'All host threads' is the intent, but on a shared 128-thread host one
    oversubscribed OpenMP team can be 10x slower than a quarter of it (trip 1: 24.6 s vs ~2 s for the same step): time one
    forward of the depth-2 sample at {cores/8, cores/4, cores/2, cores} (>= 8), ascending, and keep the fastest."""
    if cores in _THREADS:
        return _THREADS[cores]
    cands = sorted({max(8, min(cores, c)) for c in (cores, cores // 2, cores // 4, cores // 8)})
    if len(cands) == 1:
        _THREADS[cores] = (cands[0], {})
        return _THREADS[cores]
    fn, _ = _cpu_step_fn(CPU_DEPTHS[0], True, CPU_BATCH, backward=False)   # forward only: a third of the cost, same scaling
    timing = {}
    for t in cands:                                  # ascending; stop as soon as more threads stopped helping
        torch.set_num_threads(t)
        fn()
        t0 = time.perf_counter()
        fn()
        timing[t] = time.perf_counter() - t0
        if timing[t] > 1.25 * min(timing.values()):
            break
    best = min(timing, key=timing.get)
    _THREADS[cores] = (best, {str(k): round(v, 2) for k, v in timing.items()})
    return _THREADS[cores]


def _cpu_step_fn(depth, flash, batch, backward=True):
    """-> (callable running ONE fp32 forward+backward of the CFM loss on the CPU, kind).  kind 'reference': the UNMODIFIED
    reference package (baseline/_ref or /root/reference, stub-imported by oracle/ref_import.py) through its own public API
    `ConditionalFlowMatcherWrapper(x)`; kind 'port': oracle/voicebox_oracle.py (math-path attention only) when no copy of the
    reference is on the box."""
    x1 = torch.randn(batch, SEQ, DIM)
    try:
        from oracle import ref_import
        if ref_import.reference_root() is None:
            raise ImportError('no reference copy')
        vp = ref_import.import_reference()
        torch.manual_seed(0)
        vb = vp.VoiceBox(dim=DIM, depth=depth, heads=HEADS, dim_head=64, attn_flash=flash, condition_on_text=False, num_cond_tokens=None)
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for n, p in vb.named_parameters():
                if 'to_gamma.weight' in n or 'to_beta.weight' in n:
                    p.normal_(0, 0.02, generator=g)
        w = vp.ConditionalFlowMatcherWrapper(voicebox=vb)

        def step():
            if not backward:
                with torch.no_grad():
                    return w(x1)
            w.zero_grad(set_to_none=True)
            w(x1).backward()
        return step, 'reference'
    except Exception as ex:
        if flash:
            raise RuntimeError(f'attn_flash=True needs the reference package on the box ({ex!r})')
        from oracle import voicebox_oracle as O
        sd = oracle_state(depth=depth)
        sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'null_cond' not in k and 'inv_freq' not in k else v)
               for k, v in sd.items()}
        cfg = dict(depth=depth, heads=HEADS, num_register_tokens=REG, qk_norm=True, condition_on_text=False)

        def step():
            if not backward:
                with torch.no_grad():
                    return O.cfm_loss(sdg, cfg, x1)
            for v in sdg.values():
                if v.grad is not None:
                    v.grad = None
            O.cfm_loss(sdg, cfg, x1).backward()
        return step, 'port'


def _median_step_seconds(step, warmup, steps):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), ts


def cpu_reference_frames_per_sec(flash=False, steps=3, warmup=2):
    """The reference's CPU path on all host threads, fp32, on a BOUNDED sample of the cfg3 workload: full width / heads /
    sequence at batch CPU_BATCH, with CPU_DEPTHS = (2, 6) of the 24 identical layers timed (always `warmup` >= 2 untimed steps,
    median of `steps` >= 3).  The per-layer cost is (t6 - t2) / 4 and the depth-independent part (embedding, conv, time MLP,
    to_pred, loss) is t2 - 2 * per_layer, so the full model is fixed + 24 * per_layer -- no cold iteration is ever counted.
    frames/s = CPU_BATCH * SEQ / full_step_seconds."""
    cores = _host_cores()
    threads, calib = _calibrate_threads(cores)
    torch.set_num_threads(threads)
    warmup, steps = max(2, warmup), max(3, steps)
    med, raw, kind = {}, {}, None
    for d in CPU_DEPTHS:
        fn, kind = _cpu_step_fn(d, flash, CPU_BATCH)
        med[d], raw[d] = _median_step_seconds(fn, warmup, steps)
        del fn
    d0, d1 = CPU_DEPTHS
    per_layer = max((med[d1] - med[d0]) / (d1 - d0), 1e-9)
    fixed = max(med[d0] - d0 * per_layer, 0.0)
    full = fixed + DEPTH * per_layer
    sampled = sum(sum(v) for v in raw.values())
    return dict(value=CPU_BATCH * SEQ / full, unit='frames/s', cores=threads, host_cpus=cores, thread_calibration_s=calib, kind=kind,
                attn_flash=bool(flash),
                est_full_step_s=full, per_layer_s=per_layer, fixed_s=fixed, timed_s={str(k): v for k, v in med.items()},
                sample=f'{"reference package" if kind == "reference" else "oracle port"} fp32 fwd+bwd of the CFM loss, dim{DIM} seq{SEQ} heads{HEADS} '
                       f'batch {CPU_BATCH}, attn_flash={bool(flash)}: depth {d0} -> {med[d0]:.2f} s, depth {d1} -> {med[d1]:.2f} s (median of {steps} '
                       f'after {warmup} warm-ups each) => {per_layer:.3f} s/layer + {fixed:.2f} s fixed => {full:.1f} s per full-depth step; '
                       f'{sampled:.0f} s of timed CPU work, threads={torch.get_num_threads()}'), full


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only).  A 'step' of this
    arm is one bounded sample (see cpu_reference_frames_per_sec); `steps` / `ms_per_step` describe what actually ran."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    steps = max(3, min(args.steps, 3))
    t_start = time.perf_counter()
    res = {}
    for flash in (True, False):
        try:
            res[flash], _ = cpu_reference_frames_per_sec(flash=flash, steps=steps, warmup=2)
        except Exception as ex:
            res[flash] = dict(value=None, error=repr(ex))
    best = max((r for r in res.values() if r.get('value')), key=lambda r: r['value'])
    n_timed = steps * len(CPU_DEPTHS) * sum(1 for r in res.values() if r.get('value'))
    wall = time.perf_counter() - t_start
    line = dict(metric=METRIC, value=best['value'], unit='frames/s', n_gpus=args.gpus, steps=n_timed, warmup=2,
                ms_per_step=wall * 1e3 / max(n_timed, 1), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='fp32',
                data='synthetic', impl='reference',
                config=dict(workload=f'VoiceBox dim{DIM} depth{DEPTH} heads{HEADS} seq{SEQ} CFM train step on the host CPU; bounded '
                                     f'sample: batch {CPU_BATCH}, depths {CPU_DEPTHS} timed, linear in depth (see cpu_baseline.sample); '
                                     f'steps = timed sample-steps actually run, ms_per_step = wall time of the whole arm / steps',
                            est_full_step_s=best['est_full_step_s']),
                cpu_baseline=best, attn_flash_true=res[True], attn_flash_false=res[False],
                e2e=dict(value=best['value'], unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def _time_cuda(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def sdpa_baseline(B, H, Np, dev):
    """torch.nn.functional.scaled_dot_product_attention on THIS GPU at the bench geometry (bf16 q,k,v [B,H,N',64], already
    normed / rotated, scale 10 as in the trunk): what the reference's `attn_flash=True` path reaches (attend.py:71-98) per
    backend.  fwd_us = forward alone; bwd_us = (forward+backward) - forward, the number to hold vbx_attn_bwd against."""
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    q, k, v = (torch.randn(B, H, Np, 64, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
    do = torch.randn(B, H, Np, 64, device=dev, dtype=torch.bfloat16)
    out = {}
    for name, be in (('cudnn', SDPBackend.CUDNN_ATTENTION), ('flash', SDPBackend.FLASH_ATTENTION),
                     ('mem_efficient', SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel(be):
                def fwd():
                    with torch.no_grad():
                        return F.scaled_dot_product_attention(q, k, v, scale=10.0)

                def fwdbwd():
                    o = F.scaled_dot_product_attention(q, k, v, scale=10.0)
                    o.backward(do)
                    q.grad = k.grad = v.grad = None
                f_us, fb_us = _time_cuda(fwd), _time_cuda(fwdbwd)
            fl = 4.0 * B * H * Np * Np * 64
            out[name] = dict(fwd_us=f_us, bwd_us=fb_us - f_us, fwd_tflops=fl / f_us / 1e6, bwd_tflops=2.5 * fl / (fb_us - f_us) / 1e6)
        except Exception as ex:
            out[name] = f'unavailable: {type(ex).__name__}: {str(ex)[:120]}'
    return out


def run_durpred(args):
    """BASELINE configs[4]: DurationPredictor transformer dim 512, depth 10, heads 8, seq 512, batch 128 per GPU, eval forward
    (conv positional embedding + plain-RMSNorm trunk with key-padding mask, bf16 residual in the reference).  Replicas only:
    every rank runs its own batch, no collective on the data path.  value = phoneme positions/s over all ranks."""
    import torch.distributed as dist
    import voicebox_pytorch_b200 as vbx
    rank, local, world = (int(os.environ.get(k, 0)) for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'))
    world = max(world, 1)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    Bd, Nd, Dd = 128, 512, 512
    torch.manual_seed(0)
    dp = vbx.DurationPredictor(num_phoneme_tokens=256, dim_phoneme_emb=512, dim=Dd, depth=10, heads=8).to(dev).eval()
    torch.manual_seed(3 + rank)
    ids_host = torch.randint(0, 256, (Bd, Nd))
    for b in range(Bd):
        pad = int(torch.randint(0, 128, (1,)))
        if pad:
            ids_host[b, Nd - pad:] = -1
    ids_host = ids_host.pin_memory()
    cond_host = torch.randn(Bd, Nd, Dd).pin_memory()
    cmask = torch.zeros(Bd, Nd, dtype=torch.bool, device=dev)
    ids, cond = ids_host.to(dev), cond_host.to(dev)

    def fwd():
        with torch.no_grad():
            return dp(cond=cond, phoneme_ids=ids, cond_mask=cmask)

    def fwd_e2e():
        with torch.no_grad():
            return dp(cond=cond_host.to(dev, non_blocking=True), phoneme_ids=ids_host.to(dev, non_blocking=True), cond_mask=cmask).cpu()

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
    for _ in range(max(args.warmup, 3)):
        fwd()
    l0 = vbx._lib.launch_count
    ms = timed(fwd, args.steps)
    launches = vbx._lib.launch_count - l0
    fwd_e2e()
    ms_e2e = timed(fwd_e2e, args.steps)
    if rank == 0:
        units = Bd * Nd * world * args.steps
        line = dict(metric='durpred_eval_positions_per_sec', value=units / (ms / 1e3), unit='positions/s', n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='bf16', data='synthetic', gpu_launches=launches,
                    tflops=4.881 * world * args.steps / (ms / 1e3) / 1e0,
                    config=dict(workload=f'DurationPredictor dim{Dd} depth10 heads8 seq{Nd} batch {Bd}/GPU eval forward (BASELINE configs[4]), '
                                         f'replicas only', parallelism=f'replicas{world}', l2='activations (134 MB per tensor) exceed L2'),
                    e2e=dict(value=units / (ms_e2e / 1e3), unit='positions/s', h2d_bytes_per_step=cond_host.numel() * 4 + ids_host.numel() * 8,
                             d2h_bytes_per_step=Bd * Nd * 4))
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            pass

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons, power = [], [], set(), []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='batch per GPU (BASELINE configs[2]: 64)')
    ap.add_argument('--depth', type=int, default=DEPTH)
    ap.add_argument('--no-sample', action='store_true')
    ap.add_argument('--optimizer', default=os.environ.get('VBX_OPTIMIZER', 'flat'), choices=['torch', 'flat'],
                    help="'flat' (default; validated against torch Adam in tests/test_gpu_kernels.py): voicebox_pytorch_b200.FlatAdam = "
                         "clip + Adam in one vbx_adam_step launch over the flat buffers; 'torch': torch.optim.Adam(fused=True) with the "
                         "clip folded into its grad_scale")
    ap.add_argument('--allreduce', default=os.environ.get('VBX_ALLREDUCE', 'after'), choices=['overlap', 'after'],
                    help='gradient exchange: ONE all-reduce of the flat bucket after backward (default; measured faster: NCCL CTAs '
                         'otherwise take SMs from the 1-CTA/SM backward kernels), or chunked all-reduce overlapped with backward')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sdpa', action='store_true')
    ap.add_argument('--graph', default=os.environ.get('VBX_STEP_GRAPH', 'off'), choices=['on', 'off'],
                    help="'on': capture ONE whole training step (zero_grad, forward, backward, all-reduce, clip, Adam) as a CUDA graph and "
                         "time its replays; falls back to eager launches if the capture fails (reported in config.step_graph)")
    ap.add_argument('--sample-steps', type=int, default=64, help='midpoint solver steps timed in the sampling leg (configs[3]: 64)')
    ap.add_argument('--workload', default='train', choices=['train', 'durpred'],
                    help="'durpred': BASELINE configs[4] -- DurationPredictor dim512 depth10 seq512 batch 128/GPU eval forward, replicas")
    ap.add_argument('--profile-only', action='store_true',
                    help='for ncu launch lists only: allows --warmup < 3, skips the e2e / sampling / CPU legs; the printed number is NOT a bench value')
    args = ap.parse_args()
    # Only the JSON line may reach stdout: keep a private handle on the real stdout and point fd 1 at stderr, so banners printed
    # by libraries (e.g. "NCCL version ...") cannot precede it.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        return run_reference_arm(args)
    if args.workload == 'durpred':
        return run_durpred(args)
    assert args.warmup >= 3 or args.profile_only, 'timing rules: at least 3 warm-up steps'
    if args.profile_only:
        args.no_sample = args.no_cpu_baseline = True

    import torch.distributed as dist
    import voicebox_pytorch_b200 as vbx
    from voicebox_pytorch_b200.dist import FlatGradBucket

    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner (NCCL_DEBUG >= VERSION) to stdout by default
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()
    B, N, D = args.batch, SEQ, DIM

    # ---- model, optimizer, gradient bucket ---------------------------------------------------------------------------
    torch.manual_seed(0)
    vb = vbx.VoiceBox(dim=D, depth=args.depth, heads=HEADS, condition_on_text=False)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n_, p in vb.named_parameters():
            if 'to_gamma.weight' in n_ or 'to_beta.weight' in n_:
                p.normal_(0, 0.02, generator=g)
    w = vbx.ConditionalFlowMatcherWrapper(voicebox=vb).to(dev)
    bucket = FlatGradBucket(w, overlap=(args.allreduce == 'overlap'))
    bucket.broadcast_parameters(w)
    n_params = sum(p.numel() for p in w.parameters() if p.requires_grad)
    if args.graph == 'on' and args.optimizer == 'flat':
        args.optimizer = 'torch'     # FlatAdam keeps the step count on the host (bias corrections would be frozen into the graph)
    if args.optimizer == 'flat':
        opt = vbx.FlatAdam(bucket, lr=3e-4, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=0.5)
    else:
        opt = torch.optim.Adam([p for p in w.parameters() if p.requires_grad], lr=3e-4, betas=(0.9, 0.99), fused=True,
                               capturable=(args.graph == 'on'))
    torch.manual_seed(2 + rank)
    x_host = torch.randn(B, N, D).pin_memory()
    x_dev = x_host.to(dev)

    found_inf = torch.zeros((), device=dev)

    def train_step(x):
        bucket.zero_grad()
        loss = w(x)                                   # ConditionalFlowMatcherWrapper.forward (public API)
        loss.backward()                               # chunked all-reduce overlaps with this (N>1)
        bucket.finish()
        if args.optimizer == 'flat':                  # clip coefficient + Adam in ONE launch over the flat buffers
            opt.step()
            return loss
        gnorm = bucket.flat.norm()                    # clip_grad_norm_(0.5), trainer.py:274-275, on the flat bucket:
        # the clip coefficient c = min(1, 0.5/(norm+1e-6)) is applied INSIDE the fused Adam kernel (its grad_scale input
        # divides every gradient by 1/c) instead of a separate 2 x 2.85 GB scaling pass
        opt.grad_scale = torch.clamp((gnorm + 1e-6) / 0.5, min=1.0)
        opt.found_inf = found_inf
        opt.step()
        return loss

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    torch.manual_seed(1000 + rank * 10 ** 6)
    for _ in range(args.warmup):
        train_step(x_dev)
    torch.cuda.synchronize()

    # ---- optional: the whole step as ONE CUDA graph (same launches, replayed without per-launch host work) ------------------
    step_graph, graph_note, x_static, static_loss, launches_per_graph = None, 'off', None, None, 0
    if args.graph == 'on':
        try:
            x_static = x_dev.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):       # capture prerequisites: every lazily-built table / cache exists, on a side stream
                for _ in range(2):
                    train_step(x_static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            l0 = vbx._lib.launch_count
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                static_loss = train_step(x_static)
            launches_per_graph = vbx._lib.launch_count - l0
            step_graph, graph_note = g_, 'on'
            for _ in range(2):
                g_.replay()
            torch.cuda.synchronize()
        except Exception as ex:
            step_graph, graph_note = None, f'capture failed, eager launches: {ex!r}'[:300]
            torch.cuda.synchronize()

    def run_step(x):
        """One training step on batch x (device tensor): graph replay when captured, eager launches otherwise."""
        if step_graph is None:
            return train_step(x)
        if x is not x_static:
            x_static.copy_(x, non_blocking=True)
        step_graph.replay()
        return static_loss

    # ---- device-resident timed region (value) --------------------------------------------------------------------------------
    tracked = ['vbx_attn_fwd', 'vbx_attn_bwd', 'vbx_adarms_fwd', 'vbx_adarms_bwd', 'vbx_geglu_fwd', 'vbx_geglu_bwd',
               'vbx_qkrope_fwd', 'vbx_qkrope_bwd', 'vbx_convpos_fwd', 'vbx_convpos_bwd', 'vbx_ff1_geglu', 'vbx_ff2_dgrad_geglu_bwd',
               'vbx_pack_bf16',
               'vbx_accum_bf16_2d', 'vbx_accum_bf16_table', 'vbx_adam_step']
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = vbx._lib.launch_count
    if step_graph is None:
        vbx._lib.profile_start(tracked)     # per-kernel CUDA events inside the timed region (eager launches only)
    torch.cuda.nvtx.range_push('timed')   # `ncu --nvtx --nvtx-include "timed/"` captures exactly this region
    ms_dev = timed(lambda: run_step(x_static if step_graph is not None else x_dev), args.steps)
    torch.cuda.nvtx.range_pop()
    launches = (vbx._lib.launch_count - launches0) if step_graph is None else launches_per_graph * args.steps
    clocks = sampler.stop() if sampler else None
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    ms_eager = None
    if step_graph is None:
        prof = vbx._lib.profile_stop()
    else:
        # events cannot be placed inside a replayed graph: the per-kernel timings (roofline) come from a second timed region of
        # the SAME steps launched eagerly, reported as `eager_ms_per_step` next to the graph-replay `ms_per_step`
        vbx._lib.profile_start(tracked)
        ms_eager = timed(lambda: train_step(x_dev), args.steps)
        prof = vbx._lib.profile_stop()

    # ---- end-to-end timed region: host batch in, loss out, every step --------------------------------------------------
    def e2e_step():
        if step_graph is not None:
            x_static.copy_(x_host, non_blocking=True)
            return run_step(x_static).item()
        x = x_host.to(dev, non_blocking=True)
        return train_step(x).item()
    if args.profile_only:
        ms_e2e = float('nan')
    else:
        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, args.steps)

    frames = B * N * world * args.steps
    value = frames / (ms_dev / 1e3)
    e2e_value = frames / (ms_e2e / 1e3)

    # ---- roofline of the dominant hand-written kernel ------------------------------------------------------------------
    Np, T = N + REG, B * (N + REG)
    attn_fl = 4.0 * B * HEADS * Np * Np * 64
    ms_ref = ms_eager if ms_eager is not None else ms_dev   # the timed region the per-kernel events were taken in
    work = {  # algorithmic FLOPs / bytes per launch
        'vbx_attn_fwd': ('tensor', attn_fl), 'vbx_attn_bwd': ('tensor', 2.5 * attn_fl),
        'vbx_adarms_fwd': ('hbm', T * D * 12.0), 'vbx_adarms_bwd': ('hbm', T * D * 16.0),
        'vbx_geglu_fwd': ('hbm', T * 2752 * 6.0), 'vbx_geglu_bwd': ('hbm', T * 2752 * 10.0),
        'vbx_qkrope_fwd': ('hbm', T * 2048 * 4.0), 'vbx_qkrope_bwd': ('hbm', T * 1024 * (2 + 2 + 2 + 4 + 2 + 2.0)),
        'vbx_convpos_fwd': ('hbm', B * N * D * 6.0), 'vbx_convpos_bwd': ('hbm', B * N * D * 10.0),
        'vbx_ff1_geglu': ('tensor', 2.0 * T * D * 2 * 2752), 'vbx_ff2_dgrad_geglu_bwd': ('tensor', 2.0 * T * D * 2752),
        'vbx_pack_bf16': ('hbm', n_params * 6.0), 'vbx_adam_step': ('hbm', n_params * 28.0),
    }
    kernels = {}
    for name, (cnt, tot_ms) in prof.items():
        if cnt == 0:
            continue
        if name not in work:
            kernels[name] = dict(launches_per_step=cnt / args.steps, avg_us=tot_ms / cnt * 1e3, ms_per_step=tot_ms / args.steps)
            continue
        kind, units = work[name]
        avg_ms = tot_ms / cnt
        ach = units / (avg_ms * 1e-3) / (1e12 if kind == 'tensor' else 1e9)
        peak = peaks['tf_sustained'] if kind == 'tensor' else peaks['hbm']
        kernels[name] = dict(bound=kind, launches_per_step=cnt / args.steps, avg_us=avg_ms * 1e3, ms_per_step=tot_ms / args.steps,
                             achieved=ach, peak=peak, unit='TFLOP/s' if kind == 'tensor' else 'GB/s', frac=ach / peak)
    own = [k for k in kernels if 'frac' in kernels[k] and k not in ('vbx_ff1_geglu', 'vbx_ff2_dgrad_geglu_bwd')]
    dom = max(own, key=lambda k: kernels[k]['ms_per_step']) if own else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(dom)
    except Exception:
        pass
    roofline = None
    if dom:
        k = kernels[dom]
        roofline = dict(kernel=dom, bound=k['bound'], achieved=k['achieved'], peak=k['peak'], unit=k['unit'], frac=k['frac'],
                        traffic=traffic, peak_source=peaks['source'] + (' sustained bf16' if k['bound'] == 'tensor' else ' hbm copy'),
                        share_of_step=k['ms_per_step'] / (ms_ref / args.steps))

    # ---- same-box SDPA baseline for the attention kernels (what the reference's attn_flash=True reaches on this GPU) --------
    sdpa = None
    if rank == 0 and not args.profile_only and not args.no_sdpa:
        try:
            sdpa = sdpa_baseline(B, HEADS, N + REG, dev)
            for kname, key in (('vbx_attn_fwd', 'fwd_us'), ('vbx_attn_bwd', 'bwd_us')):
                if kname in kernels:
                    best = min((v[key] for v in sdpa.values() if isinstance(v, dict) and v.get(key)), default=None)
                    kernels[kname]['sdpa_us'] = {b: (v.get(key) if isinstance(v, dict) else v) for b, v in sdpa.items()}
                    kernels[kname]['speedup_vs_best_sdpa'] = (best / kernels[kname]['avg_us']) if best else None
        except Exception as ex:
            sdpa = dict(error=repr(ex))

    # ---- sampling (BASELINE configs[3]): all 64 midpoint steps, CUDA-graph replay of one captured solver step ---------------
    sample = None
    if not args.no_sample:
        step_graph = static_loss = x_static = None      # releases the captured step's private memory pool
        del opt
        bucket.flat = None
        for p in w.parameters():
            p.grad = None
        torch.cuda.empty_cache()
        from voicebox_pytorch_b200 import ode as vode
        SB, SN, solver_steps = 16, 2048, args.sample_steps
        cond_host = torch.randn(SB, SN, D).pin_memory()
        cmask = torch.zeros(SB, SN, dtype=torch.bool, device=dev)
        cmask[:, int(0.3 * SN):] = True
        cond = cond_host.to(dev)
        w.sample(cond=cond, cond_mask=cmask, steps=6)            # warm-up: weights cast + cached, one solver step captured
        ms_s = timed(lambda: w.sample(cond=cond, cond_mask=cmask, steps=solver_steps + 1), 1)
        info = dict(vode.last_run_info)

        def sample_e2e():                                       # public API, host conditioning in, host sample out
            out = w.sample(cond=cond_host.to(dev, non_blocking=True), cond_mask=cmask, steps=solver_steps + 1)
            return out.cpu()
        ms_s_e2e = timed(sample_e2e, 1)
        nfe_tflop = 26.864 * (args.depth / DEPTH)               # per evaluation at B=16, N=2048
        sample = dict(metric='sample_ode_steps_per_sec', value=world * solver_steps / (ms_s / 1e3), unit='ODE-steps/s',
                      ms_per_ode_step=ms_s / solver_steps, solver_steps_timed=solver_steps, cuda_graph=info,
                      tflops=2 * nfe_tflop * solver_steps / (ms_s / 1e3),
                      frac_of_sustained_bf16=(2 * nfe_tflop * solver_steps / (ms_s / 1e3)) / peaks['tf_sustained'],
                      e2e=dict(value=world * solver_steps / (ms_s_e2e / 1e3), unit='ODE-steps/s',
                               h2d_bytes_per_call=cond_host.numel() * 4, d2h_bytes_per_call=cond_host.numel() * 4),
                      config=dict(workload=f'ConditionalFlowMatcherWrapper.sample dim{D} depth{args.depth} seq{SN} batch {SB}/GPU '
                                           f'midpoint (2 NFE/step), cond masked 70%, all {solver_steps} solver steps timed'))

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.depth == DEPTH:
            try:
                cpu, _ = cpu_reference_frames_per_sec(flash=True, steps=3, warmup=2)
            except Exception as ex:  # the baseline must never take the GPU number down with it
                cpu = dict(value=None, unit='frames/s', cores=os.cpu_count(), kind='port', sample=f'failed: {ex!r}')
        line = dict(metric=METRIC, value=value, unit='frames/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
                    data='synthetic',
                    config=dict(workload=f'VoiceBox dim{D} depth{args.depth} heads{HEADS} seq{N} (+{REG} register tokens) '
                                         f'batch {B}/GPU, CFM loss fwd+bwd+allreduce+clip+Adam', global_batch=B * world,
                                seq_len=N, parallelism=f'dp{world}', allreduce=args.allreduce, optimizer=args.optimizer, step_graph=graph_note, params=n_params, l2='inputs (268 MB/step) exceed the 126 MB L2',
                                peak_mem_gib=round(peak_mem, 1)),
                    eager_ms_per_step=(ms_eager / args.steps if ms_eager is not None else None),
                    e2e=dict(value=e2e_value, unit='frames/s', ms_per_step=ms_e2e / args.steps,
                             h2d_bytes_per_step=x_host.numel() * 4, d2h_bytes_per_step=4),
                    profile_only=bool(args.profile_only), gpu_launches=launches, roofline=roofline, kernels=kernels, sdpa_baseline=sdpa, clocks=clocks,
                    sample=sample, cpu_baseline=cpu)
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

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


CPU_SAMPLE_DEPTH = 2   # layers timed on the CPU; the full model has DEPTH of them


def cpu_reference_frames_per_sec(batch=1, steps=1, warmup=1):
    """fp32 forward+backward of the CFM loss through oracle/voicebox_oracle.py on all host threads, on a BOUNDED sample of the
    workload: same width / heads / sequence, batch 1, CPU_SAMPLE_DEPTH of the DEPTH transformer layers.  The per-layer cost
    dominates (embedding, conv and loss are < 3 % of a layer pair), CPU time is linear in depth and batch at these sizes, so
    frames/s of the full model = batch*SEQ / (t_sample * DEPTH / CPU_SAMPLE_DEPTH); the fixed parts are over-counted by that
    scaling, i.e. the CPU figure is slightly pessimistic.  (A full-depth step takes minutes on a 128-thread host.)"""
    from oracle import voicebox_oracle as O
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = oracle_state(depth=CPU_SAMPLE_DEPTH)
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'null_cond' not in k and 'inv_freq' not in k else v)
           for k, v in sd.items()}
    cfg = dict(depth=CPU_SAMPLE_DEPTH, heads=HEADS, num_register_tokens=REG, qk_norm=True, condition_on_text=False)
    x1 = torch.randn(batch, SEQ, DIM)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = O.cfm_loss(sdg, cfg, x1)
        loss.backward()
        dt = time.perf_counter() - t0
        for v in sdg.values():
            if v.grad is not None:
                v.grad = None
        if i >= warmup:
            times.append(dt)
        elif dt > 15.0:      # slow host: keep the sample bounded, count the (cold) first step instead of repeating it
            times.append(dt)
            warmup, steps = 0, 1
            break
    sec = statistics.median(times) * DEPTH / CPU_SAMPLE_DEPTH
    return dict(value=batch * SEQ / sec, unit='frames/s', cores=cores, kind='port',
                sample=f'oracle fp32 fwd+bwd, dim{DIM} seq{SEQ} batch {batch}, {CPU_SAMPLE_DEPTH} of {DEPTH} layers timed '
                       f'({statistics.median(times):.2f} s, {steps} step(s) after {warmup} warm-up) and scaled x{DEPTH // CPU_SAMPLE_DEPTH} '
                       f'in depth -> {sec:.1f} s per full step, threads={torch.get_num_threads()}'), sec


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    base, sec = cpu_reference_frames_per_sec(batch=1, steps=max(1, min(args.steps, 3)), warmup=1)
    line = dict(metric=METRIC, value=base['value'], unit='frames/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='fp32', data='synthetic',
                impl='reference',
                config=dict(workload=f'VoiceBox dim{DIM} depth{DEPTH} heads{HEADS} seq{SEQ} CFM train step; CPU sample: batch 1, '
                                     f'{CPU_SAMPLE_DEPTH}/{DEPTH} layers timed and scaled (see cpu_baseline.sample)'),
                cpu_baseline=base, e2e=dict(value=base['value'], unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


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
    ap.add_argument('--optimizer', default=os.environ.get('VBX_OPTIMIZER', 'torch'), choices=['torch', 'flat'],
                    help="'torch': torch.optim.Adam(fused=True) with the clip folded into its grad_scale; 'flat': "
                         "voicebox_pytorch_b200.FlatAdam = one vbx_adam_step launch over the flat buffers (experiment)")
    ap.add_argument('--allreduce', default=os.environ.get('VBX_ALLREDUCE', 'after'), choices=['overlap', 'after'],
                    help='gradient exchange: ONE all-reduce of the flat bucket after backward (default; measured faster: NCCL CTAs '
                         'otherwise take SMs from the 1-CTA/SM backward kernels), or chunked all-reduce overlapped with backward')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
    if args.optimizer == 'flat':
        opt = vbx.FlatAdam(bucket, lr=3e-4, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=0.5)
    else:
        opt = torch.optim.Adam([p for p in w.parameters() if p.requires_grad], lr=3e-4, betas=(0.9, 0.99), fused=True)
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

    # ---- device-resident timed region (value) + per-kernel CUDA-event timing -------------------------------------------
    tracked = ['vbx_attn_fwd', 'vbx_attn_bwd', 'vbx_adarms_fwd', 'vbx_adarms_bwd', 'vbx_geglu_fwd', 'vbx_geglu_bwd',
               'vbx_qkrope_fwd', 'vbx_qkrope_bwd', 'vbx_convpos_fwd', 'vbx_convpos_bwd']
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = vbx._lib.launch_count
    vbx._lib.profile_start(tracked)
    torch.cuda.nvtx.range_push('timed')   # `ncu --nvtx --nvtx-include "timed/"` captures exactly this region
    ms_dev = timed(lambda: train_step(x_dev), args.steps)
    torch.cuda.nvtx.range_pop()
    prof = vbx._lib.profile_stop()
    launches = vbx._lib.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---- end-to-end timed region: host batch in, loss out, every step --------------------------------------------------
    def e2e_step():
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
    work = {  # algorithmic FLOPs / bytes per launch
        'vbx_attn_fwd': ('tensor', attn_fl), 'vbx_attn_bwd': ('tensor', 2.5 * attn_fl),
        'vbx_adarms_fwd': ('hbm', T * D * 12.0), 'vbx_adarms_bwd': ('hbm', T * D * 16.0),
        'vbx_geglu_fwd': ('hbm', T * 2752 * 6.0), 'vbx_geglu_bwd': ('hbm', T * 2752 * 10.0),
        'vbx_qkrope_fwd': ('hbm', T * 2048 * 4.0), 'vbx_qkrope_bwd': ('hbm', T * 1024 * (2 + 2 + 2 + 4 + 2 + 2.0)),
        'vbx_convpos_fwd': ('hbm', B * N * D * 6.0), 'vbx_convpos_bwd': ('hbm', B * N * D * 10.0),
    }
    kernels = {}
    for name, (cnt, tot_ms) in prof.items():
        if cnt == 0:
            continue
        kind, units = work[name]
        avg_ms = tot_ms / cnt
        ach = units / (avg_ms * 1e-3) / (1e12 if kind == 'tensor' else 1e9)
        peak = peaks['tf_sustained'] if kind == 'tensor' else peaks['hbm']
        kernels[name] = dict(bound=kind, launches_per_step=cnt / args.steps, avg_us=avg_ms * 1e3, ms_per_step=tot_ms / args.steps,
                             achieved=ach, peak=peak, unit='TFLOP/s' if kind == 'tensor' else 'GB/s', frac=ach / peak)
    dom = max(kernels, key=lambda k: kernels[k]['ms_per_step']) if kernels else None
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
                        share_of_step=k['ms_per_step'] / (ms_dev / args.steps))

    # ---- sampling (BASELINE configs[3]) ---------------------------------------------------------------------------------
    sample = None
    if not args.no_sample:
        del opt
        bucket.flat = None
        for p in w.parameters():
            p.grad = None
        torch.cuda.empty_cache()
        SB, SN, solver_steps = 16, 2048, 4
        cond = torch.randn(SB, SN, D, device=dev)
        cmask = torch.zeros(SB, SN, dtype=torch.bool, device=dev)
        cmask[:, int(0.3 * SN):] = True
        w.sample(cond=cond, cond_mask=cmask, steps=2)            # warm-up: 1 solver step (weights cast + cached)
        w.sample(cond=cond, cond_mask=cmask, steps=2)
        ms_s = timed(lambda: w.sample(cond=cond, cond_mask=cmask, steps=solver_steps + 1), 1)
        sample = dict(metric='sample_ode_steps_per_sec', value=world * solver_steps / (ms_s / 1e3), unit='ODE-steps/s',
                      ms_per_ode_step=ms_s / solver_steps,
                      config=dict(workload=f'ConditionalFlowMatcherWrapper.sample dim{D} depth{args.depth} seq{SN} batch {SB}/GPU '
                                           f'midpoint (2 NFE/step), cond masked 70%, {solver_steps} timed solver steps of the 64'))

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.depth == DEPTH:
            try:
                cpu, _ = cpu_reference_frames_per_sec(batch=1, steps=1, warmup=1)
            except Exception as ex:  # the baseline must never take the GPU number down with it
                cpu = dict(value=None, unit='frames/s', cores=os.cpu_count(), kind='port', sample=f'failed: {ex!r}')
        line = dict(metric=METRIC, value=value, unit='frames/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
                    data='synthetic',
                    config=dict(workload=f'VoiceBox dim{D} depth{args.depth} heads{HEADS} seq{N} (+{REG} register tokens) '
                                         f'batch {B}/GPU, CFM loss fwd+bwd+allreduce+clip+Adam', global_batch=B * world,
                                seq_len=N, parallelism=f'dp{world}', allreduce=args.allreduce, optimizer=args.optimizer, params=n_params, l2='inputs (268 MB/step) exceed the 126 MB L2',
                                peak_mem_gib=round(peak_mem, 1)),
                    e2e=dict(value=e2e_value, unit='frames/s', ms_per_step=ms_e2e / args.steps,
                             h2d_bytes_per_step=x_host.numel() * 4, d2h_bytes_per_step=4),
                    profile_only=bool(args.profile_only), gpu_launches=launches, roofline=roofline, kernels=kernels, clocks=clocks, sample=sample, cpu_baseline=cpu)
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Development tool: torch.profiler over one cfg3 training step (all threads, fwd+bwd+optimizer) -> top CUDA kernels."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voicebox_pytorch_b200 as vbx  # noqa: E402
from voicebox_pytorch_b200.dist import FlatGradBucket  # noqa: E402
B = int(os.environ.get('SP_B', 64))
torch.manual_seed(0)
vb = vbx.VoiceBox(dim=1024, depth=24, heads=16, condition_on_text=False)
with torch.no_grad():
    for n, p in vb.named_parameters():
        if 'to_gamma.weight' in n or 'to_beta.weight' in n:
            p.normal_(0, 0.02)
w = vbx.ConditionalFlowMatcherWrapper(voicebox=vb).cuda()
bucket = FlatGradBucket(w)
opt = torch.optim.Adam([p for p in w.parameters() if p.requires_grad], lr=3e-4, betas=(0.9, 0.99), fused=True)
x = torch.randn(B, 1024, 1024, device='cuda')


def step():
    bucket.zero_grad()
    loss = w(x)
    loss.backward()
    bucket.finish()
    g = bucket.flat.norm()
    bucket.flat.mul_(torch.clamp(0.5 / (g + 1e-6), max=1.0))
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA'] if hasattr(prof.key_averages()[0], 'device_type') else prof.key_averages()
rows = sorted(prof.key_averages(), key=lambda e: -getattr(e, 'self_device_time_total', 0))
tot = sum(getattr(e, 'self_device_time_total', 0) for e in rows)
print(f'total self device time {tot / 1e3:.1f} ms')
for e in rows[:45]:
    t = getattr(e, 'self_device_time_total', 0)
    if t <= 0:
        continue
    print(f'{t / 1e3:9.2f} ms {100 * t / tot:5.1f}%  n={e.count:5d}  {e.key[:110]}')

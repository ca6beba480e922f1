"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from /root/reference via
oracle/ref_import.py) in fp32 on CPU.  Run in the build container only:

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these files ARE the pin for both
oracle/voicebox_oracle.py (tests/test_oracle.py) and the CUDA path (tests/test_gpu_*.py).  Zero-initialised
to_gamma/to_beta weights are perturbed (N(0, 0.02^2)) so the time-conditioning path is live (SURVEY.md hard part 4).
All RNG draws the reference makes internally are taken from the CPU generator; tensors a GPU test cannot re-draw
bit-identically (x0, times, masks, y0) are stored explicitly.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_import import import_reference  # noqa: E402

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def perturb_adaptive(model, seed=1):
    g = torch.Generator().manual_seed(seed)
    for name, p in model.named_parameters():
        if ('to_gamma.weight' in name) or ('to_beta.weight' in name):
            p.data.normal_(0, 0.02, generator=g)
        if name.endswith('q_norm.gamma') or name.endswith('k_norm.gamma') or name.endswith('final_norm.gamma'):
            p.data.add_(0.1 * torch.randn(p.shape, generator=g))


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB')


def voicebox_case(vp, name, *, dim, depth, heads, batch, seq, time_hidden_dim, sigma=0., qk_norm=True):
    torch.manual_seed(0)
    vb = vp.VoiceBox(dim=dim, depth=depth, dim_head=64, heads=heads, time_hidden_dim=time_hidden_dim,
                     num_cond_tokens=None, condition_on_text=False, attn_qk_norm=qk_norm)
    perturb_adaptive(vb)
    w = vp.ConditionalFlowMatcherWrapper(voicebox=vb, sigma=sigma)
    sd = {k: v.detach().clone() for k, v in w.state_dict().items()}

    torch.manual_seed(2)
    x1 = torch.randn(batch, seq, dim)
    torch.manual_seed(1000)
    x0 = torch.randn_like(x1)
    times = torch.rand((batch,))
    frac = torch.zeros((batch,)).float().uniform_(0.7, 1.0)
    rand = torch.zeros((batch,)).float().uniform_(0, 1)
    lengths = (frac * seq).long()
    start = ((seq - lengths) * rand).clamp(min=0)
    cond_mask = vp.mask_from_start_end_indices(seq, start, start + lengths)

    # ---- training loss through the reference's own VoiceBox.forward (vp.py:987-1115) with explicit draws ----
    t = times[:, None, None]
    wt = (1 - (1 - sigma) * t) * x0 + t * x1
    flow = x1 - (1 - sigma) * x0
    vb.train()
    vb.zero_grad()
    loss = vb(wt, times=times, target=flow, cond_token_ids=None, cond_mask=cond_mask, cond_drop_prob=0.)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in vb.named_parameters() if p.grad is not None}
    vb.eval()
    with torch.no_grad():
        # prediction (no target): cond passed explicitly = flow, exactly what the training call conditions on
        pred = vb(wt, times=times, cond=flow, cond_token_ids=None, cond_mask=cond_mask, cond_drop_prob=0.)

    # ---- the wrapper's own forward, RNG drawn inside (CPU generator): pins draw order ----
    torch.manual_seed(4242)
    loss_wrapper = w(x1)

    # ---- sampling (vp.py:1175-1330) : midpoint + euler, cond masked (first 30 % kept as prompt) ----
    torch.manual_seed(3)
    cond = torch.randn(batch, seq, dim)
    smask = torch.zeros(batch, seq, dtype=torch.bool)
    smask[:, int(0.3 * seq):] = True
    y0 = torch.randn(batch, seq, dim, generator=torch.Generator().manual_seed(77))
    samples = {}
    for method, steps in (('midpoint', 3), ('euler', 4)):
        w.odeint_kwargs['method'] = method
        real_randn_like = torch.randn_like
        torch.randn_like = lambda ref, **kw: y0.clone()  # inject y0: the CUDA generator cannot reproduce CPU draws
        try:
            samples[method] = w.sample(cond=cond, cond_mask=smask, steps=steps)
        finally:
            torch.randn_like = real_randn_like

    gsel = ['transformer.layers.0.3.to_qkv.weight', 'transformer.layers.1.2.to_gamma.weight',
            'transformer.layers.1.4.to_beta.bias',
            *(['transformer.layers.0.3.q_norm.gamma', 'transformer.layers.1.3.k_norm.gamma'] if qk_norm else []), 'transformer.layers.0.5.0.weight', 'transformer.layers.1.5.3.bias',
            'conv_embed.dw_conv1d.0.weight', 'conv_embed.dw_conv1d.0.bias', 'transformer.register_tokens',
            'sinu_pos_emb.0.weights', 'sinu_pos_emb.1.weight', 'to_pred.weight', 'to_embed.weight',
            'transformer.final_norm.gamma', 'transformer.layers.0.3.to_out.weight']
    arrays = {f'sd/{k}': v for k, v in sd.items()}
    arrays.update({f'grad/{k}': grads[k] for k in gsel})
    arrays.update(dict(x1=x1, x0=x0, times=times, frac=frac, rand=rand, cond_mask=cond_mask, loss=loss.detach(),
                       pred=pred, loss_wrapper_seed4242=loss_wrapper.detach(), cond=cond, sample_cond_mask=smask, y0=y0,
                       sample_midpoint_steps3=samples['midpoint'], sample_euler_steps4=samples['euler'],
                       cfg=np.array([dim, depth, heads, batch, seq, time_hidden_dim]), sigma=np.float32(sigma),
                       qk_norm=np.array(int(qk_norm))))
    save(name, **arrays)


def text_conditioned_case(vp, name, *, dim, depth, heads, batch, seq, n_tok, dim_emb, tok_len):
    """Text-conditioned VoiceBox (vp.py:1058-1070: to_cond_emb gather + interpolate_1d to the latent length) in eval mode, and
    classifier-free guidance through forward_with_cond_scale (vp.py:972-985; cond_drop_prob in {0,1} -> no RNG)."""
    torch.manual_seed(0)
    vb = vp.VoiceBox(dim=dim, depth=depth, dim_head=64, heads=heads, time_hidden_dim=dim, num_cond_tokens=n_tok,
                     dim_cond_emb=dim_emb, condition_on_text=True)
    perturb_adaptive(vb)
    vb.eval()
    sd = {k: v.detach().clone() for k, v in vb.state_dict().items()}
    torch.manual_seed(6)
    x = torch.randn(batch, seq, dim)
    cond = torch.randn(batch, seq, dim)
    times = torch.rand(batch)
    ids = torch.randint(0, n_tok, (batch, tok_len))
    cond_mask = torch.zeros(batch, seq, dtype=torch.bool)
    cond_mask[:, seq // 3:] = True
    with torch.no_grad():
        pred = vb(x, times=times, cond_token_ids=ids, cond=cond, cond_mask=cond_mask, cond_drop_prob=0.)
        guided = vb.forward_with_cond_scale(x, times=times, cond_token_ids=ids, cond=cond, cond_mask=cond_mask, cond_scale=1.3)
    arrays = {f'sd/{k}': v for k, v in sd.items()}
    arrays.update(dict(x=x, cond=cond, times=times, cond_token_ids=ids, cond_mask=cond_mask, pred=pred, guided=guided,
                       cfg=np.array([dim, depth, heads, batch, seq, n_tok, dim_emb, tok_len])))
    save(name, **arrays)


def duration_case(vp, name, *, dim, depth, heads, batch, seq, n_tok, dim_emb):
    torch.manual_seed(0)
    dp = vp.DurationPredictor(num_phoneme_tokens=n_tok, dim_phoneme_emb=dim_emb, dim=dim, depth=depth, heads=heads)
    g = torch.Generator().manual_seed(1)
    for n, p in dp.named_parameters():
        if n.endswith('gamma'):
            p.data.add_(0.1 * torch.randn(p.shape, generator=g))
    dp.eval()
    sd = {k: v.detach().clone() for k, v in dp.state_dict().items() if not k.startswith('aligner')}
    torch.manual_seed(5)
    cond = torch.randn(batch, seq, dim)
    ids = torch.randint(0, n_tok, (batch, seq))
    for b in range(batch):
        pad = int(torch.randint(0, seq // 4, (1,)))
        if pad:
            ids[b, seq - pad:] = -1
    cond_mask = torch.zeros(batch, seq, dtype=torch.bool)
    cond_mask[:, seq // 2:] = True
    with torch.no_grad():
        dur = dp(cond=cond, phoneme_ids=ids, cond_mask=cond_mask)
    arrays = {f'sd/{k}': v for k, v in sd.items()}
    arrays.update(dict(cond=cond, phoneme_ids=ids, cond_mask=cond_mask, durations=dur,
                       cfg=np.array([dim, depth, heads, batch, seq, n_tok, dim_emb])))
    save(name, **arrays)


def mask_kats(vp):
    """Bit-exact mask / index known answers (vp.py:68-74, 121-150)."""
    out = {}
    torch.manual_seed(1234)
    fl = torch.zeros(4).float().uniform_(0.7, 1.0)
    m = vp.mask_from_frac_lengths(1024, fl)
    out['kat1234_frac'] = fl
    out['kat1234_mask'] = m
    torch.manual_seed(7)
    out['kat7_prob_mask'] = vp.prob_mask_like((8,), 0.3, 'cpu')
    # a sweep with stored rand so the device path can be checked bit-exactly without sharing a generator
    g = torch.Generator().manual_seed(99)
    for seq in (17, 512, 1024, 2048):
        frac = torch.rand(64, generator=g) * 0.9 + 0.1
        rand = torch.rand(64, generator=g)
        lengths = (frac * seq).long()
        start = ((seq - lengths) * rand).clamp(min=0)
        out[f'sweep{seq}_frac'] = frac
        out[f'sweep{seq}_rand'] = rand
        out[f'sweep{seq}_mask'] = np.packbits(vp.mask_from_start_end_indices(seq, start, start + lengths).numpy(), axis=-1)
    rot = vp.RotaryEmbedding(64)(torch.tensor([-10000, 0, 1, 1023]))
    out['rotary_m10000_0_1_1023'] = rot
    out['geglu_1234'] = vp.GEGLU()(torch.tensor([[1., 2., 3., 4.]]))
    save('kats', **out)


if __name__ == '__main__':
    vp = import_reference()
    only = sys.argv[1:]
    if only == ['smooth']:
        # no qk-norm => softmax scale 1/8: a well-conditioned fixture on which bf16 noise is NOT amplified (tight tolerances)
        voicebox_case(vp, 'voicebox_d128_l2_h4_n200_noqknorm', dim=128, depth=2, heads=4, batch=2, seq=200, time_hidden_dim=128,
                      qk_norm=False)
        sys.exit(0)
    if only == ['text']:
        text_conditioned_case(vp, 'voicebox_text_d64_l2_h2_n120', dim=64, depth=2, heads=2, batch=2, seq=120, n_tok=20, dim_emb=32,
                              tok_len=45)
        sys.exit(0)
    mask_kats(vp)
    # N'=216 = one full 128-key tile + an 88-key tail; heads*64 != dim; F = int(128*8/3) = 341 (not 8-aligned)
    voicebox_case(vp, 'voicebox_d128_l2_h4_n200', dim=128, depth=2, heads=4, batch=2, seq=200, time_hidden_dim=128)
    # sigma > 0 and a second tile geometry (N' = 16+300 = 316 -> 2 full tiles + 60)
    voicebox_case(vp, 'voicebox_d64_l2_h2_n300_sigma', dim=64, depth=2, heads=2, batch=3, seq=300, time_hidden_dim=64,
                  sigma=0.1)
    duration_case(vp, 'durpred_d128_l2_h2_n100', dim=128, depth=2, heads=2, batch=3, seq=100, n_tok=50, dim_emb=64)

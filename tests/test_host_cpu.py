"""CPU (`-m "not gpu"`): the host-side logic and the C-ABI boundary -- state_dict contract, bit-exact mask generation,
exported symbols, loud failure without a GPU, and the flat-bucket gradient exchange on a 2-rank gloo group."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import voicebox_pytorch_b200 as vbx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'vbx.h')).read()
    declared = set(re.findall(r'^(?:int|const char\*)\s+(vbx_\w+)\s*\(', header, flags=re.M))
    assert len(declared) >= 18
    lib = ctypes.CDLL(vbx._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/vbx.h but not exported'
    assert declared == set(vbx._lib.EXPORTS)
    assert vbx._lib.load().vbx_version() == 100
    assert b'NULL' in vbx._lib.load().vbx_strerror(-1)


@pytest.mark.parametrize('name,dim,depth,heads,thd', [('voicebox_d128_l2_h4_n200', 128, 2, 4, 128),
                                                     ('voicebox_d64_l2_h2_n300_sigma', 64, 2, 2, 64)])
def test_state_dict_contract_voicebox(golden, name, dim, depth, heads, thd):
    """Parameter/buffer names and shapes identical to the reference's: checkpoints load both ways."""
    _, sd = golden(name)
    vb = vbx.VoiceBox(dim=dim, depth=depth, heads=heads, time_hidden_dim=thd, condition_on_text=False)
    w = vbx.ConditionalFlowMatcherWrapper(voicebox=vb)
    mine = w.state_dict()
    assert set(mine) == set(sd)
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
    w.load_state_dict(sd, strict=True)


def test_state_dict_contract_duration_predictor(golden):
    a, sd = golden('durpred_d128_l2_h2_n100')
    dp = vbx.DurationPredictor(num_phoneme_tokens=50, dim_phoneme_emb=64, dim=128, depth=2, heads=2)
    mine = dp.state_dict()
    assert set(mine) == set(sd)  # the fixture excludes the third-party aligner.* keys
    dp.load_state_dict(sd, strict=True)


def test_text_conditioned_voicebox_keys():
    vb = vbx.VoiceBox(dim=64, depth=2, heads=2, num_cond_tokens=10, dim_cond_emb=32, num_register_tokens=4)
    sd = vb.state_dict()
    assert sd['to_cond_emb.weight'].shape == (11, 32)
    assert sd['to_embed.weight'].shape == (64, 64 * 2 + 32)
    assert sd['transformer.register_tokens'].shape == (4, 64)
    assert sd['transformer.layers.0.2.to_gamma.weight'].shape == (64, 256)


def test_masks_bit_exact(golden):
    a, _ = golden('kats')
    torch.manual_seed(1234)
    fl = torch.zeros(4).float().uniform_(0.7, 1.0)
    assert torch.equal(vbx.mask_from_frac_lengths(1024, fl), a['kat1234_mask'])
    torch.manual_seed(7)
    assert torch.equal(vbx.prob_mask_like((8,), 0.3, 'cpu'), a['kat7_prob_mask'])
    for seq in (17, 512, 1024, 2048):
        frac, rand = a[f'sweep{seq}_frac'], a[f'sweep{seq}_rand']
        lengths = (frac * seq).long()
        start = ((seq - lengths) * rand).clamp(min=0)
        m = vbx.mask_from_start_end_indices(seq, start, start + lengths)
        assert np.array_equal(np.packbits(m.numpy(), axis=-1), a[f'sweep{seq}_mask'].numpy())


def test_unsupported_options_raise_at_construction():
    with pytest.raises(NotImplementedError):
        vbx.Transformer(64, depth=2, use_gateloop_layers=True)
    with pytest.raises(NotImplementedError):
        vbx.Transformer(64, depth=2, attn_dropout=0.1)
    with pytest.raises(NotImplementedError):
        vbx.Transformer(64, depth=2, dim_head=32)
    vb = vbx.VoiceBox(dim=64, depth=2, heads=2, condition_on_text=False)
    with pytest.raises(NotImplementedError):
        vbx.ConditionalFlowMatcherWrapper(voicebox=vb, use_torchode=True)
    with pytest.raises(NotImplementedError):
        vbx.ConditionalFlowMatcherWrapper(voicebox=vb, torchdiffeq_ode_method='dopri5')


def test_no_cpu_fallback():
    vb = vbx.VoiceBox(dim=64, depth=2, heads=2, condition_on_text=False)
    w = vbx.ConditionalFlowMatcherWrapper(voicebox=vb)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        w(torch.randn(2, 32, 64))


# ---- world_size-2 gloo: flat gradient bucket -------------------------------------------------------------------------
def _worker(rank, world, port, out, overlap=True):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from voicebox_pytorch_b200.dist import FlatGradBucket
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    model[3].weight.requires_grad_(True)
    unused = torch.nn.Linear(3, 3)  # never used by the loss (like duration_predictor inside the wrapper)
    model.add_module('unused', unused)
    bucket = FlatGradBucket(model, chunk_bytes=256, overlap=overlap)  # tiny chunks -> several collectives when overlapped
    bucket.broadcast_parameters(model)
    assert len(bucket.chunks) > 1
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(5, 8, generator=g) for _ in range(2)]
    # micro-step 1 under no_sync, micro-step 2 synced: gradient accumulation semantics of trainer.py:261-272
    bucket.zero_grad()
    with bucket.no_sync():
        model[3](model[2](model[1](model[0](xs[0])))).pow(2).mean().backward()
    model[3](model[2](model[1](model[0](xs[1])))).pow(2).mean().backward()
    bucket.finish()
    # strip the 16-byte alignment padding between parameters: compare the tightly packed gradients
    out[rank] = torch.cat([bucket.flat[o:o + p.numel()] for p, o in zip(bucket.params, bucket.offsets)])
    assert all(o % 4 == 0 for o in bucket.offsets) and bucket.flat.numel() % 4 == 0
    pad = torch.ones(bucket.flat.numel(), dtype=torch.bool)
    for p, o in zip(bucket.params, bucket.offsets):
        pad[o:o + p.numel()] = False
    assert float(bucket.flat[pad].abs().sum()) == 0          # padding stays zero through accumulation and all-reduce
    # local (un-reduced) reference
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    ref.load_state_dict({k: v for k, v in model.state_dict().items() if not k.startswith('unused')})
    for x in xs:
        ref(x).pow(2).mean().backward()
    out[world + rank] = torch.cat([p.grad.reshape(-1) for p in reversed(list(ref.parameters()))])
    dist.destroy_process_group()


@pytest.mark.parametrize('overlap', [True, False])
def test_flat_bucket_allreduce_gloo_world2(overlap):
    """both exchange modes: chunked all-reduce from the grad hooks (overlap) and ONE all-reduce of the bucket in finish()"""
    world = 2
    port = 29500 + (os.getpid() % 2000) + (7 if overlap else 0)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, overlap), nprocs=world, join=True)
    n_unused = 3 * 3 + 3
    mean_local = (out[world + 0] + out[world + 1]) / 2
    for r in range(world):
        flat = out[r]
        assert torch.allclose(flat[:n_unused], torch.zeros(n_unused))       # unused params first (reverse order), zero
        assert torch.allclose(flat[n_unused:], mean_local, atol=1e-6)         # mean over ranks of accumulated grads
    assert torch.equal(out[0], out[1])


def test_exp2_polynomial_constants_and_accuracy():
    """csrc/attn.cu:ex2_poly (VBX_EXP_POLY experiment) -- the constants in the kernel source are the ones fitted by
    tools/fit_exp2_poly.py, and its fp32 restatement stays within 1e-4 relative of exp2 over the softmax range (bf16 P
    carries 2^-9), including positive arguments (backward: s - lse can exceed 0 by rounding) and the -125 clamp."""
    import importlib.util
    import re
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('fit_exp2_poly', os.path.join(root, 'tools', 'fit_exp2_poly.py'))
    fp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fp)
    src = open(os.path.join(root, 'voicebox-pytorch_b200', 'csrc', 'attn.cu')).read()
    body = src[src.index('VBX_DEVINL float ex2_poly'):]
    body = body[:body.index('}')]
    consts = [float(v) for v in re.findall(r'(\d+\.\d+)f', body)]
    assert consts[0] == 125.0 and consts[1] == consts[2] == 12582912.0
    c3, c2, c1, c0 = consts[3:7]
    assert (c0, c1, c2, c3) == fp.COEFFS
    x = np.concatenate([np.linspace(-124.9, 3.0, 2_000_001), -np.abs(np.random.default_rng(1).normal(0, 4, 500_000))]).astype(np.float32)
    got = fp.ex2_poly_fp32(x)
    ref = np.exp2(x.astype(np.float64))
    assert np.max(np.abs(got / ref - 1)) < 1e-4
    assert fp.ex2_poly_fp32(np.array([-np.inf, -3e38], dtype=np.float32)).max() < 3e-38  # clamp: 2^-125, never NaN / garbage


def test_bf16_cast_cache_is_pinned_to_the_tensor_object():
    """ops.cast_bf16 caches the bf16 copy of a frozen parameter per (object, version, data_ptr).  id() values, device
    addresses and version counters are all reused once a model is freed, so an entry must never be served to another tensor
    object, and an in-place update must invalidate it."""
    import torch
    from voicebox_pytorch_b200 import ops
    ops.clear_cast_cache()
    p = torch.nn.Parameter(torch.randn(8, 8))
    with torch.no_grad():
        a = ops.cast_bf16(p)
        assert ops.cast_bf16(p) is a                       # hit
        p.add_(1.0)
        b = ops.cast_bf16(p)
        assert b is not a and torch.equal(b, p.detach().to(torch.bfloat16))   # version bump -> miss
        q = torch.nn.Parameter(torch.zeros(8, 8))
        # forge the collision a freed-and-rebuilt model produces: q's key, p's (live) entry with q's version and address
        ops._cast_cache[(id(q), '')] = (ops._cast_cache[(id(p), '')][0], q._version, q.data_ptr(), b)
        assert torch.equal(ops.cast_bf16(q).float(), torch.zeros(8, 8))
    assert ops.cast_bf16(p).requires_grad                  # grad mode: differentiable cast, never the cache
    ops.clear_cast_cache()


def test_batched_gamma_beta_matches_per_norm_linears():
    """ops.batched_affine (the VBX_BATCHED_GB experiment: all adaptive norms' gamma/beta from one batched GEMM) against
    the per-norm path it replaces (AdaptiveRMSNorm.gamma_beta -> ops.linear): same bf16 rounding points, so values agree to
    a bf16 ulp and every parameter / input gradient to bf16 accumulation noise.  Pure torch: runs on CPU."""
    import torch
    from voicebox_pytorch_b200 import ops
    from voicebox_pytorch_b200.modules import AdaptiveRMSNorm
    torch.manual_seed(0)
    B, C, D, K = 5, 48, 32, 3
    norms = [AdaptiveRMSNorm(D, cond_dim=C) for _ in range(K)]
    for n in norms:
        for p in n.parameters():
            torch.nn.init.normal_(p, 0, 0.3)
    cond = torch.randn(B, C)

    def run(batched):
        for n in norms:
            n.zero_grad()
        c = cond.clone().requires_grad_()
        cb = c.to(torch.bfloat16)
        if batched:
            gbs = ops.batched_affine(cb, [w for n in norms for w in (n.to_gamma.weight, n.to_beta.weight)],
                                     [b for n in norms for b in (n.to_gamma.bias, n.to_beta.bias)])
            outs = [t for t in gbs]
        else:
            outs = [t for n in norms for t in n.gamma_beta(cb)]
        assert all(t.dtype == torch.float32 and t.is_contiguous() and t.shape == (B, D) for t in outs)
        torch.manual_seed(1)
        loss = sum((t * torch.randn_like(t)).sum() for t in outs)
        loss.backward()
        grads = [p.grad.clone() for n in norms for p in n.parameters()]
        return [t.detach() for t in outs], grads, c.grad.clone()

    o1, g1, c1 = run(False)
    o2, g2, c2 = run(True)
    for a, b in zip(o1, o2):
        assert torch.allclose(a, b, rtol=2 ** -7, atol=1e-3)
    for a, b in zip(g1, g2):
        assert a.shape == b.shape and b.dtype == torch.float32
        assert (a - b).abs().max() <= 2e-2 * a.abs().max() + 1e-6
    assert (c1 - c2).abs().max() <= 2e-2 * c1.abs().max()
    with torch.no_grad():  # sampling path: no autograd node, same values
        gbs = ops.batched_affine(cond.to(torch.bfloat16), [w for n in norms for w in (n.to_gamma.weight, n.to_beta.weight)],
                                 [b for n in norms for b in (n.to_gamma.bias, n.to_beta.bias)])
        assert all(torch.equal(a, b) for a, b in zip(gbs, o2)) and not gbs[0].requires_grad


def test_flat_adam_repoints_storage_and_keeps_the_module_contract():
    """optim.FlatAdam moves every parameter into one flat fp32 buffer (same order as the gradient bucket) without
    changing parameter objects, values, shapes or state_dict keys; stepping without the CUDA library path must raise (no CPU
    optimizer fallback)."""
    import torch
    import voicebox_pytorch_b200 as vbx
    from voicebox_pytorch_b200.dist import FlatGradBucket
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    before = {k: v.clone() for k, v in m.state_dict().items()}
    ids = [id(p) for p in m.parameters()]
    bucket = FlatGradBucket(m)
    opt = vbx.FlatAdam(bucket, lr=1e-3, max_grad_norm=0.5)
    assert [id(p) for p in m.parameters()] == ids
    assert list(m.state_dict().keys()) == list(before.keys())
    assert all(torch.equal(v, before[k]) for k, v in m.state_dict().items())
    for p, off in zip(bucket.params, bucket.offsets):   # parameters and gradients: views at the same, 16-byte aligned offsets
        assert off % 4 == 0
        assert p.data_ptr() == opt.flat_p.data_ptr() + 4 * off and p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off
    assert bucket.offsets == [0, 4, 20, 28] and opt.flat_p.numel() == bucket.flat.numel() == 28 + 36   # sizes 3, 15, 5, 35
    m(torch.randn(4, 7)).sum().backward()   # autograd still accumulates into the bucket views
    assert bucket.flat.abs().sum() > 0
    with pytest.raises(RuntimeError):
        opt.step()                          # CPU tensors: the C ABI refuses, nothing is updated
    with pytest.raises(NotImplementedError):
        vbx.FlatAdam(bucket, weight_decay=0.01, decoupled=True)

"""CPU, build container only: oracle/voicebox_oracle.py against the UNMODIFIED reference imported read-only from
/root/reference (skipped where that tree does not exist, e.g. on the GPU box)."""
import os

import pytest
import torch

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/voicebox_pytorch'), reason='reference tree absent')


def test_readme_unconditional_snippet_loss_and_sample():
    from oracle.ref_import import import_reference
    from oracle import voicebox_oracle as O
    vp = import_reference()
    torch.manual_seed(0)
    m = vp.VoiceBox(dim=128, num_cond_tokens=500, depth=2, dim_head=64, heads=4, condition_on_text=False)
    w = vp.ConditionalFlowMatcherWrapper(voicebox=m)
    x = torch.randn(2, 96, 128)
    sd = {k: v.detach() for k, v in w.state_dict().items()}
    cfg = dict(depth=2, heads=4, num_register_tokens=16, qk_norm=True, condition_on_text=False)
    torch.manual_seed(5)
    ref = w(x)
    torch.manual_seed(5)
    mine = O.cfm_loss(sd, cfg, x)
    assert abs(ref.item() - mine.item()) <= 1e-6 * abs(ref.item())
    cond = torch.randn(2, 96, 128)
    torch.manual_seed(9)
    s_ref = w.sample(cond=cond, steps=3)
    torch.manual_seed(9)
    s_or = O.cfm_sample(sd, cfg, cond=cond, steps=3)
    assert (s_ref - s_or).abs().max().item() <= 1e-5


def test_patch_reference_rebinds_reference_objects():
    """patch_reference() on objects built by the REFERENCE's own classes (entry mode 1): identity, parameters and
    state_dict are untouched, the reference's beartype-checked constructors are satisfied, and forward/sample now route to the
    fused path -- which, on this GPU-less box, must fail loudly instead of falling back to the reference's torch code."""
    from oracle.ref_import import import_reference
    import voicebox_pytorch_b200 as vbx
    vp = import_reference()
    torch.manual_seed(0)
    m = vp.VoiceBox(dim=128, num_cond_tokens=500, depth=2, dim_head=64, heads=4, condition_on_text=False)
    w = vp.ConditionalFlowMatcherWrapper(voicebox=m)          # beartype: voicebox must be the reference's VoiceBox
    keys_before = list(w.state_dict().keys())
    ptrs_before = [p.data_ptr() for p in w.parameters()]
    assert vbx.patch_reference(w) is w
    assert type(w).__name__ == 'ConditionalFlowMatcherWrapper' and isinstance(w.voicebox, vp.VoiceBox)
    assert list(w.state_dict().keys()) == keys_before and [p.data_ptr() for p in w.parameters()] == ptrs_before
    x = torch.randn(2, 96, 128)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        w(x)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        w.sample(cond=x, steps=3)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        w.voicebox.transformer(torch.randn(2, 16, 128), adaptive_rmsnorm_cond=torch.randn(2, 512))
    # unsupported reference options are refused at patch time
    m2 = vp.VoiceBox(dim=128, num_cond_tokens=500, depth=2, dim_head=64, heads=4, condition_on_text=False, ff_dropout=0.1)
    with pytest.raises(NotImplementedError):
        vbx.patch_reference(m2)

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

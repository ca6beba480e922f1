"""Import the UNMODIFIED reference (lucidrains/voicebox-pytorch) read-only: from /root/reference in the build container,
else from baseline/_ref (the `pip install --no-deps --target baseline/_ref` copy of the task contract: git-ignored, it
travels to the GPU box with the snapshot).

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate oracle/voicebox_oracle.py and (b) generate
tests/golden/*.npz (tests/golden/make_golden.py); on the GPU box only by tests/test_gpu_patch_reference.py, which
builds REFERENCE objects and rebinds them with `patch_reference` (skipped, loudly, when no copy is present).
smoke() and bench.py never import this module.

The reference cannot be imported as shipped: seven third-party packages it imports at module scope are
absent here.  None of them is on the hot path, so they are replaced by inert
stubs in sys.modules.  `torchdiffeq.odeint` is on the sampling path (vp.py:1295) and is restated in
oracle/voicebox_oracle.py:odeint_fixed_grid (parity unpinned: torchdiffeq is un-vendored, un-pinned).
"""
import sys
import types
import warnings

import torch
from torch import nn

import os

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ('/root/reference', os.path.join(_REPO, 'baseline', '_ref'))


def reference_root():
    """First directory that holds the reference package, or None."""
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, 'voicebox_pytorch', 'voicebox_pytorch.py')):
            return root
    return None


REFERENCE_ROOT = reference_root() or '/root/reference'


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Inert(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise RuntimeError('inert stub of an out-of-scope third-party module was called')


def _klass(name):
    return type(name, (_Inert,), {})


def import_reference():
    """Returns the reference's `voicebox_pytorch.voicebox_pytorch` module (stubs installed once)."""
    if 'voicebox_pytorch.voicebox_pytorch' in sys.modules:
        return sys.modules['voicebox_pytorch.voicebox_pytorch']

    from oracle.voicebox_oracle import odeint_fixed_grid

    _mod('torchode', Tsit5=object, ODETerm=object, IntegralController=object,
         AutoDiffAdjoint=object, InitialValueProblem=object)
    _mod('torchdiffeq', odeint=odeint_fixed_grid)

    class Tokenizer:  # must be a class: used in a beartype annotation (vp.py:602)
        vocab_size = 256

    _mod('naturalspeech2_pytorch')
    _mod('naturalspeech2_pytorch.aligner', Aligner=_klass('Aligner'), ForwardSumLoss=_klass('ForwardSumLoss'),
         BinLoss=_klass('BinLoss'), maximum_path=lambda *a, **k: None)
    _mod('naturalspeech2_pytorch.utils')
    _mod('naturalspeech2_pytorch.utils.tokenizer', Tokenizer=Tokenizer)
    _mod('naturalspeech2_pytorch.naturalspeech2_pytorch', generate_mask_from_repeats=lambda *a, **k: None)
    _mod('audiolm_pytorch', EncodecWrapper=_klass('EncodecWrapper'), HubertWithKmeans=_klass('HubertWithKmeans'))
    _mod('spear_tts_pytorch', TextToSemantic=_klass('TextToSemantic'))
    _mod('gateloop_transformer', SimpleGateLoopLayer=_klass('SimpleGateLoopLayer'))
    _mod('vocos', Vocos=_klass('Vocos'))
    _mod('accelerate', Accelerator=object, DistributedType=object)
    _mod('accelerate.utils', DistributedDataParallelKwargs=object)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import voicebox_pytorch.voicebox_pytorch as vp
    return vp

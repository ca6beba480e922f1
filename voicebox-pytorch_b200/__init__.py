"""voicebox_pytorch_b200: B200-native (sm_100a) drop-in for the Voicebox conditional-flow-matching hot path of
lucidrains/voicebox-pytorch -- same class surface and state_dict, hand-written CUDA underneath."""
from .modules import (  # noqa: F401
    Transformer,
    VoiceBox,
    DurationPredictor,
    ConditionalFlowMatcherWrapper,
    AudioEncoderDecoder,
    Attention,
    RMSNorm,
    AdaptiveRMSNorm,
    MultiheadRMSNorm,
    ConvPositionEmbed,
    RotaryEmbedding,
    LearnedSinusoidalPosEmb,
    GEGLU,
    FeedForward,
    mask_from_frac_lengths,
    mask_from_start_end_indices,
    prob_mask_like,
)
from .patch import patch_reference  # noqa: F401
from .optim import FlatAdam  # noqa: F401
from . import ops, ode, modules, _lib  # noqa: F401

__all__ = ['Transformer', 'VoiceBox', 'DurationPredictor', 'ConditionalFlowMatcherWrapper', 'AudioEncoderDecoder',
           'patch_reference']

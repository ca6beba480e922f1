"""ctypes binding of libvbx_sm100a.so (C ABI declared in include/vbx.h).

The library is built in-tree by `__graft_entry__.build()` / `csrc/build.sh` into `voicebox-pytorch_b200/lib/`.  There is no
fallback of any kind: if the shared object is missing, or a call is made without a CUDA device, this module raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('VBX_LIB') or os.path.join(_HERE, 'lib', 'libvbx_sm100a.so')  # VBX_LIB: trace build (tools/)

_i64, _f32, _int, _vp = ctypes.c_int64, ctypes.c_float, ctypes.c_int, ctypes.c_void_p

# name -> argtypes (all return int); mirrors include/vbx.h one to one
_SIGNATURES = {
    'vbx_adarms_fwd': [_vp, _i64, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_adarms_bwd': [_vp, _i64, _i64, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_geglu_fwd': [_vp, _vp, _i64, _i64, _vp],
    'vbx_geglu_bwd': [_vp, _vp, _vp, _vp, _i64, _i64, _vp],
    'vbx_convpos_fwd': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp],
    'vbx_convpos_bwd': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp],
    'vbx_cfm_embed': [_vp, _vp, _vp, _vp, _f32, _vp, _i64, _i64, _i64, _vp],
    'vbx_embed_concat': [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_masked_mse_fwd': [_vp, _vp, _vp, _vp, _f32, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_masked_mse_bwd': [_vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_ode_axpy': [_vp, _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_qkrope_fwd': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_qkrope_bwd': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_attn_fwd': [_vp, _vp, _vp, _i64, _i64, _vp, _f32, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_attn_bwd': [_vp, _vp, _vp, _i64, _i64, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64,
                     _vp],
    'vbx_adam_step': [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _int, _i64, _vp, _vp, _vp],
    'vbx_pack_bf16': [_vp, _vp, _i64, _i64, _vp],
    'vbx_accum_bf16_2d': [_vp, _i64, _vp, _i64, _i64, _i64, _vp],
    'vbx_accum_bf16_table': [_vp, _vp, _i64, _i64, _vp],
    'vbx_gemm_bf16': [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_ff1_geglu': [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_ff2_dgrad_geglu_bwd': [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    'vbx_umma_selftest': [_vp, _vp, _vp, _int, _vp],
}

EXPORTS = tuple(_SIGNATURES) + ('vbx_version', 'vbx_strerror')

_lib = None


def load():
    """dlopen the library (once).  Raises a RuntimeError naming the build command if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           f'(or voicebox-pytorch_b200/csrc/build.sh). There is no fallback path.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _int
    lib.vbx_version.restype = _int
    lib.vbx_strerror.restype = ctypes.c_char_p
    lib.vbx_strerror.argtypes = [_int]
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Tensors handed to the C ABI must be contiguous CUDA tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('voicebox_pytorch_b200 kernels run on CUDA tensors only (no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError('non-contiguous tensor passed to the vbx C ABI')
    if t.device.index != torch.cuda.current_device():
        # launches go to torch's current stream of the CURRENT device: a foreign pointer there is an illegal address at best
        raise RuntimeError(f'tensor on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}: '
                           f'wrap the call in torch.cuda.device(...) (one process per GPU is the supported layout)')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# ---- instrumentation used by bench.py (off by default: zero cost beyond one `is None` test) ----------------------------
launch_count = 0          # kernels launched through the C ABI since import (vbx_attn_bwd = 2 launches)
_profile = None           # {entry point name: [(start_event, end_event), ...]} while bench.py times selected kernels


def profile_start(names):
    global _profile
    _profile = {n: [] for n in names}


def profile_stop():
    """-> {name: (launches, total_ms)}; call after torch.cuda.synchronize()."""
    global _profile
    out = {n: (len(ev), sum(a.elapsed_time(b) for a, b in ev)) for n, ev in (_profile or {}).items()}
    _profile = None
    return out


def call(name, *args):
    global launch_count
    lib = load()
    launch_count += 2 if name == 'vbx_attn_bwd' else 1
    if _profile is not None and name in _profile:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        _profile[name].append((e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name}: {lib.vbx_strerror(rc).decode()} (rc={rc})')

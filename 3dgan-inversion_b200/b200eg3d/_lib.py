"""ctypes binding of libb200eg3d.so -- the C-ABI boundary of the CUDA hot path (include/b200eg3d.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('B200EG3D_LIB') or os.path.join(_HERE, 'libb200eg3d.so')     # env override: kernel-tuning variants

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float

# b200_version() this binding table was written for.  Bumped together with csrc/conv_api.cu whenever a prototype changes: a
# stale or variant .so (B200EG3D_LIB) with other argument lists would otherwise be called with the wrong stack layout.
EXPECTED_VERSION = 207

# name -> argument ctypes (every function returns int status; 0 = ok)
SIGNATURES = {
    'b200_conv_fwd': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'b200_conv_dgrad': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'b200_conv_wgrad': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'b200_bank_styles_fwd': [_P, _I, _P, _I, _I, _I, _P],
    'b200_bank_weights_fwd': [_P, _I, _I, _P],
    'b200_bank_weights_bwd': [_P, _I, _I, _P],
    'b200_bank_styles_bwd': [_P, _I, _P, _P, _I, _I, _I, _P],
    'b200_split_bf16': [_P, _P, _P, _L, _P],
    'b200_conv_fwd_tc': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'b200_conv_fwd_tc_act': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P],
    'b200_conv_dgrad_tc': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'b200_conv_wgrad_tc': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'b200_adam_step': [_P, _I, _P, _F, _F, _F, _F, _F, _P, _P, _P],
    'b200_modconv_weight_prep': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200_modconv_weight_prep_bwd': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200_bias_act': [_P, _P, _P, _P, _P, _P, _I, _L, _L, _I, _I, _F, _F, _F, _P],
    'b200_layer_act_fwd': [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _F, _F, _F, _P],
    'b200_layer_act_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P],
    'b200_layer_act_bwd_sum2': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P],
    'b200_layer_act_bwd_sum2_supported': [_I, _I, _I, _I, _L],
    'b200_conv1x1_wgrad_split': [_P, _P, _P, _P, _I, _L, _I, _I, _P],
    'b200_conv1x1_fwd_thin': [_P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P],
    'b200_upfirdn2d': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    'b200_upfirdn2d_fused': [_P] * 6 + [_I] * 13 + [_F, _I, _P, _P, _P, _L, _I, _F, _F, _F, _I, _P],
    'b200_triplane_mlp_fwd': [_P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _L, _F, _P, _P, _P, _P, _F, _P, _P, _P, _P],
    'b200_triplane_mlp_bwd': [_P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _L, _F, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P],
    'b200_pti_loss_fwd': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _P, _P],
    'b200_pti_loss_bwd': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    'b200_warp_uv_fwd': [_P, _P, _P, _P, _P, _I, _P, _P, _P],
    'b200_warp_uv_bwd': [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P],
    'b200_noise_pyramid_fwd': [_I, _P, _P, _P, _P, _P, _P, _P],
    'b200_noise_pyramid_bwd': [_I, _P, _P, _P, _P, _P, _P, _P, _P],
    'b200_noise_normalize': [_I, _P, _P, _P, _P],
    'b200_ray_sampler_fwd': [_P, _P, _I, _I, _P, _P, _P],
    'b200_ray_sampler_bwd': [_P, _P, _I, _I, _P, _P, _P, _P],
    'b200_ray_depths_coarse': [_P, _P, _P, _L, _I, _F, _P, _P],
    'b200_depth_minmax': [_P, _L, _P, _P],
    'b200_ray_importance': [_P, _P, _P, _P, _L, _I, _I, _P],
    'b200_ray_composite_fwd': [_P, _P, _P, _I, _P, _P, _P, _I, _P, _I, _L, _P, _P, _P, _P],
    'b200_ray_composite_bwd': [_P, _P, _P, _I, _P, _P, _P, _I, _P, _I, _L, _P, _P, _P, _P, _P, _P, _P, _P],
}



class BankLayer(ctypes.Structure):
    """B200BankLayer of include/b200eg3d.h."""
    _fields_ = [(k, ctypes.c_void_p) for k in ('affine_w', 'affine_b', 'weight', 'styles', 'dcoef', 'wmod', 'w_hi', 'w_lo', 'dwmod',
                                               'd_weight', 'd_styles', 'd_affine_w', 'd_affine_b')] + \
               [(k, ctypes.c_int) for k in ('widx', 'cin', 'cout', 'taps', 'demod')] + [('post_scale', ctypes.c_float)]


_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: run `python __graft_entry__.py build` (no CPU/PyTorch fallback exists)')
    lib = ctypes.CDLL(LIB_PATH)
    lib.b200_version.restype = ctypes.c_int
    lib.b200_version.argtypes = []
    if lib.b200_version() != EXPECTED_VERSION:
        raise RuntimeError(f'{LIB_PATH} reports version {lib.b200_version()}, this binding expects {EXPECTED_VERSION}: '
                           'rebuild with `python __graft_entry__.py build`')
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int
    lib.b200_last_error.restype = ctypes.c_char_p
    lib.b200_last_error.argtypes = []
    lib.b200_version.restype = ctypes.c_int
    lib.b200_version.argtypes = []
    lib.b200_set_pdl.restype = ctypes.c_int
    lib.b200_set_pdl.argtypes = [_I]
    lib.b200_set_conv_pair.restype = ctypes.c_int
    lib.b200_set_conv_pair.argtypes = [_I]
    lib.b200_set_mlp_passes.restype = ctypes.c_int
    lib.b200_set_mlp_passes.argtypes = [_I]
    lib.b200_set_triplane_impl.restype = ctypes.c_int
    lib.b200_set_triplane_impl.argtypes = [_I]
    lib.b200_conv_tc_supported.restype = ctypes.c_int
    lib.b200_triplane_bwd_workspace_bytes.restype = ctypes.c_long
    lib.b200_triplane_bwd_workspace_bytes.argtypes = [_I, _L]
    lib.b200_triplane_fsave_bytes.restype = ctypes.c_long
    lib.b200_triplane_fsave_bytes.argtypes = [_I, _L]
    lib.b200_conv_tc_supported.argtypes = [_I] * 7
    lib.b200_conv1x1_thin_supported.restype = ctypes.c_int
    lib.b200_conv1x1_thin_supported.argtypes = [_I, _I]
    lib.b200_conv1x1_fwd_thin_supported.restype = ctypes.c_int
    lib.b200_conv1x1_fwd_thin_supported.argtypes = [_I, _I]
    lib.b200_conv_tc_ksplit.restype = ctypes.c_int
    lib.b200_conv_tc_ksplit.argtypes = [_I] * 8
    lib.b200_conv_tc_act_fusable.restype = ctypes.c_int
    lib.b200_conv_tc_act_fusable.argtypes = [_I] * 6
    lib.b200_noise_pyramid_work_floats.restype = ctypes.c_long
    lib.b200_noise_pyramid_work_floats.argtypes = [_I, _P]
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be a contiguous fp32/int32 CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('b200eg3d: tensor is not on a CUDA device (there is no CPU path)')
    if not t.is_contiguous():
        raise RuntimeError('b200eg3d: tensor must be contiguous')
    return t.data_ptr()


def stream(device=None):
    """Raw handle of torch's current stream on `device` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """Make the device of `t` current for the enclosed C-ABI calls and restore the previous one afterwards (the reference's
    plugins do this with OptionalCUDAGuard(device_of(x)), bias_act.cpp:58).  A no-op when it already is current."""

    def __init__(self, t):
        self.idx = t.device.index if (t is not None and t.is_cuda) else None
        self.prev = None

    def __enter__(self):
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def device_guard(fn):
    """Decorator for autograd.Function.forward / backward: run with the device of the first CUDA tensor argument current, so
    that kernels, torch.empty(...) allocations and stream() all refer to the tensors' device even when another one is current."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = next((a for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if t is None and args and hasattr(args[0], 'saved_tensors'):
            try:
                t = next((a for a in args[0].saved_tensors if isinstance(a, torch.Tensor) and a.is_cuda), None)
            except Exception:
                t = None
        with on_device(t):
            return fn(*args, **kwargs)
    return wrapper


LAUNCHES = 0        # number of C-ABI kernel calls made (bench.py reads it for `gpu_launches`)
PROFILE = None      # when a dict: name -> [(start_event, end_event), ...] recorded around every call (bench.py roofline leg)


def call(name, *args):
    global LAUNCHES
    lib = load()
    LAUNCHES += 1
    if PROFILE is not None:
        # per-kernel timing leg of bench.py: drain the device first so that the events bracket this call's kernels only
        # (no overlap with earlier asynchronous work, no second-stream branch running beside it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed: {lib.b200_last_error().decode()}')

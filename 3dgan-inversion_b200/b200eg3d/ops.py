"""Host-side operators of the EG3D hot path: thin autograd wrappers over the C-ABI CUDA library.

Public names mirror the reference's operator layer so that its call sites read the same:
    bias_act(x, b, dim, act, alpha, gain, clamp)          torch_utils/ops/bias_act.py:54
    upfirdn2d / upsample2d / setup_filter                  torch_utils/ops/upfirdn2d.py:72,120,315
    modulated_conv2d-based layers                          training/networks_stylegan2.py:34-91,311-357
Internally the synthesis stack keeps activations NHWC ([N,H,W,C] fp32): one channel vector per pixel is the
K-contiguous GEMM operand and the coalescing unit of every elementwise kernel.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import call, device_guard, ptr, stream

_ACT = {'linear': (1, 0.0, 1.0), 'relu': (2, 0.0, math.sqrt(2)), 'lrelu': (3, 0.2, math.sqrt(2)), 'tanh': (4, 0.0, 1.0),
        'sigmoid': (5, 0.0, 1.0), 'elu': (6, 0.0, 1.0), 'selu': (7, 0.0, 1.0), 'softplus': (8, 0.0, 1.0),
        'swish': (9, 0.0, math.sqrt(2))}
_ACT_REF = {'linear': '', 'relu': 'y', 'lrelu': 'y', 'tanh': 'y', 'sigmoid': 'y', 'elu': 'y', 'selu': 'y', 'softplus': 'y',
            'swish': 'x'}


# Numerics of the convolution stack.  tc: use the tcgen05 kernels when the shape allows; *_passes: 3 = split-bf16
# (hi*hi + hi*lo + lo*hi, ~fp32 accuracy), 1 = plain bf16 operands.  B200EG3D_TC=0 forces the exact-fp32 SIMT kernels.
# Defaults: forward and dgrad 3-pass (their errors would compound through 17 layers and fail the 1e-3 bar), wgrad 1-pass: a weight
# gradient is a sum over all pixels of products whose bf16 rounding errors (2^-9 relative, zero-mean) are independent, so the sum
# keeps ~1e-3 relative accuracy -- measured per parameter by tests/test_gpu_golden.py (channel norms and strided samples of every
# gradient tensor, worst 2.7e-3 against the 1e-2 bar).  B200EG3D_WGRAD_PASSES=3 buys fp32-equivalent weight gradients for ~0.5 ms.
CONFIG = {'tc': os.environ.get('B200EG3D_TC', '1') != '0',
          'fwd_passes': int(os.environ.get('B200EG3D_FWD_PASSES', '3')),
          'bank': os.environ.get('B200EG3D_BANK', '1') != '0',       # batch styles + weight prep of all layers into one launch
          'dgrad_passes': int(os.environ.get('B200EG3D_DGRAD_PASSES', '3')),
          'wgrad_passes': int(os.environ.get('B200EG3D_WGRAD_PASSES', '1')),
          # run the ToRGB / skip-upsample chain on a second stream, concurrently with the next block's convolutions
          'overlap': os.environ.get('B200EG3D_OVERLAP', '1') != '0',
          # walk 8x16-pixel ray patches front to back in the tri-plane kernels instead of ray after ray (measured slower: see DESIGN.md)
          'ray_patch_order': os.environ.get('B200EG3D_RAY_PATCH_ORDER', '0') != '0',
          # weight-gradient GEMMs on their own stream (they only feed the bank's backward): their split-K reduction tails and launch
          # latencies overlap the dgrad chain; d wmod of all layers is one pool, zeroed once per step off the critical path
          'wgrad_stream': os.environ.get('B200EG3D_WGRAD_STREAM', '1') != '0',
          # inside a synthesis network keep activations only as the split-bf16 pair the next tensor-core conv reads: no fp32 copy is
          # written by the layer epilogues or re-read by the activation backward (the returned fp32 tensor is then a placeholder)
          'lean_acts': os.environ.get('B200EG3D_LEAN_ACTS', '1') != '0',
          # styles + modulated weights of the super-resolution module on the second stream, beside the renderer (forward and backward)
          'early_sr_bank': os.environ.get('B200EG3D_EARLY_SR_BANK', '1') != '0',
          # an activation with two consumers hands each its own handle: the two gradients are summed in the activation backward
          'fork_grads': os.environ.get('B200EG3D_FORK_GRADS', '1') != '0',
          # one zero fill per network pass for the outputs of all split-K convolutions instead of a memset before each launch
          'zero_pools': os.environ.get('B200EG3D_ZERO_POOLS', '1') != '0',
          'ranges': os.environ.get('B200EG3D_RANGES', '0') != '0'}


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class _NullRange:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL_RANGE = _NullRange()


def prof_range(name):
    """torch.autograd.profiler.record_function(name) with the range names the reference emits (misc.profiled_function,
    torch_utils/misc.py:102-107: 'modulated_conv2d', 'conv2d_resample', ...; networks_stylegan2.py:236-262,505: 'input',
    'broadcast', 'truncate', 'split_ws'), so traces of the two implementations line up.  Free when no profiler is recording
    (and in graph replay, where no Python runs at all); B200EG3D_RANGES=1 forces the ranges on (NVTX via emit_nvtx)."""
    if CONFIG['ranges'] or torch.autograd.profiler._is_profiler_enabled:
        return torch.autograd.profiler.record_function(name)
    return _NULL_RANGE


def _tc_ok(kind, h, w, cin, cout, k, up):
    return CONFIG['tc'] and _lib.load().b200_conv_tc_supported(kind, h, w, cin, cout, k, up) == 1


def _split(t, need_lo):
    """fp32 tensor -> (hi, lo) bf16 tensors with t ~= hi + lo (lo None when not needed)."""
    hi = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    lo = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16) if need_lo else None
    call('b200_split_bf16', ptr(t), ptr(hi), ptr(lo), t.numel(), stream())
    return hi, lo


def _conv_fwd(x, wmod, y, n, h, w, cin, cout, k, up):
    if _tc_ok(0, h, w, cin, cout, k, up):
        npass = CONFIG['fwd_passes']
        xh, xl = _split(x, npass == 3)
        wh, wl = _split(wmod, npass == 3)
        call('b200_conv_fwd_tc', ptr(xh), ptr(xl), ptr(wh), ptr(wl), ptr(y), n, h, w, cin, cout, k, up, npass, 0, stream())
    else:
        call('b200_conv_fwd', ptr(x), ptr(wmod), ptr(y), n, h, w, cin, cout, k, up, stream())


def _conv_dgrad(dy, wmod, dx, n, h, w, cin, cout, k, up):
    if _tc_ok(1, h, w, cin, cout, k, up):
        npass = CONFIG['dgrad_passes']
        dh, dl = _split(dy, npass == 3)
        wh, wl = _split(wmod, npass == 3)
        call('b200_conv_dgrad_tc', ptr(dh), ptr(dl), ptr(wh), ptr(wl), ptr(dx), n, h, w, cin, cout, k, up, npass, 0, stream())
    else:
        call('b200_conv_dgrad', ptr(dy), ptr(wmod), ptr(dx), n, h, w, cin, cout, k, up, stream())


def _conv_wgrad(x, dy, dwmod, n, h, w, cin, cout, k, up):
    if _tc_ok(2, h, w, cin, cout, k, up):
        npass = CONFIG['wgrad_passes']
        xh, xl = _split(x, npass == 3)
        dh, dl = _split(dy, npass == 3)
        call('b200_conv_wgrad_tc', ptr(xh), ptr(xl), ptr(dh), ptr(dl), ptr(dwmod), n, h, w, cin, cout, k, up, npass, 0, stream())
    else:
        call('b200_conv_wgrad', ptr(x), ptr(dy), ptr(dwmod), n, h, w, cin, cout, k, up, stream())


# ----------------------------------------------------------------------------------------------
# bias_act (generic, any layout; reference plugin signature bias_act.cpp:36)

class _BiasAct(torch.autograd.Function):
    @staticmethod
    @device_guard
    def forward(ctx, x, b, dim, act, alpha, gain, clamp):
        code = _ACT[act][0]
        xc = _f32c(x)
        y = torch.empty_like(xc)
        step = int(np.prod(xc.shape[dim + 1:])) if xc.ndim > dim + 1 else 1
        bc = _f32c(b) if b is not None else None
        call('b200_bias_act', ptr(xc), ptr(bc), None, None, None, ptr(y), 0, xc.numel(), step,
             xc.shape[dim] if bc is not None else 1, code, alpha, gain, clamp, stream())
        ctx.cfg = (dim, act, alpha, gain, clamp, step, b is not None)
        ctx.save_for_backward(xc if _ACT_REF[act] == 'x' else None, bc if _ACT_REF[act] == 'x' else None,
                              y if _ACT_REF[act] == 'y' or clamp >= 0 else None)
        return y

    @staticmethod
    @device_guard
    def backward(ctx, dy):
        dim, act, alpha, gain, clamp, step, has_b = ctx.cfg
        xref, bref, yref = ctx.saved_tensors
        dyc = _f32c(dy)
        dx = torch.empty_like(dyc)
        call('b200_bias_act', ptr(dyc), ptr(bref), ptr(xref), ptr(yref), None, ptr(dx), 1, dyc.numel(), step,
             dyc.shape[dim] if bref is not None else 1, _ACT[act][0], alpha, gain, clamp, stream())
        db = dx.sum([i for i in range(dx.ndim) if i != dim]) if has_b else None
        return dx, db, None, None, None, None, None


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain).  Same argument meaning as the reference op; CUDA only."""
    if act not in _ACT:
        raise ValueError(f'unknown activation {act!r}')
    _, a0, g0 = _ACT[act]
    alpha = float(alpha if alpha is not None else a0)
    gain = float(gain if gain is not None else g0)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None and (b.ndim != 1 or b.shape[0] != x.shape[dim]):
        raise ValueError('bias must be a vector matching x.shape[dim]')
    with prof_range('bias_act'):
        return _BiasAct.apply(x, b, dim, act, alpha, gain, clamp)


# ----------------------------------------------------------------------------------------------
# upfirdn2d

def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Same contract as upfirdn2d.setup_filter (upfirdn2d.py:72-116)."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _filter2d(f, device):
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    f = f.detach().to(device=device, dtype=torch.float32)
    if f.ndim == 1:
        f = f.ger(f)                       # separable filters are applied as their outer product
    return f.contiguous()


def _pair(v):
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _pad4(p):
    if isinstance(p, int):
        return p, p, p, p
    p = [int(v) for v in p]
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    return tuple(p)


def _upfirdn_nhwc_raw(x, f2, up, down, pad, flip, gain, add=None):
    """x [N,H,W,C] contiguous fp32 -> [N,OH,OW,C]; no autograd."""
    n, h, w, c = x.shape
    (upx, upy), (dx, dy) = up, down
    px0, px1, py0, py1 = pad
    fh, fw = f2.shape
    oh = (h * upy + py0 + py1 - fh + dy) // dy
    ow = (w * upx + px0 + px1 - fw + dx) // dx
    y = torch.empty([n, oh, ow, c], device=x.device, dtype=torch.float32)
    call('b200_upfirdn2d', ptr(x), ptr(f2), ptr(add), ptr(y), n, h, w, c, fh, fw, upx, upy, dx, dy, px0, px1, py0, py1,
         int(flip), float(gain), stream())
    return y


class _UpfirdnNHWC(torch.autograd.Function):
    @staticmethod
    @device_guard
    def forward(ctx, x, f2, up, down, pad, flip, gain):
        xc = _f32c(x)
        ctx.cfg = (up, down, pad, flip, gain, xc.shape)
        ctx.save_for_backward(f2)
        return _upfirdn_nhwc_raw(xc, f2, up, down, pad, flip, gain)

    @staticmethod
    @device_guard
    def backward(ctx, dy):
        (f2,) = ctx.saved_tensors
        up, down, pad, flip, gain, xs = ctx.cfg
        n, ih, iw, c = xs
        oh, ow = dy.shape[1:3]
        fh, fw = f2.shape
        px0, px1, py0, py1 = pad
        # upfirdn2d.py:257-268: the adjoint is the same op with up/down swapped and the filter flipped
        p = (fw - px0 - 1, iw * up[0] - ow * down[0] + px0 - up[0] + 1, fh - py0 - 1, ih * up[1] - oh * down[1] + py0 - up[1] + 1)
        dx = _upfirdn_nhwc_raw(_f32c(dy), f2, down, up, p, not flip, gain)
        return dx, None, None, None, None, None, None


def upfirdn2d_nhwc(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    f2 = _filter2d(f, x.device)
    return _UpfirdnNHWC.apply(x, f2, _pair(up), _pair(down), _pad4(padding), bool(flip_filter), float(gain))


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """NCHW front-end with the reference's signature (upfirdn2d.py:120): every (n, c) image is one single-channel map."""
    n, c, h, w = x.shape
    with prof_range('upfirdn2d'):
        y = upfirdn2d_nhwc(x.reshape(n * c, h, w, 1), f, up, down, padding, flip_filter, gain)
    return y.reshape(n, c, y.shape[1], y.shape[2])


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """upfirdn2d.py:315-350."""
    upx, upy = _pair(up)
    px0, px1, py0, py1 = _pad4(padding)
    fh, fw = (1, 1) if f is None else ((f.shape[0], f.shape[0]) if f.ndim == 1 else (f.shape[0], f.shape[1]))
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


# ----------------------------------------------------------------------------------------------
# Modulated convolution layer (SynthesisLayer body) on NHWC activations

def _fir4(device):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64)
    f = torch.outer(k, k)
    return (f / f.sum()).to(device=device, dtype=torch.float32).contiguous()


_FIR_CACHE = {}


def fir_filter(device):
    key = str(device)
    if key not in _FIR_CACHE:
        _FIR_CACHE[key] = _fir4(device)
    return _FIR_CACHE[key]


# ----------------------------------------------------------------------------------------------
# Weight bank: styles + modulated weights of every layer of a synthesis network in one launch per stage

class BankSpec:
    """Static description of one modulated-conv layer for the bank (built by the generator modules)."""

    def __init__(self, affine, weight, widx, demod, post_scale, h, w, up):
        self.affine, self.weight, self.widx, self.demod, self.post_scale = affine, weight, widx, bool(demod), float(post_scale)
        cout, cin, k, _ = weight.shape
        self.cin, self.cout, self.k, self.taps, self.h, self.w, self.up = cin, cout, k, k * k, h, w, up
        tc_f = _tc_ok(0, h, w, cin, cout, k, up)
        tc_b = _tc_ok(1, h, w, cin, cout, k, up) and _tc_ok(2, h, w, cin, cout, k, up)
        if demod:                       # 3x3 layers: all-or-nothing, as in _ModConvLayer
            tc_f = tc_b = tc_f and tc_b
        self.tc_f, self.tc_b = tc_f, tc_b


class WeightBank:
    """Per-forward container of the bank's outputs (and, during backward, of the per-layer d wmod)."""

    def __init__(self, specs):
        self.specs = specs
        n = len(specs)
        self.styles, self.dcoef, self.wmod, self.w_hi, self.w_lo = [None] * n, [None] * n, [None] * n, [None] * n, [None] * n
        self.dwmod = [None] * n
        self.need_wgrad = False
        self.token = None
        self.ztotal = 0
        self.zpool = None          # one zero-filled buffer for the small backward accumulators (d bias, d noise_strength) of all layers
        self.zoff = []
        self.side = None           # second stream some layers ran on (CONFIG['overlap']); the bank's backward joins it
        # outputs of the split-K convolutions (the 4x4 .. 32x32 blocks) start from zero: two pools, filled once per network pass,
        # instead of a memset in front of every such launch (forward outputs / backward input gradients)
        self.zf, self.zf_off, self.zb, self.zb_off = None, [], None, []
        self.dwpool = None         # d wmod of every tensor-core layer in one buffer, zeroed once per step (the wgrad kernels accumulate)
        self.dwoff, self.dwtotal, self.dwpool_home = [], 0, None
        self.wstream = None        # stream the weight-gradient GEMMs run on (CONFIG['wgrad_stream']); the bank's backward joins it

    def zeros(self, lidx, cout):
        """(d bias [cout], d strength []) views into the pool: one fill per network and pass instead of two per layer.
        The pool made by the forward serves ONE backward pass (the views are handed to autograd as gradients and may end up as
        param.grad); _Bank.backward drops it, so a second pass over a retained graph gets a fresh, zeroed pool here."""
        if self.zpool is None:
            torch.cuda.synchronize()                       # rare path (retain_graph): order the fill against both stream branches the blunt way
            self.zpool = torch.zeros([self.ztotal], device=self.styles[0].device, dtype=torch.float32)
            if self.side is not None:
                self.zpool.record_stream(self.side)
            torch.cuda.synchronize()
        o = self.zoff[lidx]
        return self.zpool[o:o + cout], self.zpool[o + cout]

    def take_f(self, lidx, shape):
        """Zero-filled forward output of layer lidx from the pool, or None (the launch does not split K / no pool)."""
        return self._take(self.zf, self.zf_off, lidx, shape)

    def take_b(self, lidx, shape):
        """Zero-filled input gradient of layer lidx from the backward pool, or None.  The pool serves ONE backward pass."""
        return self._take(self.zb, self.zb_off, lidx, shape)

    @staticmethod
    def _take(pool, offs, lidx, shape):
        if pool is None or lidx < 0 or offs[lidx] < 0:
            return None
        cnt = int(np.prod(shape))
        return pool[offs[lidx]:offs[lidx] + cnt].view(shape)

    def dwslice(self, lidx, n):
        """Zero-initialised d wmod [n, taps, cout, cin] of layer lidx inside the pool, or None when the layer is not pooled."""
        if self.dwoff[lidx] < 0:
            return None
        if self.dwpool is None:
            # first weight gradient of this backward pass: one fill for all layers, on the stream the wgrad kernels run on
            cur = torch.cuda.current_stream()
            self.dwpool = torch.empty([self.dwtotal], device=self.styles[0].device, dtype=torch.float32)
            if CONFIG['wgrad_stream']:
                self.wstream = wgrad_stream(self.dwpool.device)
                self.wstream.wait_stream(cur)           # the block may still be in use by earlier work of the allocating stream
                with torch.cuda.stream(self.wstream):
                    self.dwpool.zero_()
                self.dwpool.record_stream(self.wstream)
            else:
                self.dwpool.zero_()
            self.dwpool_home = cur
        sp = self.specs[lidx]
        o = self.dwoff[lidx]
        return self.dwpool[o:o + n * sp.taps * sp.cout * sp.cin].view(n, sp.taps, sp.cout, sp.cin)

    def run_wgrad(self, tensors, fn):
        """Run fn() (one wgrad launch reading `tensors`) on the weight-gradient stream, after the work queued on the current one."""
        ws = self.wstream
        if ws is None:
            cur = torch.cuda.current_stream()
            if self.dwpool_home is not None and cur != self.dwpool_home:
                cur.wait_stream(self.dwpool_home)       # the pool was zeroed on the stream of the first layer backward
            fn()
            return
        ws.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(ws):
            fn()
        for t in tensors:
            if t is not None:
                t.record_stream(ws)


def _bank_array(bank, n, dev, bwd=None):
    arr = (_lib.BankLayer * len(bank.specs))()
    for l, sp in enumerate(bank.specs):
        e = arr[l]
        e.affine_w, e.affine_b, e.weight = ptr(sp._aw), ptr(sp._ab), ptr(sp._W)
        e.styles, e.dcoef, e.wmod, e.w_hi, e.w_lo = ptr(bank.styles[l]), ptr(bank.dcoef[l]), ptr(bank.wmod[l]), ptr(bank.w_hi[l]), ptr(bank.w_lo[l])
        e.widx, e.cin, e.cout, e.taps, e.demod, e.post_scale = sp.widx, sp.cin, sp.cout, sp.taps, int(sp.demod), sp.post_scale
        if bwd is not None:
            e.dwmod, e.d_weight, e.d_styles, e.d_affine_w, e.d_affine_b = bwd[l]
    return arr


class _Bank(torch.autograd.Function):
    """token = bank(ws, params...): fills `bank` (styles, dcoef, modulated weights of all layers).  The backward runs after
    every layer's backward has stored its d wmod in bank.dwmod and turns them into parameter / latent gradients."""

    @staticmethod
    @device_guard
    def forward(ctx, ws, bank, *params):
        ws = _f32c(ws)
        n, num_ws, w_dim = ws.shape
        dev = ws.device
        keep_lo = CONFIG['fwd_passes'] == 3 or CONFIG['dgrad_passes'] == 3
        for l, sp in enumerate(bank.specs):
            sp._aw, sp._ab, sp._W = _f32c(params[3 * l]), _f32c(params[3 * l + 1]), _f32c(params[3 * l + 2])
            bank.styles[l] = torch.empty([n, sp.cin], device=dev, dtype=torch.float32)
            bank.dcoef[l] = torch.empty([n, sp.cout], device=dev, dtype=torch.float32) if sp.demod else None
            shape = [n, sp.taps, sp.cout, sp.cin]
            if not (sp.tc_f and sp.tc_b):
                bank.wmod[l] = torch.empty(shape, device=dev, dtype=torch.float32)
            if sp.tc_f or sp.tc_b:
                bank.w_hi[l] = torch.empty(shape, device=dev, dtype=torch.bfloat16)
                bank.w_lo[l] = torch.empty(shape, device=dev, dtype=torch.bfloat16) if keep_lo else None
        bank.zoff, tot = [], 0
        for sp in bank.specs:
            bank.zoff.append(tot)
            tot += (sp.cout + 1 + 3) // 4 * 4
        bank.ztotal = tot
        bank.zpool = torch.zeros([tot], device=dev, dtype=torch.float32)
        arr = _bank_array(bank, n, dev)
        call('b200_bank_styles_fwd', ctypes.addressof(arr), len(bank.specs), ptr(ws), n, num_ws, w_dim, stream())
        call('b200_bank_weights_fwd', ctypes.addressof(arr), len(bank.specs), n, stream())
        ctx.bank = bank
        ctx.save_for_backward(ws)
        bank.need_wgrad = any(ctx.needs_input_grad)
        # zero pools for the split-K launches (the library tells which shapes split K at this batch size)
        lib = _lib.load()
        bank.zf_off, bank.zb_off, tf, tb = [], [], 0, 0
        for sp in bank.specs:
            raw = sp.h * sp.w if sp.up == 1 else (2 * sp.h + 1) * (2 * sp.w + 1)
            kf = lib.b200_conv_tc_ksplit(0, n, sp.h, sp.w, sp.cin, sp.cout, sp.k, sp.up) if (sp.tc_f and CONFIG['zero_pools']) else 1
            kd = lib.b200_conv_tc_ksplit(1, n, sp.h, sp.w, sp.cin, sp.cout, sp.k, sp.up) if (sp.tc_b and CONFIG['zero_pools']) else 1
            bank.zf_off.append(tf if kf > 1 else -1)
            tf += n * raw * sp.cout if kf > 1 else 0
            bank.zb_off.append(tb if kd > 1 else -1)
            tb += n * sp.h * sp.w * sp.cin if kd > 1 else 0
        bank.zf = torch.zeros([tf], device=dev, dtype=torch.float32) if tf else None
        bank.zb = torch.zeros([tb], device=dev, dtype=torch.float32) if (tb and any(ctx.needs_input_grad)) else None
        bank.dwpool, bank.wstream = None, None
        bank.dwoff, tot = [], 0
        for sp in bank.specs:                       # layout of the d wmod pool (allocated and zeroed by the first wgrad of the backward)
            bank.dwoff.append(tot if sp.tc_b else -1)
            tot += n * sp.taps * sp.cout * sp.cin if sp.tc_b else 0
        bank.dwtotal = tot
        return torch.zeros([1], device=dev, dtype=torch.float32)

    @staticmethod
    @device_guard
    def backward(ctx, _dtok):
        bank = ctx.bank
        (ws,) = ctx.saved_tensors
        n, num_ws, w_dim = ws.shape
        dev = ws.device
        need = ctx.needs_input_grad
        specs = bank.specs
        if bank.side is not None:                   # d wmod of the side-stream layers carries no autograd edge: join explicitly
            torch.cuda.current_stream().wait_stream(bank.side)
        if bank.wstream is not None:
            torch.cuda.current_stream().wait_stream(bank.wstream)
        if bank.dwpool is not None:
            if bank.wstream is None and bank.dwpool_home != torch.cuda.current_stream():
                torch.cuda.current_stream().wait_stream(bank.dwpool_home)      # filled on the stream of the first layer backward
            bank.dwpool.record_stream(torch.cuda.current_stream())
        d_ws = torch.zeros_like(ws) if need[0] else None
        total_cin = sum(sp.cin for sp in specs)
        ds_all = torch.zeros([total_cin * n], device=dev, dtype=torch.float32)
        grads, bwd, off = [], [], 0
        for l, sp in enumerate(specs):
            gaw = torch.empty_like(sp._aw) if need[2 + 3 * l] else None
            gab = torch.empty_like(sp._ab) if need[3 + 3 * l] else None
            gW = torch.empty_like(sp._W) if need[4 + 3 * l] else None
            dw = bank.dwmod[l]
            if dw is None:                                   # layer not reached by the backward pass: zero gradients
                gaw = torch.zeros_like(sp._aw) if gaw is not None else None
                gab = torch.zeros_like(sp._ab) if gab is not None else None
                gW = torch.zeros_like(sp._W) if gW is not None else None
            ds = ds_all[off:off + n * sp.cin]
            off += n * sp.cin
            bwd.append((ptr(dw), ptr(gW), ptr(ds), ptr(gaw), ptr(gab)))
            grads += [gaw, gab, gW]
        arr = _bank_array(bank, n, dev, bwd)
        call('b200_bank_weights_bwd', ctypes.addressof(arr), len(specs), n, stream())
        call('b200_bank_styles_bwd', ctypes.addressof(arr), len(specs), ptr(ws), ptr(d_ws), n, num_ws, w_dim, stream())
        bank.dwmod = [None] * len(specs)
        bank.zb = None             # served this backward pass; a second pass over a retained graph lets the kernels clear their outputs
        bank.dwpool = None
        bank.zpool = None          # its slices now belong to autograd (possibly as param.grad): never accumulate into them again
        bank.token = None          # token -> grad_fn -> ctx.bank -> bank was a reference cycle keeping the per-step weight tensors alive
        return (d_ws, None, *grads)


_SIDE = {}
_WGRAD = {}


def _device_stream(table, device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in table:
        table[key] = torch.cuda.Stream(device=key)
    return table[key]


def side_stream(device):
    """The per-device second stream of CONFIG['overlap'] (created on first use)."""
    return _device_stream(_SIDE, device)


def wgrad_stream(device):
    """The per-device stream of CONFIG['wgrad_stream'] (created on first use)."""
    return _device_stream(_WGRAD, device)


def make_bank(ws, specs):
    """Run the bank for `specs` on latents ws [N, num_ws, w_dim]; returns the filled WeightBank."""
    bank = WeightBank(specs)
    params = []
    for sp in specs:
        params += [sp.affine.weight, sp.affine.bias, sp.weight]
    bank.token = _Bank.apply(ws, bank, *params)
    return bank


def _bf16_like(t):
    return torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)


def _fir_fused(x, f2, up, down, pad, flip, gain, want_f32=True, want_split=False, need_lo=True, add=None, act=None, separable=False):
    """b200_upfirdn2d_fused on NHWC x.  act = (bias, noise, strength, noise_bs, act_gain, clamp) or None.
    Returns (y fp32 | None, y_hi | None, y_lo | None)."""
    n, h, w, c = x.shape
    px0, px1, py0, py1 = pad
    fh, fw = f2.shape
    oh = (h * up + py0 + py1 - fh + down) // down
    ow = (w * up + px0 + px1 - fw + down) // down
    y = torch.empty([n, oh, ow, c], device=x.device, dtype=torch.float32) if want_f32 else None
    yh = torch.empty([n, oh, ow, c], device=x.device, dtype=torch.bfloat16) if want_split else None
    yl = torch.empty([n, oh, ow, c], device=x.device, dtype=torch.bfloat16) if (want_split and need_lo) else None
    if act is None:
        a = (0, None, None, None, 0, 0, 0.0, 1.0, -1.0)
    else:
        bias, noise, strength, nbs, act_gain, clamp = act
        a = (1, ptr(bias), ptr(noise), ptr(strength), nbs, 1, 0.2, float(act_gain), float(clamp))
    call('b200_upfirdn2d_fused', ptr(x), ptr(f2), ptr(add), ptr(y), ptr(yh), ptr(yl), n, h, w, c, fh, fw, up, down, px0, px1, py0, py1,
         int(flip), float(gain), *a, int(separable), stream())
    return y, yh, yl


class _ModConvLayer(torch.autograd.Function):
    """z = clamp(lrelu(modconv(x, W, styles) [-> FIR if up=2] + noise*strength + bias) * act_gain, +-clamp)

    x [N,H,W,Cin] NHWC (+ optional split-bf16 copies x_hi/x_lo made by the producing layer), weight [Cout,Cin,3,3],
    styles [N,Cin], noise [H',W'] or [N,1,H',W'] or None, strength 0-dim tensor.
    Returns (z, z_hi, z_lo); the bf16 pair is None on the exact-fp32 path.  networks_stylegan2.py:311-330 + :34-91.
    """

    @staticmethod
    @device_guard
    def forward(ctx, x, x_hi, x_lo, weight, styles, bias, noise, strength, up, act_gain, clamp, token=None, bank=None, lidx=-1,
                want_z=True, fork=False):
        ctx.set_materialize_grads(False)       # no zero-filled gradients for the non-differentiable bf16 outputs
        n, h, w, cin = x.shape
        have_split = x_hi is not None and (x_lo is not None or CONFIG['fwd_passes'] != 3)
        if not have_split:
            x = _f32c(x)                       # (with the producer's split pair at hand x may be a shape-only placeholder: never touch it)
        if bank is not None:
            sp = bank.specs[lidx]
            cout, k = sp.cout, sp.k
            W, s = sp._W, bank.styles[lidx]
        else:
            cout, _, k, _ = weight.shape
            W = _f32c(weight)
            s = _f32c(styles)
        taps = k * k
        dev = x.device
        b = _f32c(bias)
        clampf = float(clamp if clamp is not None else -1)
        oh, ow = h * up, w * up
        nz = st = None
        nbs = 0
        if noise is not None:
            nz = _f32c(noise)
            nbs = oh * ow if nz.ndim == 4 else 0
            st = _f32c(strength)
        if bank is not None:
            tc = sp.tc_f
            dcoef = bank.dcoef[lidx]
        else:
            tc = _tc_ok(0, h, w, cin, cout, k, up) and _tc_ok(1, h, w, cin, cout, k, up) and _tc_ok(2, h, w, cin, cout, k, up)
            dcoef = torch.empty([n, cout], device=dev, dtype=torch.float32)
        lean = tc and not want_z               # keep the output only as its split-bf16 pair
        if lean:
            z = _placeholder([n, oh, ow, cout], dev)
        else:
            z = torch.empty([n, oh, ow, cout], device=dev, dtype=torch.float32)
        zp = None if lean else ptr(z)
        if tc:
            fp = CONFIG['fwd_passes']
            keep_lo = fp == 3 or CONFIG['dgrad_passes'] == 3
            if bank is not None:
                w_hi, w_lo = bank.w_hi[lidx], bank.w_lo[lidx]
            else:
                w_hi = torch.empty([n, taps, cout, cin], device=dev, dtype=torch.bfloat16)
                w_lo = torch.empty_like(w_hi) if keep_lo else None
                call('b200_modconv_weight_prep', ptr(W), ptr(s), None, ptr(w_hi), ptr(w_lo), ptr(dcoef), n, cout, cin, taps, 1, stream())
            if x_hi is None or (fp == 3 and x_lo is None):
                x_hi, x_lo = _split(x, fp == 3)
            z_hi = torch.empty([n, oh, ow, cout], device=dev, dtype=torch.bfloat16)
            z_lo = torch.empty_like(z_hi)
            if up == 1 and _lib.load().b200_conv_tc_act_fusable(n, h, w, cin, cout, k) == 1:
                # epilogue applied while the accumulator leaves tensor memory: no fp32 round trip of the raw conv output
                call('b200_conv_fwd_tc_act', ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), zp, ptr(z_hi), ptr(z_lo), ptr(b), ptr(nz), ptr(st),
                     nbs, n, h, w, cin, cout, k, fp, 0.2, float(act_gain), clampf, stream())
            elif up == 1:
                y = bank.take_f(lidx, [n, oh, ow, cout]) if bank is not None else None
                pz = int(y is not None)
                if y is None:
                    y = torch.empty([n, oh, ow, cout], device=dev, dtype=torch.float32)
                call('b200_conv_fwd_tc', ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(y), n, h, w, cin, cout, k, 1, fp, pz, stream())
                call('b200_layer_act_fwd', ptr(y), zp, ptr(z_hi), ptr(z_lo), ptr(b), ptr(nz), ptr(st), nbs, n, oh * ow, cout, 1, 0.2,
                     float(act_gain), clampf, stream())
            else:
                zt = bank.take_f(lidx, [n, 2 * h + 1, 2 * w + 1, cout]) if bank is not None else None
                pz = int(zt is not None)
                if zt is None:
                    zt = torch.empty([n, 2 * h + 1, 2 * w + 1, cout], device=dev, dtype=torch.float32)
                call('b200_conv_fwd_tc', ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(zt), n, h, w, cin, cout, k, 2, fp, pz, stream())
                # 4x4 FIR (pad 1, gain 4) fused with the layer epilogue and the bf16 split for the next conv
                call('b200_upfirdn2d_fused', ptr(zt), ptr(fir_filter(dev)), None, zp, ptr(z_hi), ptr(z_lo), n, 2 * h + 1, 2 * w + 1,
                     cout, 4, 4, 1, 1, 1, 1, 1, 1, 0, 4.0, 1, ptr(b), ptr(nz), ptr(st), nbs, 1, 0.2, float(act_gain), clampf, 1, stream())
            # the activation backward needs the OUTPUT (sign and clamp): the fp32 copy, or -- lean -- the split pair the consumer keeps anyway
            ctx.save_for_backward(x_hi, x_lo if CONFIG['wgrad_passes'] == 3 else None, W, s, w_hi, w_lo, dcoef,
                                  None if lean else z, nz, st, z_hi if lean else None, z_lo if lean else None)
            ctx.mark_non_differentiable(z_hi, z_lo)
        else:
            if bank is not None:
                wmod = bank.wmod[lidx]
            else:
                wmod = torch.empty([n, taps, cout, cin], device=dev, dtype=torch.float32)
                call('b200_modconv_weight_prep', ptr(W), ptr(s), ptr(wmod), None, None, ptr(dcoef), n, cout, cin, taps, 1, stream())
            if up == 1:
                y = torch.empty_like(z)
                call('b200_conv_fwd', ptr(x), ptr(wmod), ptr(y), n, h, w, cin, cout, k, 1, stream())
            else:
                zt = torch.empty([n, 2 * h + 1, 2 * w + 1, cout], device=dev, dtype=torch.float32)
                call('b200_conv_fwd', ptr(x), ptr(wmod), ptr(zt), n, h, w, cin, cout, k, 2, stream())
                y = _upfirdn_nhwc_raw(zt, fir_filter(dev), (1, 1), (1, 1), (1, 1, 1, 1), False, 4.0)
            call('b200_layer_act_fwd', ptr(y), ptr(z), None, None, ptr(b), ptr(nz), ptr(st), nbs, n, oh * ow, cout, 1, 0.2,
                 float(act_gain), clampf, stream())
            z_hi = z_lo = None
            ctx.save_for_backward(x, None, W, s, wmod, None, dcoef, z, nz, st, None, None)
        ctx.cfg = (tc, up, float(act_gain), clampf, k, nbs, (n, h, w, cin, cout))
        ctx.bank, ctx.lidx = bank, lidx
        if fork:
            # a second handle of the (lean) output for its second consumer: the two gradients then arrive separately in backward
            # and are summed inside the activation-backward kernel instead of in an accumulation pass of autograd's
            assert lean, 'fork needs a lean output (consumers that read the split pair only)'
            return z, z_hi, z_lo, _placeholder([n, oh, ow, cout], dev)
        return z, z_hi, z_lo

    @staticmethod
    @device_guard
    def backward(ctx, dz, _dhi, _dlo, dz2=None):
        if dz is None:
            dz, dz2 = dz2, None
        if dz is None:
            return (None,) * 16
        xs, xs_lo, W, s, wm, wm_lo, dcoef, z, nz, st, zr_hi, zr_lo = ctx.saved_tensors
        tc, up, act_gain, clamp, k, nbs, (n, h, w, cin, cout) = ctx.cfg
        bank, lidx = ctx.bank, ctx.lidx
        taps = k * k
        dev = dz.device
        oh, ow = h * up, w * up
        zref = (ptr(z), ptr(zr_hi), ptr(zr_lo))          # saved output: fp32, or (lean) its split pair
        need = ctx.needs_input_grad
        need_x, need_w = need[0], ((need[3] or need[4]) if bank is None else bank.need_wgrad)
        dzc = _f32c(dz)
        has_noise = nz is not None
        if dz2 is not None:
            dz2 = _f32c(dz2)
            lo2 = (need_x and CONFIG['dgrad_passes'] == 3) or (need_w and CONFIG['wgrad_passes'] == 3)
            if not (tc and up == 1 and lo2 and _lib.load().b200_layer_act_bwd_sum2_supported(n, oh * ow, cout, int(has_noise), nbs) == 1):
                dzc, dz2 = dzc + dz2, None          # shapes / modes the two-addend kernel does not take: the pass autograd would have run
        if bank is not None:
            dbias, dstr = bank.zeros(lidx, cout)
            dstr = dstr if has_noise else None
        else:
            dbias = torch.zeros([cout], device=dev, dtype=torch.float32)
            dstr = torch.zeros([], device=dev, dtype=torch.float32) if has_noise else None
        dnoise = torch.zeros_like(nz) if (has_noise and need[6]) else None
        dx = bank.take_b(lidx, [n, h, w, cin]) if (need_x and tc and bank is not None) else None
        pz = int(dx is not None)
        if dx is None and need_x:
            dx = torch.empty([n, h, w, cin], device=dev, dtype=torch.float32)
        dW = ds = None
        if tc:
            dp, wp = CONFIG['dgrad_passes'], CONFIG['wgrad_passes']
            lo = (need_x and dp == 3) or (need_w and wp == 3)
            if up == 1:
                dy_hi, dy_lo = _bf16_like(dzc), (_bf16_like(dzc) if lo else None)
                if dz2 is not None:
                    call('b200_layer_act_bwd_sum2', ptr(dzc), ptr(dz2), *zref, None, ptr(dy_hi), ptr(dy_lo), ptr(dbias), ptr(nz), ptr(st), nbs,
                         ptr(dstr), ptr(dnoise), n, oh * ow, cout, 1, 0.2, act_gain, clamp, stream())
                else:
                    call('b200_layer_act_bwd', ptr(dzc), *zref, None, ptr(dy_hi), ptr(dy_lo), ptr(dbias), ptr(nz), ptr(st), nbs, ptr(dstr),
                         ptr(dnoise), n, oh * ow, cout, 1, 0.2, act_gain, clamp, stream())
            else:
                dy = torch.empty_like(dzc)
                call('b200_layer_act_bwd', ptr(dzc), *zref, ptr(dy), None, None, ptr(dbias), ptr(nz), ptr(st), nbs, ptr(dstr),
                     ptr(dnoise), n, oh * ow, cout, 1, 0.2, act_gain, clamp, stream())
                # adjoint of the FIR (pad 2, flipped filter) straight to split bf16 on the (2h+1)x(2w+1) grid
                _, dy_hi, dy_lo = _fir_fused(dy, fir_filter(dev), 1, 1, (2, 2, 2, 2), True, 4.0, want_f32=False, want_split=True, need_lo=lo,
                                             separable=True)          # fir_filter() is the outer product [1,3,3,1] x [1,3,3,1] / 64
            if need_x:
                call('b200_conv_dgrad_tc', ptr(dy_hi), ptr(dy_lo), ptr(wm), ptr(wm_lo), ptr(dx), n, h, w, cin, cout, k, up, dp, pz, stream())
            if need_w:
                dwmod = bank.dwslice(lidx, n) if bank is not None else None
                if dwmod is not None:
                    bank.run_wgrad((xs, xs_lo, dy_hi, dy_lo), lambda: call(
                        'b200_conv_wgrad_tc', ptr(xs), ptr(xs_lo), ptr(dy_hi), ptr(dy_lo), ptr(dwmod), n, h, w, cin, cout, k, up, wp, 1, stream()))
                else:
                    dwmod = torch.empty([n, taps, cout, cin], device=dev, dtype=torch.float32)
                    call('b200_conv_wgrad_tc', ptr(xs), ptr(xs_lo), ptr(dy_hi), ptr(dy_lo), ptr(dwmod), n, h, w, cin, cout, k, up, wp, 0, stream())
        else:
            dy = torch.empty_like(dzc)
            call('b200_layer_act_bwd', ptr(dzc), *zref, ptr(dy), None, None, ptr(dbias), ptr(nz), ptr(st), nbs, ptr(dstr), ptr(dnoise),
                 n, oh * ow, cout, 1, 0.2, act_gain, clamp, stream())
            if up == 2:
                dy = _upfirdn_nhwc_raw(dy, fir_filter(dev), (1, 1), (1, 1), (2, 2, 2, 2), True, 4.0)
            if need_x:
                call('b200_conv_dgrad', ptr(dy), ptr(wm), ptr(dx), n, h, w, cin, cout, k, up, stream())
            if need_w:
                dwmod = torch.empty([n, taps, cout, cin], device=dev, dtype=torch.float32)
                call('b200_conv_wgrad', ptr(xs), ptr(dy), ptr(dwmod), n, h, w, cin, cout, k, up, stream())
        dtok = None
        if need_w and bank is not None:
            bank.dwmod[lidx] = dwmod                    # consumed by _Bank.backward (runs after every layer's backward)
            if lidx == 0:
                dtok = torch.zeros([1], device=dev, dtype=torch.float32)
        elif need_w:
            dW = torch.empty_like(W)
            ds = torch.empty_like(s)
            call('b200_modconv_weight_prep_bwd', ptr(W), ptr(s), ptr(dcoef), ptr(dwmod), ptr(dW), ptr(ds), n, cout, cin, taps, 1, stream())
        return dx, None, None, dW, ds, dbias, dnoise, dstr, None, None, None, dtok, None, None, None, None


def _placeholder(shape, device):
    """Shape-only fp32 stand-in (zero strides, one element of storage) for an activation that exists only as its split-bf16 pair:
    autograd routes the gradient of that shape through it, nothing ever reads its values."""
    key = str(device)
    if key not in _PLACEHOLDER:
        _PLACEHOLDER[key] = torch.zeros([1], device=device, dtype=torch.float32)
    return _PLACEHOLDER[key].expand(shape)


_PLACEHOLDER = {}


def lean_ok(bank, consumers):
    """True when every consumer layer (bank indices) of an activation reads it through the split-bf16 pair in both directions, so the
    producer may skip the fp32 copy (CONFIG['lean_acts']).  A consumer qualifies when its forward runs on the tensor cores and its
    backward does too, or is the thin 1x1 path that takes the pair (the 3-channel ToRGB layers)."""
    if bank is None or not CONFIG['lean_acts'] or not CONFIG['tc']:
        return False
    lib = _lib.load()
    for l in consumers:
        if l >= len(bank.specs):
            continue
        sp = bank.specs[l]
        thin = sp.k == 1 and sp.up == 1 and lib.b200_conv1x1_thin_supported(sp.cin, sp.cout) == 1
        if not (sp.tc_f and (sp.tc_b or thin)):
            return False
    return True


def modconv_layer(x, weight, styles, bias, noise, strength, up, act_gain, clamp, x_split=None, bank=None, lidx=-1, want_z=True,
                  fork=False):
    """Returns (z, (z_hi, z_lo) or None) -- with fork=True (z, pair, z_second): a second handle of a lean output for its second
    consumer, so that the two gradients meet inside the activation backward (b200_layer_act_bwd_sum2) instead of in an add pass.
    x_split: the producer's split-bf16 copies of x, if it made them.
    bank / lidx: take styles and modulated weights from a WeightBank entry instead of (weight, styles).
    want_z=False (tensor-core layers only): do not write the fp32 output; z is then a shape-only placeholder and the pair is the
    activation (see lean_ok)."""
    xh, xl = x_split if x_split is not None else (None, None)
    with prof_range('modulated_conv2d'):
        if bank is not None:
            if fork and not want_z and bank.specs[lidx].tc_f:      # (a layer off the tensor cores writes its fp32 output: one handle, autograd adds)
                z, zh, zl, z2 = _ModConvLayer.apply(x, xh, xl, None, None, bias, noise, strength, up, act_gain, clamp, bank.token, bank,
                                                    lidx, False, True)
                return z, (zh, zl), z2
            z, zh, zl = _ModConvLayer.apply(x, xh, xl, None, None, bias, noise, strength, up, act_gain, clamp, bank.token, bank, lidx,
                                            bool(want_z))
        else:
            z, zh, zl = _ModConvLayer.apply(x, xh, xl, weight, styles, bias, noise, strength, up, act_gain, clamp)
    if fork:
        return z, ((zh, zl) if zh is not None else None), z
    return z, ((zh, zl) if zh is not None else None)


def _off_chain(bank, tensors, fn):
    """Run fn() (a weight-gradient launch whose only consumer is the bank's backward) on the weight-gradient stream behind the work
    queued on the current one; in place when there is no bank / no such stream."""
    if bank is None or not CONFIG['wgrad_stream']:
        fn()
        return
    ws = wgrad_stream(tensors[0].device)
    bank.wstream = ws                           # _Bank.backward joins it (the same per-device stream WeightBank.dwslice picks)
    ws.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ws):
        fn()
    for t in tensors:
        if t is not None:
            t.record_stream(ws)


class _ToRGB(torch.autograd.Function):
    """img = [upsample2d(img_prev)] + clamp(conv1x1(x, W*styles) + bias, +-clamp)     (networks_stylegan2.py:353-357, 451-457)

    x [N,H,W,Cin] (+ optional split copies), weight [Cimg,Cin,1,1], styles [N,Cin] (already scaled by 1/sqrt(Cin)),
    img_prev [N,H/2,W/2,Cimg] or None.
    """

    @staticmethod
    @device_guard
    def forward(ctx, x, x_hi, x_lo, weight, styles, bias, img_prev, clamp, token=None, bank=None, lidx=-1):
        ctx.set_materialize_grads(False)
        n, h, w, cin = x.shape
        dev = x.device
        fp = CONFIG['fwd_passes']
        have_split = x_hi is not None and x_lo is not None
        if not have_split:
            x = _f32c(x)                       # (with the producer's split pair at hand x may be a shape-only placeholder: never touch it)
        if bank is not None:
            sp = bank.specs[lidx]
            cimg, W, s = sp.cout, sp._W, bank.styles[lidx]
            tc_f, tc_b = sp.tc_f, sp.tc_b
            wmod, w_hi, w_lo = bank.wmod[lidx], bank.w_hi[lidx], bank.w_lo[lidx]
        else:
            cimg = weight.shape[0]
            W = _f32c(weight)
            s = _f32c(styles)
            tc_f = _tc_ok(0, h, w, cin, cimg, 1, 1)
            tc_b = _tc_ok(1, h, w, cin, cimg, 1, 1) and _tc_ok(2, h, w, cin, cimg, 1, 1)
            wmod = w_hi = w_lo = None
            if not (tc_f and tc_b):
                wmod = torch.empty([n, 1, cimg, cin], device=dev, dtype=torch.float32)
            if tc_f or tc_b:
                w_hi = torch.empty([n, 1, cimg, cin], device=dev, dtype=torch.bfloat16)
                w_lo = torch.empty_like(w_hi)
            call('b200_modconv_weight_prep', ptr(W), ptr(s), ptr(wmod), ptr(w_hi), ptr(w_lo), None, n, cimg, cin, 1, 0, stream())
        y = bank.take_f(lidx, [n, h, w, cimg]) if (bank is not None and tc_f) else None
        pz = int(y is not None)
        if y is None:
            y = torch.empty([n, h, w, cimg], device=dev, dtype=torch.float32)
        if (tc_f or tc_b) and (x_hi is None or x_lo is None):
            x_hi, x_lo = _split(x, True)
        b = _f32c(bias)
        cl = float(clamp if clamp is not None else -1)
        if have_split and wmod is not None and _lib.load().b200_conv1x1_fwd_thin_supported(cin, cimg) == 1:
            # 3-channel images from 64 / 128 input channels: a streaming dot product over the split pair with bias and clamp on the
            # way out, instead of a 128-wide tensor-core tile (125 columns of padding) plus a bias / clamp pass
            call('b200_conv1x1_fwd_thin', ptr(x_hi), ptr(x_lo), ptr(wmod), ptr(b), ptr(y), n, h * w, cin, cimg, cl, stream())
        else:
            if tc_f:
                call('b200_conv_fwd_tc', ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(y), n, h, w, cin, cimg, 1, 1, fp, pz, stream())
            else:
                call('b200_conv_fwd', ptr(_f32c(x)), ptr(wmod), ptr(y), n, h, w, cin, cimg, 1, 1, stream())
            call('b200_bias_act', ptr(y), ptr(b), None, None, None, ptr(y), 0, y.numel(), 1, cimg, 1, 0.0, 1.0, cl, stream())
        if img_prev is not None:
            img = _upfirdn_nhwc_raw(_f32c(img_prev), fir_filter(dev), (2, 2), (1, 1), (2, 1, 2, 1), False, 4.0, add=y)
        else:
            img = y
        # fp32 backward (channel counts the tensor cores cannot take): the thin 1x1 weight gradient reads x from the pair when it exists
        split_x = (not tc_b) and have_split and _lib.load().b200_conv1x1_thin_supported(cin, cimg) == 1
        ctx.cfg = (cl, img_prev is not None, tc_b, (n, h, w, cin, cimg), split_x)
        ctx.bank, ctx.lidx = bank, lidx
        if tc_b:
            ctx.save_for_backward(x_hi, x_lo if CONFIG['wgrad_passes'] == 3 else None, W, s, w_hi, w_lo, y)
        elif split_x:
            ctx.save_for_backward(x_hi, x_lo, W, s, wmod, None, y)
        else:
            ctx.save_for_backward(_f32c(x), None, W, s, wmod, None, y)
        return img

    @staticmethod
    @device_guard
    def backward(ctx, dimg):
        if dimg is None:
            return (None,) * 11
        xs, xs_lo, W, s, wm, wm_lo, y = ctx.saved_tensors
        cl, has_prev, tc_b, (n, h, w, cin, cimg), split_x = ctx.cfg
        bank, lidx = ctx.bank, ctx.lidx
        dev = y.device
        need = ctx.needs_input_grad
        need_x, need_w = need[0], ((need[3] or need[4]) if bank is None else bank.need_wgrad)
        dimg = _f32c(dimg)
        dbias = bank.zeros(lidx, cimg)[0] if bank is not None else torch.zeros([cimg], device=dev, dtype=torch.float32)
        dx = bank.take_b(lidx, [n, h, w, cin]) if (need_x and tc_b and bank is not None) else None
        pz = int(dx is not None)
        if dx is None and need_x:
            dx = torch.empty([n, h, w, cin], device=dev, dtype=torch.float32)
        dW = ds = None
        if need_w:
            dwmod = bank.dwslice(lidx, n) if (bank is not None and tc_b) else None
            pooled = dwmod is not None
            if not pooled:
                dwmod = torch.empty([n, 1, cimg, cin], device=dev, dtype=torch.float32)
        # gradient of bias + clamp from the saved clamped output (bias_act.cu:143-145), d bias reduced in the same pass
        if tc_b:
            dp, wp = CONFIG['dgrad_passes'], CONFIG['wgrad_passes']
            lo = (need_x and dp == 3) or (need_w and wp == 3)
            dy_hi, dy_lo = _bf16_like(dimg), (_bf16_like(dimg) if lo else None)
            call('b200_layer_act_bwd', ptr(dimg), ptr(y), None, None, None, ptr(dy_hi), ptr(dy_lo), ptr(dbias), None, None, 0, None, None,
                 n, h * w, cimg, 0, 0.0, 1.0, cl, stream())
            if need_x:
                call('b200_conv_dgrad_tc', ptr(dy_hi), ptr(dy_lo), ptr(wm), ptr(wm_lo), ptr(dx), n, h, w, cin, cimg, 1, 1, dp, pz, stream())
            if need_w and pooled:
                bank.run_wgrad((xs, xs_lo, dy_hi, dy_lo), lambda: call(
                    'b200_conv_wgrad_tc', ptr(xs), ptr(xs_lo), ptr(dy_hi), ptr(dy_lo), ptr(dwmod), n, h, w, cin, cimg, 1, 1, wp, 1, stream()))
            elif need_w:
                call('b200_conv_wgrad_tc', ptr(xs), ptr(xs_lo), ptr(dy_hi), ptr(dy_lo), ptr(dwmod), n, h, w, cin, cimg, 1, 1, wp, 0, stream())
        else:
            dy = torch.empty_like(dimg)
            call('b200_layer_act_bwd', ptr(dimg), ptr(y), None, None, ptr(dy), None, None, ptr(dbias), None, None, 0, None, None,
                 n, h * w, cimg, 0, 0.0, 1.0, cl, stream())
            if need_x:
                call('b200_conv_dgrad', ptr(dy), ptr(wm), ptr(dx), n, h, w, cin, cimg, 1, 1, stream())
            # the thin weight gradients feed only the bank's backward: on the weight-gradient stream, so that this node (and with it
            # the main chain waiting for dx) ends after the dgrad
            if need_w and split_x:
                _off_chain(bank, (xs, xs_lo, dy, dwmod), lambda: call(
                    'b200_conv1x1_wgrad_split', ptr(xs), ptr(xs_lo), ptr(dy), ptr(dwmod), n, h * w, cin, cimg, stream()))
            elif need_w:
                _off_chain(bank, (xs, dy, dwmod), lambda: call(
                    'b200_conv_wgrad', ptr(xs), ptr(dy), ptr(dwmod), n, h, w, cin, cimg, 1, 1, stream()))
        if need_w and bank is not None:
            bank.dwmod[lidx] = dwmod
        elif need_w:
            dW = torch.empty_like(W)
            ds = torch.empty_like(s)
            call('b200_modconv_weight_prep_bwd', ptr(W), ptr(s), None, ptr(dwmod), ptr(dW), ptr(ds), n, cimg, cin, 1, 0, stream())
        dprev = None
        if has_prev and need[6]:
            dprev = _upfirdn_nhwc_raw(dimg, fir_filter(dev), (1, 1), (2, 2), (1, 1, 1, 1), True, 4.0)
        return dx, None, None, dW, ds, dbias, dprev, None, None, None, None


def torgb_layer(x, weight, styles, bias, img_prev, clamp, x_split=None, bank=None, lidx=-1):
    xh, xl = x_split if x_split is not None else (None, None)
    with prof_range('modulated_conv2d'):
        if bank is not None:
            return _ToRGB.apply(x, xh, xl, None, None, bias, img_prev, clamp, bank.token, bank, lidx)
        return _ToRGB.apply(x, xh, xl, weight, styles, bias, img_prev, clamp)


# ----------------------------------------------------------------------------------------------
# Renderer: fused tri-plane sampling + decoder, per-ray hierarchical sampling and compositing

def _decoder_params(decoder):
    fc0, fc1 = decoder.net[0], decoder.net[2]
    return fc0.weight, fc0.bias, fc1.weight, fc1.bias, float(fc0.lr_multiplier)


class _RunModel(torch.autograd.Function):
    """rgb, sigma = decoder(sample_from_planes(planes, coords))   (renderer.py:197-203 without density noise)

    planes_nhwc [N,H,W,96]; coords [N,P,3].
    """

    @staticmethod
    @device_guard
    def forward(ctx, planes, coords, W1, b1, W2, b2, lr_mul, box_warp):
        pl = _f32c(planes)
        co = _f32c(coords)
        n, hp, wp, _ = pl.shape
        P = co.shape[1]
        rgb = torch.empty([n, P, 32], device=pl.device, dtype=torch.float32)
        sigma = torch.empty([n, P, 1], device=pl.device, dtype=torch.float32)
        w = [_f32c(t) for t in (W1, b1, W2, b2)]
        fs = _fsave(n, P, pl.device) if any(ctx.needs_input_grad) else None
        call('b200_triplane_mlp_fwd', ptr(pl), n, hp, wp, ptr(co), None, None, None, 0, 0, P, float(box_warp), *map(ptr, w),
             float(lr_mul), ptr(rgb), ptr(sigma), ptr(fs), stream())
        ctx.cfg = (float(lr_mul), float(box_warp))
        ctx.save_for_backward(pl, co, *w, fs)
        return rgb, sigma

    @staticmethod
    @device_guard
    def backward(ctx, d_rgb, d_sigma):
        pl, co, W1, b1, W2, b2, fs = ctx.saved_tensors
        lr_mul, box_warp = ctx.cfg
        n, hp, wp, _ = pl.shape
        P = co.shape[1]
        need = ctx.needs_input_grad
        d_planes = torch.zeros_like(pl) if need[0] else None
        d_coords = torch.empty_like(co) if need[1] else None
        wg = any(need[2:6])
        dws = [torch.zeros_like(t) for t in (W1, b1, W2, b2)] if wg else [None] * 4
        ws_bytes = _lib.load().b200_triplane_bwd_workspace_bytes(n, P)
        work = torch.empty([ws_bytes], device=pl.device, dtype=torch.uint8)
        call('b200_triplane_mlp_bwd', ptr(pl), n, hp, wp, ptr(co), None, None, None, 0, 0, P, box_warp, ptr(W1), ptr(b1), ptr(W2),
             ptr(b2), lr_mul, ptr(_f32c(d_rgb)), ptr(_f32c(d_sigma)), ptr(fs), ptr(d_planes), ptr(d_coords), None, None, *map(ptr, dws),
             ptr(work), ws_bytes, stream())
        return (d_planes, d_coords, *dws, None, None)


def _fsave(n, P, device):
    """Forward -> backward hand-off buffer of the fused sampler + decoder (b200_triplane_fsave_bytes)."""
    return torch.empty([_lib.load().b200_triplane_fsave_bytes(n, P)], device=device, dtype=torch.uint8)


def run_model_nhwc(planes_nhwc, decoder, coords, box_warp):
    W1, b1, W2, b2, lr_mul = _decoder_params(decoder)
    return _RunModel.apply(planes_nhwc, coords, W1, b1, W2, b2, lr_mul, box_warp)


_MM_INIT = {}


def _minmax_init(device):
    key = str(device)
    if key not in _MM_INIT:
        _MM_INIT[key] = torch.tensor([-1, 0], device=device, dtype=torch.int32)
    return _MM_INIT[key]


class _Render(torch.autograd.Function):
    """ImportanceRenderer.forward as one autograd node (renderer.py:143-195, numeric ray_start / ray_end).

    planes [N,H,W,96]; ray_o, ray_d [N,M,3]; t_base [S] (linspace(ray_start, ray_end, S));
    u_strat [N,M,S,1], u_imp [N*M,S_imp] are the two uniform draws of the reference (renderer.py:245,292).
    Returns feat [N,M,32], depth [N,M,1], wsum [N,M,1].
    """

    @staticmethod
    @device_guard
    def forward(ctx, planes, ray_o, ray_d, W1, b1, W2, b2, lr_mul, box_warp, t_base, delta, u_strat, u_imp, white_back,
                density_noise, t_coarse=None):
        ctx.set_materialize_grads(False)
        pl = _f32c(planes)
        ro, rd = _f32c(ray_o), _f32c(ray_d)
        dev = pl.device
        n, hp, wp, _ = pl.shape
        M = ro.shape[1]
        S = t_base.numel() if t_coarse is None else t_coarse.shape[2]
        S2 = 0 if u_imp is None else u_imp.shape[1]
        w = [_f32c(t) for t in (W1, b1, W2, b2)]
        st = stream()
        # global depth range for the clamp of ray_marcher.py:50, as two order-preserving uint words {0xFFFFFFFF, 0} = (+inf, -inf)
        minmax = _minmax_init(dev).clone()
        if t_coarse is not None:          # per-ray limits / disparity sampling: depths prepared by the caller (renderer.py:230-242)
            t_c = _f32c(t_coarse).reshape(n, M, S)
            call('b200_depth_minmax', ptr(t_c), t_c.numel(), ptr(minmax), st)
        else:
            t_c = torch.empty([n, M, S], device=dev, dtype=torch.float32)
            call('b200_ray_depths_coarse', ptr(_f32c(t_base)), ptr(_f32c(u_strat)), ptr(t_c), n * M, S, float(delta), ptr(minmax), st)
        rgb_c = torch.empty([n, M, S, 32], device=dev, dtype=torch.float32)
        sig_c = torch.empty([n, M, S], device=dev, dtype=torch.float32)
        rw = 0
        if CONFIG['ray_patch_order'] and int(round(math.sqrt(M))) ** 2 == M:
            rw = int(round(math.sqrt(M)))
        want_grad = any(ctx.needs_input_grad)
        fs_c = _fsave(n, M * S, dev) if want_grad else None          # features of every point, kept for the backward
        call('b200_triplane_mlp_fwd', ptr(pl), n, hp, wp, None, ptr(ro), ptr(rd), ptr(t_c), S, rw, M * S, float(box_warp),
             *map(ptr, w), float(lr_mul), ptr(rgb_c), ptr(sig_c), ptr(fs_c), st)
        if density_noise > 0:
            sig_c += (torch.randn_like(sig_c.view(n, M * S, 1)) * density_noise).view(n, M, S)      # renderer.py:201-202 (same draw shape)
        t_f = rgb_f = sig_f = fs_f = None
        if S2 > 0:
            t_f = torch.empty([n, M, S2], device=dev, dtype=torch.float32)
            call('b200_ray_importance', ptr(t_c), ptr(sig_c), ptr(_f32c(u_imp)), ptr(t_f), n * M, S, S2, st)
            rgb_f = torch.empty([n, M, S2, 32], device=dev, dtype=torch.float32)
            sig_f = torch.empty([n, M, S2], device=dev, dtype=torch.float32)
            fs_f = _fsave(n, M * S2, dev) if want_grad else None
            call('b200_triplane_mlp_fwd', ptr(pl), n, hp, wp, None, ptr(ro), ptr(rd), ptr(t_f), S2, rw, M * S2, float(box_warp),
                 *map(ptr, w), float(lr_mul), ptr(rgb_f), ptr(sig_f), ptr(fs_f), st)
            if density_noise > 0:
                sig_f += (torch.randn_like(sig_f.view(n, M * S2, 1)) * density_noise).view(n, M, S2)
            # no second pass over t_f: every fine depth is an interpolation between two mid-points of its ray's coarse depths
            # (renderer.py:297-307), so the range of the merged samples IS the range of the coarse ones
        feat = torch.empty([n, M, 32], device=dev, dtype=torch.float32)
        depth = torch.empty([n, M, 1], device=dev, dtype=torch.float32)
        wsum = torch.empty([n, M, 1], device=dev, dtype=torch.float32)
        call('b200_ray_composite_fwd', ptr(t_c), ptr(sig_c), ptr(rgb_c), S, ptr(t_f), ptr(sig_f), ptr(rgb_f), S2, ptr(minmax),
             int(bool(white_back)), n * M, ptr(feat), ptr(depth), ptr(wsum), st)
        ctx.cfg = (float(lr_mul), float(box_warp), int(bool(white_back)), S, S2, rw)
        ctx.save_for_backward(pl, ro, rd, *w, t_c, sig_c, rgb_c, t_f, sig_f, rgb_f, minmax, fs_c, fs_f)
        return feat, depth, wsum

    @staticmethod
    @device_guard
    def backward(ctx, d_feat, d_depth, d_wsum):
        pl, ro, rd, W1, b1, W2, b2, t_c, sig_c, rgb_c, t_f, sig_f, rgb_f, minmax, fs_c, fs_f = ctx.saved_tensors
        lr_mul, box_warp, white_back, S, S2, rw = ctx.cfg
        n, hp, wp, _ = pl.shape
        M = ro.shape[1]
        dev = pl.device
        st = stream()
        need = ctx.needs_input_grad
        if d_feat is None:
            d_feat = torch.zeros([n, M, 32], device=dev, dtype=torch.float32)
        d_rgb_c = torch.empty_like(rgb_c)
        d_sig_c = torch.empty_like(sig_c)
        d_rgb_f = torch.empty_like(rgb_f) if S2 > 0 else None
        d_sig_f = torch.empty_like(sig_f) if S2 > 0 else None
        call('b200_ray_composite_bwd', ptr(t_c), ptr(sig_c), ptr(rgb_c), S, ptr(t_f), ptr(sig_f), ptr(rgb_f), S2, ptr(minmax),
             white_back, n * M, ptr(_f32c(d_feat)), ptr(_f32c(d_depth) if d_depth is not None else None),
             ptr(_f32c(d_wsum) if d_wsum is not None else None), ptr(d_rgb_c), ptr(d_sig_c),
             ptr(d_rgb_f), ptr(d_sig_f), st)
        d_planes = torch.zeros_like(pl) if need[0] else None
        want_rays = need[1] or need[2]
        wg = any(need[3:7])
        dws = [torch.zeros_like(t) for t in (W1, b1, W2, b2)] if wg else [None] * 4
        d_ro = d_rd = None
        if want_rays:
            d_ro = torch.zeros_like(ro)
            d_rd = torch.zeros_like(rd)
        ws_bytes = _lib.load().b200_triplane_bwd_workspace_bytes(n, M * max(S, S2))
        work = torch.empty([ws_bytes], device=dev, dtype=torch.uint8)
        for t, d_rgb, d_sig, s, fs in ((t_c, d_rgb_c, d_sig_c, S, fs_c), (t_f, d_rgb_f, d_sig_f, S2, fs_f)):
            if s == 0:
                continue
            # the per-ray sums d ray_o = sum_k d point, d ray_d = sum_k t_k * d point (point = o + t*d, renderer.py:161,178) are
            # reduced inside the C-ABI call
            call('b200_triplane_mlp_bwd', ptr(pl), n, hp, wp, None, ptr(ro), ptr(rd), ptr(t), s, rw, M * s, box_warp, ptr(W1),
                 ptr(b1), ptr(W2), ptr(b2), lr_mul, ptr(d_rgb), ptr(d_sig), ptr(fs), ptr(d_planes), None, ptr(d_ro), ptr(d_rd),
                 *map(ptr, dws), ptr(work), ws_bytes, st)
        return (d_planes, d_ro, d_rd, *dws, None, None, None, None, None, None, None, None, None)


def render(planes_nhwc, decoder, ray_o, ray_d, box_warp, t_base, delta, u_strat, u_imp, white_back=False, density_noise=0.0,
           t_coarse=None):
    W1, b1, W2, b2, lr_mul = _decoder_params(decoder)
    return _Render.apply(planes_nhwc, ray_o, ray_d, W1, b1, W2, b2, lr_mul, box_warp, t_base, delta, u_strat, u_imp,
                         white_back, density_noise, t_coarse)


class _RaySampler(torch.autograd.Function):
    """RaySampler.forward (ray_sampler.py:24-73) as one kernel; gradient to cam2world (the w-projection optimises the pose)."""

    @staticmethod
    @device_guard
    def forward(ctx, cam2world, intrinsics, resolution):
        c2w = _f32c(cam2world).reshape(-1, 16)
        K = _f32c(intrinsics).reshape(-1, 9)
        n, M = c2w.shape[0], resolution * resolution
        ray_o = torch.empty([n, M, 3], device=c2w.device, dtype=torch.float32)
        ray_d = torch.empty_like(ray_o)
        call('b200_ray_sampler_fwd', ptr(c2w), ptr(K), n, resolution, ptr(ray_o), ptr(ray_d), stream())
        ctx.res, ctx.shape = resolution, cam2world.shape
        ctx.save_for_backward(c2w, K)
        return ray_o, ray_d

    @staticmethod
    @device_guard
    def backward(ctx, d_o, d_d):
        c2w, K = ctx.saved_tensors
        g = torch.zeros_like(c2w)
        call('b200_ray_sampler_bwd', ptr(c2w), ptr(K), c2w.shape[0], ctx.res, ptr(_f32c(d_o) if d_o is not None else None),
             ptr(_f32c(d_d) if d_d is not None else None), ptr(g), stream())
        return g.reshape(ctx.shape), None, None


def ray_sampler(cam2world, intrinsics, resolution):
    return _RaySampler.apply(cam2world, intrinsics, resolution)


def library_info():
    lib = _lib.load()
    return {'path': _lib.LIB_PATH, 'version': lib.b200_version()}

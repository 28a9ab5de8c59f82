"""Look-alikes of torch_utils/ops/{bias_act,upfirdn2d,conv2d_resample,conv2d_gradfix,fma}.py on the b200eg3d kernels.

Same names, argument meaning and error behaviour as the reference functions (file:line cited per function); NCHW in / out like
the reference, NHWC inside.  Everything here is CUDA-only: a CPU tensor raises (there is no reference fallback to hide behind).
"""
import contextlib

import torch

from .. import ops
from .._lib import call, device_guard, ptr, stream

activation_funcs = {k: {'def_alpha': v[1], 'def_gain': v[2], 'cuda_idx': v[0], 'ref': ops._ACT_REF[k]} for k, v in ops._ACT.items()}

# ---------------------------------------------------------------------------------------------------------------------
# bias_act.py:54, upfirdn2d.py:72,120,279,315,354

bias_act = ops.bias_act
setup_filter = ops.setup_filter
upfirdn2d = ops.upfirdn2d
upsample2d = ops.upsample2d


def _filter_size(f):
    if f is None:
        return 1, 1
    return (int(f.shape[-1]), int(f.shape[0]))


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """upfirdn2d.py:279-311: same-size FIR filtering."""
    px0, px1, py0, py1 = ops._pad4(padding)
    fw, fh = _filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return ops.upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """upfirdn2d.py:354-386."""
    dx, dy = ops._pair(down)
    px0, px1, py0, py1 = ops._pad4(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return ops.upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


# ---------------------------------------------------------------------------------------------------------------------
# fma.py:17

class _FMA(torch.autograd.Function):
    """a * b + c with broadcasting (fma.py:17-60)."""

    @staticmethod
    def forward(ctx, a, b, c):
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return torch.addcmul(c, a, b)

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        need = ctx.needs_input_grad
        da = _unbroadcast(dout * b, a.shape) if need[0] else None
        db = _unbroadcast(dout * a, b.shape) if need[1] else None
        dc = _unbroadcast(dout, ctx.c_shape) if need[2] else None
        return da, db, dc


def _unbroadcast(x, shape):
    extra = x.ndim - len(shape)
    dims = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra or shape[i - extra] == 1)]
    if dims:
        x = x.sum(dim=dims, keepdim=True)
    if extra:
        x = x.reshape(-1, *x.shape[extra + 1:])
    return x.reshape(shape)


def fma(a, b, c):
    return _FMA.apply(a, b, c)


# ---------------------------------------------------------------------------------------------------------------------
# conv2d_gradfix.py:37,42 and conv2d_resample.py:48 on the b200eg3d convolution kernels

_NO_WGRAD = [False]


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    """conv2d_gradfix.py:27-34."""
    old = _NO_WGRAD[0]
    if disable:
        _NO_WGRAD[0] = True
    yield
    _NO_WGRAD[0] = old


class _GroupedConv(torch.autograd.Function):
    """y[g] = conv(x[g], w[g]) for G independent (sample, weight) pairs -- what the reference expresses as one grouped convolution
    (networks_stylegan2.py:83-86 with groups = batch) or, for G = 1 weight shared by the batch, as a plain convolution.

    x [G, H, W, I] NHWC, w [G, taps, O, I] in the library's GEMM layout (tap = ky * k + kx, correlation order for up = 1,
    convolution order for the stride-2 transposed case).  up = 2 returns the (2H+1) x (2W+1) transposed-convolution output."""

    @staticmethod
    @device_guard
    def forward(ctx, x, w, k, up):
        x, w = ops._f32c(x), ops._f32c(w)
        g, h, wd, cin = x.shape
        cout = w.shape[2]
        oh, ow = (h, wd) if up == 1 else (2 * h + 1, 2 * wd + 1)
        y = torch.empty([g, oh, ow, cout], device=x.device, dtype=torch.float32)
        ops._conv_fwd(x, w, y, g, h, wd, cin, cout, k, up)
        ctx.cfg = (g, h, wd, cin, cout, k, up)
        ctx.save_for_backward(x, w)
        ctx.no_wgrad = _NO_WGRAD[0]
        return y

    @staticmethod
    @device_guard
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        g, h, wd, cin, cout, k, up = ctx.cfg
        dy = ops._f32c(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops._conv_dgrad(dy, w, dx, g, h, wd, cin, cout, k, up)
        if ctx.needs_input_grad[1] and not ctx.no_wgrad:
            dw = torch.empty_like(w)
            ops._conv_wgrad(x, dy, dw, g, h, wd, cin, cout, k, up)
        return dx, dw, None, None


def _to_gemm_weight(w, groups, flip):
    """[G*O, I, k, k] (reference layout) -> [G, k*k, O, I]; flip = rotate the kernel by 180 degrees."""
    go, cin, kh, kw = w.shape
    if flip:
        w = w.flip([2, 3])
    return w.reshape(groups, go // groups, cin, kh * kw).permute(0, 3, 1, 2).contiguous()


def _conv_core(x, w, k, up, groups, flip):
    """x [N, G*I, H, W] NCHW with the reference's grouping convention -> [N, G*O, H', W'] NCHW through _GroupedConv."""
    n, gc, h, wd = x.shape
    cin = w.shape[1]
    if groups * cin != gc:
        raise ValueError('conv: input channels do not match weight / groups')
    if groups > 1 and n != 1:
        raise NotImplementedError('b200eg3d conv shim: groups > 1 is supported in the reference\'s "one group per sample" form (batch folded into channels, N = 1)')
    wg = _to_gemm_weight(w.to(torch.float32), groups, flip)
    if groups == 1:
        xg = x.permute(0, 2, 3, 1)                                           # [N, H, W, I]; one weight shared by the batch
        wg = wg.expand(n, -1, -1, -1)
        y = _GroupedConv.apply(xg, wg, k, up)                                # [N, H', W', O]
        return y.permute(0, 3, 1, 2)
    xg = x.reshape(groups, cin, h, wd).permute(0, 2, 3, 1)                   # [G, H, W, I]
    y = _GroupedConv.apply(xg, wg, k, up)                                    # [G, H', W', O]
    return y.permute(0, 3, 1, 2).reshape(1, -1, y.shape[1], y.shape[2])


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """conv2d_gradfix.py:37-40 for the shapes of the EG3D synthesis stack: k in {1, 3}, stride 1, 'same' padding, dilation 1."""
    k = int(weight.shape[-1])
    pad = padding if isinstance(padding, int) else (padding[0] if padding[0] == padding[-1] else -1)
    if stride not in (1, (1, 1), [1, 1]) or dilation not in (1, (1, 1), [1, 1]) or weight.shape[-2] != k or k not in (1, 3) or pad != k // 2:
        raise NotImplementedError('b200eg3d conv2d shim: only k in {1,3}, stride 1, padding k//2, dilation 1 (the EG3D synthesis stack)')
    y = _conv_core(input, weight, k, 1, groups, flip=False)                   # F.conv2d is a correlation
    return y if bias is None else y + bias.reshape(1, -1, 1, 1)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """conv2d_gradfix.py:42-45 for the one shape the up-sampling layers use: 3x3, stride 2, padding 0 -> (2H+1) x (2W+1).
    weight is [G*I, O, 3, 3] as F.conv_transpose2d expects."""
    k = int(weight.shape[-1])
    if stride not in (2, (2, 2), [2, 2]) or padding not in (0, (0, 0), [0, 0]) or output_padding not in (0, (0, 0), [0, 0]) or k != 3 or \
            dilation not in (1, (1, 1), [1, 1]):
        raise NotImplementedError('b200eg3d conv_transpose2d shim: only 3x3, stride 2, padding 0 (conv2d_resample.py:113-127)')
    gi, o = weight.shape[0], weight.shape[1]
    w = weight.reshape(groups, gi // groups, o, k, k).transpose(1, 2).reshape(groups * o, gi // groups, k, k)   # back to [G*O, I, k, k]
    y = _conv_core(input, w, k, 2, groups, flip=False)                        # conv_transpose2d scatters w unflipped = true convolution
    return y if bias is None else y + bias.reshape(1, -1, 1, 1)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:48-143 for the branches the EG3D stack takes: (up, down) = (1, 1) with 'same' padding and
    (2, 1) with the [1,3,3,1] filter; other combinations raise NotImplementedError (the reference falls back to generic
    upfirdn2d + conv there, none of which inversion ever reaches)."""
    if not (isinstance(x, torch.Tensor) and x.ndim == 4 and isinstance(w, torch.Tensor) and w.ndim == 4):
        raise AssertionError('conv2d_resample: x and w must be 4-D tensors')
    kh, kw = int(w.shape[2]), int(w.shape[3])
    fw, fh = _filter_size(f)
    px0, px1, py0, py1 = ops._pad4(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down != 1 or kh != kw or kh not in (1, 3) or up not in (1, 2):
        raise NotImplementedError('b200eg3d conv2d_resample shim: down = 1, square 1x1 / 3x3 kernels, up in {1, 2}')
    if up == 1:
        if not (px0 == px1 == py0 == py1 == kh // 2):
            raise NotImplementedError('b200eg3d conv2d_resample shim: up = 1 needs padding = kernel_size // 2')
        return _conv_core(x, w, kh, 1, groups, flip=not flip_weight)          # flip_weight=True is the correlation F.conv2d computes
    if kh != 3:                                                               # conv2d_resample.py:99-102: 1x1 -> convolve, then upsample
        y = _conv_core(x, w, 1, 1, groups, flip=False)
        return ops.upfirdn2d(y, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    # conv2d_resample.py:113-131: stride-2 transposed convolution (flip_weight = not flip_weight) then the FIR with the leftover padding
    px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
    pxt, pyt = max(min(-px0, -px1), 0), max(min(-py0, -py1), 0)
    if pxt != 0 or pyt != 0:
        raise NotImplementedError('b200eg3d conv2d_resample shim: transposed-convolution padding other than 0')
    y = _conv_core(x, w, 3, 2, groups, flip=flip_weight)                      # kernel in convolution order unless the caller asked for correlation
    return ops.upfirdn2d(y, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)

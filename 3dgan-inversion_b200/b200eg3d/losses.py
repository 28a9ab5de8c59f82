"""Caller-side losses of the inversion loops as fused CUDA reductions.

    pti_loss(generated_images, real_images)      training/coaches/base_coach.py:101-126 (calc_loss) without the LPIPS term
    compute_tv_norm(values)                      training/coaches/base_coach.py:294-305

LPIPS needs pretrained AlexNet/VGG weights and stays with the caller; everything else of calc_loss -- the area
down-sampling of the target, both MSE terms and the depth total-variation term -- is one forward and one backward kernel.
"""
import ctypes

import torch

from ._lib import call, device_guard, ptr, stream


def _strides(t):
    arr = (ctypes.c_long * 4)(*[int(s) for s in t.stride()])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def _dense_like_nchw(t):
    """Gradient buffer for an NCHW-shaped tensor in channels-last memory (what the NHWC generator consumes without a copy)."""
    n, c, h, w = t.shape
    return torch.empty([n, h, w, c], device=t.device, dtype=torch.float32).permute(0, 3, 1, 2)


class _PTILoss(torch.autograd.Function):
    @staticmethod
    @device_guard
    def forward(ctx, image, image_raw, image_depth, real, l2_lambda, tv_lambda):
        real = real.detach().to(torch.float32).contiguous()
        n, c, h, w = real.shape
        img = image.detach().to(torch.float32) if image is not None else None
        raw = image_raw.detach().to(torch.float32) if image_raw is not None else None
        dep = image_depth.detach().to(torch.float32).contiguous() if image_depth is not None else None
        r = raw.shape[-1] if raw is not None else (dep.shape[-1] if dep is not None else h)
        if img is not None and tuple(img.shape) != (n, c, h, w):
            raise ValueError('image and real_images must have the same shape')
        if raw is not None and tuple(raw.shape) != (n, c, r, r):
            raise ValueError('image_raw must be [N, C, R, R]')
        if dep is not None and dep.numel() != n * r * r:
            raise ValueError('image_depth must be [N, 1, R, R]')
        out = torch.zeros([4], device=real.device, dtype=torch.float32)
        keep = [_strides(t) if t is not None else (None, None) for t in (img, raw)]
        call('b200_pti_loss_fwd', ptr_any(img), keep[0][1], ptr_any(raw), keep[1][1], ptr(dep), ptr(real), n, c, h, w, r,
             float(l2_lambda), float(tv_lambda), ptr(out), stream())
        ctx.cfg = (n, c, h, w, r, float(l2_lambda), float(tv_lambda))
        ctx.save_for_backward(img, raw, dep, real)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    @device_guard
    def backward(ctx, dloss, _dparts):
        img, raw, dep, real = ctx.saved_tensors
        n, c, h, w, r, l2, tv = ctx.cfg
        need = ctx.needs_input_grad
        img = img if need[0] else None
        raw = raw if need[1] else None
        dep = dep if need[2] else None
        g = dloss.detach().to(torch.float32).reshape(1).contiguous()
        d_img = _dense_like_nchw(img) if img is not None else None
        d_raw = _dense_like_nchw(raw) if raw is not None else None
        d_dep = torch.empty_like(dep) if dep is not None else None
        keep = [_strides(t) if t is not None else (None, None) for t in (img, raw, d_img, d_raw)]
        call('b200_pti_loss_bwd', ptr_any(img), keep[0][1], ptr_any(raw), keep[1][1], ptr(dep), ptr(real), n, c, h, w, r, l2, tv,
             ptr(g), ptr_any(d_img), keep[2][1], ptr_any(d_raw), keep[3][1], ptr(d_dep), stream())
        return d_img, d_raw, d_dep, None, None, None


def ptr_any(t):
    """Device pointer of a possibly strided CUDA tensor (the kernel receives its element strides)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('b200eg3d: tensor is not on a CUDA device (there is no CPU path)')
    return t.data_ptr()


def pti_loss(generated_images, real_images, pt_l2_lambda=1.0, depth_tv_lambda=1.0, return_parts=False):
    """loss = pt_l2_lambda * (mse(image, real) + mse(image_raw, area_R(real))) + depth_tv_lambda * tv(image_depth).

    generated_images: the dict G.synthesis returns ('image', 'image_raw', 'image_depth'); real_images [N, C, H, W].
    With return_parts, also returns the tensor [total, mse_image, mse_raw, tv] (detached) for logging."""
    loss, parts = _PTILoss.apply(generated_images.get('image'), generated_images.get('image_raw'),
                                 generated_images.get('image_depth'), real_images, pt_l2_lambda, depth_tv_lambda)
    return (loss, parts) if return_parts else loss


def compute_tv_norm(values):
    """base_coach.py:294-305 on a depth map [..., R, R] (mean of squared forward differences)."""
    v = values.reshape(-1, 1, values.shape[-2], values.shape[-1])
    dummy = torch.zeros([v.shape[0], 1, v.shape[-1], v.shape[-1]], device=v.device)
    loss, _ = _PTILoss.apply(None, None, v, dummy, 0.0, 1.0)
    return loss

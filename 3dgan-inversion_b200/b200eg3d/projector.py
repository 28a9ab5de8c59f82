"""Stage-1 (w-projection) caller-side operators on the b200eg3d CUDA library.

Mirrors, with the reference's names and argument order, the per-step pieces of training/projectors/w_projector.py:145-270
that are not already inside G.synthesis:

    calc_warping_loss(...)          training/warping_loss.py:6-56   (geometry fused into one kernel, features stay with the caller)
    LinePlaneCollision(...)         training/warping_loss.py:58-72  (kept for callers that use it directly)
    noise_regularizer(bufs)         w_projector.py:221-237          (one launch per pyramid level for all buffers)
    normalize_noise_(bufs)          w_projector.py:262-268          (two launches for all buffers)

Feature networks (VGG16 / LPIPS) need pretrained weights and remain ordinary torch modules supplied by the caller.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import call, ptr, stream

_BIG = 3.0e38


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class _WarpUV(torch.autograd.Function):
    """pred_uv [R*R, 2] of warping_loss.py:18-43 from (extrinsic, init_ext, intrinsic, depth); gradients to extrinsic and depth."""

    @staticmethod
    def forward(ctx, extrinsic, init_ext, intrinsic, depth):
        ext = _f32c(extrinsic).reshape(16)
        ini = _f32c(init_ext).reshape(16)
        K = _f32c(intrinsic).reshape(9)
        dep = _f32c(depth)
        R = dep.shape[-1]
        if dep.numel() != R * R:
            raise ValueError('warp_uv handles one depth map [1, 1, R, R] (the reference path is batch 1)')
        w2c = torch.linalg.inv(ini.reshape(4, 4)).contiguous()               # warping_loss.py:38
        uv = torch.empty([R * R, 2], device=dep.device, dtype=torch.float32)
        mn = torch.full([1], _BIG, device=dep.device, dtype=torch.float32)
        call('b200_warp_uv_fwd', ptr(ext), ptr(ini), ptr(w2c), ptr(K), ptr(dep), R, ptr(uv), ptr(mn), stream())
        ctx.R = R
        ctx.shapes = (extrinsic.shape, depth.shape)
        ctx.save_for_backward(ext, ini, w2c, K, dep)
        ctx.mark_non_differentiable(mn)
        return uv, mn

    @staticmethod
    def backward(ctx, d_uv, _dmn):
        ext, ini, w2c, K, dep = ctx.saved_tensors
        d_ext = torch.zeros([16], device=dep.device, dtype=torch.float32)
        d_dep = torch.empty_like(dep)
        call('b200_warp_uv_bwd', ptr(ext), ptr(ini), ptr(w2c), ptr(K), ptr(dep), ctx.R, ptr(_f32c(d_uv)), ptr(d_ext), ptr(d_dep), stream())
        es, ds = ctx.shapes
        return d_ext.reshape(es), None, None, d_dep.reshape(ds)


def warp_uv(extrinsic, init_ext, intrinsic, depth, check_intersection=True, epsilon=1e-6):
    """Canonical-view uv in [-1, 1] of every pixel's surface point.  check_intersection reproduces the reference's
    RuntimeError (warping_loss.py:66-67); it synchronises, so switch it off inside CUDA-graph capture."""
    uv, mn = _WarpUV.apply(extrinsic, init_ext, intrinsic, depth)
    if check_intersection and mn.item() < epsilon:
        raise RuntimeError('no intersection or line is within plane')
    return uv


def LinePlaneCollision(planeNormal, planePoint, rayDirection, rayPoint, epsilon=1e-6):
    """warping_loss.py:58-72 (inputs [N, 3]); plain torch, used by callers outside the fused path."""
    ndotu = (planeNormal * rayDirection).sum(-1, keepdim=True)
    if abs(torch.min(ndotu)) < epsilon:
        raise RuntimeError('no intersection or line is within plane')
    w_vec = rayPoint - planePoint
    si = -(planeNormal * w_vec).sum(-1, keepdim=True) / ndotu
    return w_vec + si * rayDirection + planePoint


def get_features(x, model, layers):
    """warping_loss.py:74-111: activations after child number 7 / 14 / 21 of a torchvision-VGG-style `features` module."""
    stop = {'7': 7, '14': 14, '21': 21}.get(str(layers))
    if stop is None:
        print('layers must be multipliers of 7')
        raise ValueError
    for idx, layer in enumerate(model.children()):
        x = layer(x)
        if idx == stop:
            return x
    raise ValueError('feature network has fewer children than the requested layer')


def photometric_reconstruction_loss(tgt_img, ref_img, depth_mask, explainability_mask=None):
    """explainability_network/loss_functions.py:9-20."""
    return ((tgt_img - ref_img) * depth_mask).abs().mean()


def calc_warping_loss(ws, canonical_cam, extrinsic, init_ext, intrinsic, depth, target_images, G, torch_vgg, ray_generator=None,
                      layers='14', check_intersection=True):
    """training/warping_loss.py:6-56 with the same arguments.  Returns (loss, warped canonical image)."""
    canonical_dict = G.synthesis(ws, canonical_cam, noise_mode='const', force_fp32=True)
    can_images = canonical_dict['image']
    if can_images.shape[2] > 256:
        can_images = F.interpolate(can_images, size=(256, 256), mode='area')
    depth_mean = torch.mean(depth)
    masked_depths = torch.where(depth < depth_mean, torch.ones_like(depth_mean), torch.zeros_like(depth_mean))   # foreground only
    pred_uv = warp_uv(extrinsic, init_ext, intrinsic, depth, check_intersection=check_intersection)
    torch_target_features = get_features(target_images, torch_vgg, layers)
    torch_synth_features = get_features(can_images, torch_vgg, layers)
    res = depth.shape[-1]
    fr = torch_target_features.shape[-1]
    pred_uv_resized = F.interpolate(pred_uv.reshape(1, res, res, -1).permute(0, 3, 1, 2), size=(fr, fr), mode='bilinear').permute(0, 2, 3, 1)
    warpped_feature = F.grid_sample(torch_synth_features, pred_uv_resized, mode='bilinear', align_corners=False)
    warpped_image = F.grid_sample(can_images, pred_uv.reshape(1, res, res, -1), mode='bilinear', align_corners=False)
    masked_depths = F.interpolate(masked_depths, size=(fr, fr), mode='bilinear')
    loss = photometric_reconstruction_loss(warpped_feature, torch_target_features, masked_depths)
    return loss, warpped_image


# ----------------------------------------------------------------------------------------------------------------------
# noise regulariser / normalisation over all noise buffers of a generator

def _ptr_table(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def _size_table(tensors):
    for t in tensors:
        if t.ndim != 2 or t.shape[0] != t.shape[1]:
            raise ValueError('noise buffers must be square 2-D tensors')
    arr = (ctypes.c_int * len(tensors))(*[int(t.shape[0]) for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


_SIZES_DEV = {}


def _sizes_dev(bufs):
    """Device copy of the side lengths, cached so that no host->device copy happens inside a CUDA-graph capture."""
    key = (str(bufs[0].device), tuple(int(b.shape[0]) for b in bufs))
    if key not in _SIZES_DEV:
        _SIZES_DEV[key] = torch.tensor(key[1], device=bufs[0].device, dtype=torch.int32)
    return _SIZES_DEV[key]


class _NoiseReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *bufs):
        bs = [_f32c(b) for b in bufs]
        dev = bs[0].device
        sizes, sizes_p = _size_table(bs)
        tab, tab_p = _ptr_table(bs)
        nwork = _lib.load().b200_noise_pyramid_work_floats(len(bs), sizes_p)
        work = torch.empty([max(nwork, 1)], device=dev, dtype=torch.float32)
        sums = torch.zeros([len(bs) * 8 * 2], device=dev, dtype=torch.float32)
        sizes_dev = _sizes_dev(bs)
        reg = torch.zeros([], device=dev, dtype=torch.float32)
        for b in bs:
            ptr(b)                                                  # device / contiguity checks
        call('b200_noise_pyramid_fwd', len(bs), tab_p, sizes_p, ptr(sizes_dev), ptr(work), ptr(sums), ptr(reg), stream())
        ctx.save_for_backward(work, sums, *bs)
        return reg

    @staticmethod
    def backward(ctx, dreg):
        work, sums, *bs = ctx.saved_tensors
        sizes, sizes_p = _size_table(bs)
        tab, tab_p = _ptr_table(bs)
        grads = [torch.empty_like(b) for b in bs]
        gtab, gtab_p = _ptr_table(grads)
        gwork = torch.empty_like(work)
        g = dreg.detach().to(torch.float32).reshape(1).contiguous()
        call('b200_noise_pyramid_bwd', len(bs), tab_p, sizes_p, ptr(work), ptr(sums), ptr(g), gtab_p, ptr(gwork), stream())
        return tuple(grads)


def noise_regularizer(noise_bufs):
    """reg_loss of w_projector.py:221-237 summed over all given buffers ([res, res] each; an iterable or a name->buffer dict)."""
    bufs = list(noise_bufs.values()) if isinstance(noise_bufs, dict) else list(noise_bufs)
    if not bufs:
        return torch.zeros([])
    return _NoiseReg.apply(*bufs)


def normalize_noise_(noise_bufs):
    """w_projector.py:262-268, in place on every buffer: zero mean, unit second moment."""
    bufs = list(noise_bufs.values()) if isinstance(noise_bufs, dict) else list(noise_bufs)
    if not bufs:
        return
    with torch.no_grad():
        for b in bufs:
            if b.dtype != torch.float32 or not b.is_contiguous():
                raise RuntimeError('normalize_noise_: buffers must be contiguous fp32 (they are updated in place)')
            ptr(b)
        sizes, sizes_p = _size_table(bufs)
        tab, tab_p = _ptr_table(bufs)
        stats = torch.zeros([2 * len(bufs)], device=bufs[0].device, dtype=torch.float32)
        call('b200_noise_normalize', len(bufs), tab_p, sizes_p, ptr(stats), stream())

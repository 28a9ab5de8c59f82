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
from ._lib import call, device_guard, ptr, stream

_BIG = 3.0e38


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class _WarpUV(torch.autograd.Function):
    """pred_uv [R*R, 2] of warping_loss.py:18-43 from (extrinsic, init_ext, intrinsic, depth); gradients to extrinsic and depth."""

    @staticmethod
    @device_guard
    def forward(ctx, extrinsic, init_ext, intrinsic, depth, w2c):
        ext = _f32c(extrinsic).reshape(16)
        ini = _f32c(init_ext).reshape(16)
        K = _f32c(intrinsic).reshape(9)
        dep = _f32c(depth)
        R = dep.shape[-1]
        if dep.numel() != R * R:
            raise ValueError('warp_uv handles one depth map [1, 1, R, R] (the reference path is batch 1)')
        # warping_loss.py:38; torch.linalg.inv synchronises, so callers that capture CUDA graphs pass the (constant) inverse in
        w2c = torch.linalg.inv(ini.reshape(4, 4)).contiguous() if w2c is None else _f32c(w2c).reshape(4, 4)
        uv = torch.empty([R * R, 2], device=dep.device, dtype=torch.float32)
        mn = torch.full([1], _BIG, device=dep.device, dtype=torch.float32)
        call('b200_warp_uv_fwd', ptr(ext), ptr(ini), ptr(w2c), ptr(K), ptr(dep), R, ptr(uv), ptr(mn), stream())
        ctx.R = R
        ctx.shapes = (extrinsic.shape, depth.shape)
        ctx.save_for_backward(ext, ini, w2c, K, dep)
        ctx.mark_non_differentiable(mn)
        return uv, mn

    @staticmethod
    @device_guard
    def backward(ctx, d_uv, _dmn):
        ext, ini, w2c, K, dep = ctx.saved_tensors
        d_ext = torch.zeros([16], device=dep.device, dtype=torch.float32)
        d_dep = torch.empty_like(dep)
        call('b200_warp_uv_bwd', ptr(ext), ptr(ini), ptr(w2c), ptr(K), ptr(dep), ctx.R, ptr(_f32c(d_uv)), ptr(d_ext), ptr(d_dep), stream())
        es, ds = ctx.shapes
        return d_ext.reshape(es), None, None, d_dep.reshape(ds), None


def warp_uv(extrinsic, init_ext, intrinsic, depth, check_intersection=True, epsilon=1e-6, w2c=None):
    """Canonical-view uv in [-1, 1] of every pixel's surface point.  check_intersection reproduces the reference's
    RuntimeError (warping_loss.py:66-67); it synchronises, so switch it off inside CUDA-graph capture."""
    uv, mn = _WarpUV.apply(extrinsic, init_ext, intrinsic, depth, w2c)
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
                      layers='14', check_intersection=True, w2c=None):
    """training/warping_loss.py:6-56 with the same arguments.  Returns (loss, warped canonical image)."""
    canonical_dict = G.synthesis(ws, canonical_cam, noise_mode='const', force_fp32=True)
    can_images = canonical_dict['image']
    if can_images.shape[2] > 256:
        can_images = F.interpolate(can_images, size=(256, 256), mode='area')
    depth_mean = torch.mean(depth)
    masked_depths = torch.where(depth < depth_mean, torch.ones_like(depth_mean), torch.zeros_like(depth_mean))   # foreground only
    pred_uv = warp_uv(extrinsic, init_ext, intrinsic, depth, check_intersection=check_intersection, w2c=w2c)
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
    @device_guard
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
    @device_guard
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
    with torch.no_grad(), _lib.on_device(bufs[0]):
        for b in bufs:
            if b.dtype != torch.float32 or not b.is_contiguous():
                raise RuntimeError('normalize_noise_: buffers must be contiguous fp32 (they are updated in place)')
            ptr(b)
        sizes, sizes_p = _size_table(bufs)
        tab, tab_p = _ptr_table(bufs)
        stats = torch.zeros([2 * len(bufs)], device=bufs[0].device, dtype=torch.float32)
        call('b200_noise_normalize', len(bufs), tab_p, sizes_p, ptr(stats), stream())


# ----------------------------------------------------------------------------------------------------------------------
# one iteration of the w-projection loop (w_projector.py:160-260) around a pose given as a rotation matrix

def rot6d_to_rotmat(x):
    """utils/camera_utils.py:259-273 (6-D rotation representation -> [B, 3, 3]).  The cross product is taken along the last
    dimension; the reference calls torch.cross without `dim` (legacy: first dimension of size 3), identical for every batch
    size but 3."""
    x = x.view(-1, 2, 3) + 1e-4
    a1, a2 = x[:, 0, :], x[:, 1, :]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum('bi,bi->b', b1, a2).unsqueeze(-1) * b1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-1)


def assemble_extrinsic(pred_rotmat, translation_opt, radius=2.7):
    """w_projector.py:160-171: cam2world from the predicted rotation and the optimisable translation, radius renormalised."""
    dev = pred_rotmat.device
    pred_translation = -radius * pred_rotmat[:, :3, 2]
    translation_opt_world = -torch.bmm(pred_rotmat, translation_opt.unsqueeze(-1)) * 2.7
    tmp_translation = translation_opt_world.squeeze(-1) + pred_translation
    tmp_translation = tmp_translation / torch.norm(tmp_translation, dim=-1) * 2.7
    bottom = F.pad(torch.ones([pred_rotmat.shape[0], 1, 1], device=dev), (3, 0))          # [0, 0, 0, 1] without a host->device copy
    return torch.cat([torch.cat([pred_rotmat, tmp_translation.unsqueeze(-1)], dim=2), bottom], dim=1)


def projection_step_loss(G, w_opt, pred_rotmat, translation_opt, init_ext, intrinsic, target_images, target_features, feature_fn,
                         torch_vgg, noise_bufs, noise_bufs2, w_noise=None, regularize_noise_weight=1e5, layers='14',
                         check_intersection=False, w2c=None):
    """Loss of one w-projection iteration (w_projector.py:160-240): synthesis at the predicted camera, warping loss through a
    second synthesis at the canonical camera, feature distance and noise regulariser.

    target_images: [1, 3, H, W] in [-1, 1] (the reference's target_images_contiguous); target_features = feature_fn of the
    0..255 target at <= 256 px; feature_fn / torch_vgg are the caller's (pretrained) feature networks."""
    pred_ext = assemble_extrinsic(pred_rotmat, translation_opt)
    pred_cam = torch.cat([pred_ext.reshape(-1, 16), intrinsic.reshape(1, 9)], dim=-1)
    canonical_cam = torch.cat([init_ext.reshape(-1, 16), intrinsic.reshape(1, 9)], dim=-1)
    ws_expand = (w_opt if w_noise is None else w_opt + w_noise).repeat(1, G.backbone.num_ws, 1)
    pred_dict = G.synthesis(ws_expand, pred_cam, noise_mode='const', force_fp32=True)
    pred_images = pred_dict['image'] * 127.5 + 128
    warp_loss, _ = calc_warping_loss(ws_expand.clone().detach(), canonical_cam.clone().detach(), pred_ext, init_ext, intrinsic,
                                     pred_dict['image_depth'], target_images, G, torch_vgg, None, layers=layers,
                                     check_intersection=check_intersection, w2c=w2c)
    if pred_images.shape[2] > 256:
        pred_images = F.interpolate(pred_images, size=(256, 256), mode='area')
    dist = (target_features - feature_fn(pred_images)).square().sum()
    reg_loss = noise_regularizer(list(noise_bufs.values()) + list(noise_bufs2.values()))
    return dist + reg_loss * regularize_noise_weight + warp_loss, {'dist': dist.detach(), 'warp': warp_loss.detach(), 'reg': reg_loss.detach()}

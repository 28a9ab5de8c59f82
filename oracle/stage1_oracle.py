"""CPU oracle for the stage-1 (w-projection) caller-side pieces (TEST INFRASTRUCTURE ONLY -- see eg3d_oracle.py).

Restates in plain torch what the reference computes around G.synthesis in training/projectors/w_projector.py:145-270:

    warp_uv              training/warping_loss.py:18-43 (+ ray_sampler.py:24-93, LinePlaneCollision :58-72)
    warping_loss         training/warping_loss.py:6-56 given the canonical image (the G.synthesis call is the caller's)
    noise_regularizer    w_projector.py:221-237
    normalize_noise      w_projector.py:262-268

Parity status: warp_uv / warping_loss are pinned against the reference's own calc_warping_loss, imported read-only in the
build container by oracle/make_goldens_stage1.py (fixture tests/golden/stage1_warp.npz: loss, warped image and gradients
w.r.t. the predicted extrinsic and the depth map).  The noise regulariser / normalisation are inline statements of
project() (not importable): restated line by line and checked against a literal transcription in the golden script.
"""
import torch
import torch.nn.functional as F

import eg3d_oracle as oracle


def warp_uv(extrinsic, init_ext, intrinsic, depth):
    """extrinsic, init_ext [1,4,4]; intrinsic [9] or [3,3]; depth [1,1,R,R] -> pred_uv [R*R, 2] in [-1, 1]."""
    R = depth.shape[-1]
    K = intrinsic.reshape(3, 3)
    o, d = oracle.ray_sampler(extrinsic.reshape(-1, 4, 4), K.unsqueeze(0), R)          # ray_sampler.py:24-73
    xyz = o[0] + d[0] * depth.reshape(R * R, 1)                                        # ray_sampler.py:75-93
    origin = init_ext.reshape(4, 4)[:3, 3].unsqueeze(0)                                # warping_loss.py:23-25
    vectors = xyz - origin
    normal = -origin
    plane_point = (init_ext.reshape(4, 4) @ torch.tensor([0., 0., 1., 1.], dtype=init_ext.dtype))[:3].unsqueeze(0)   # :28
    ndotu = (normal * vectors).sum(-1, keepdim=True)                                   # :64
    w_vec = origin - plane_point
    si = -(normal * w_vec).sum(-1, keepdim=True) / ndotu
    psi = w_vec + si * vectors + plane_point                                           # :69-71
    psi1 = torch.cat([psi, torch.ones_like(psi[:, :1])], dim=-1).t()
    w2c = torch.linalg.inv(init_ext.reshape(4, 4))                                     # :38
    uv = (w2c @ psi1)[:3].t()
    uv = uv / uv[:, 2:]
    uv = (K @ uv.t())[:2].t()                                                          # :41
    return (uv - 0.5) * 2, ndotu


def get_features(x, model, layers):
    stop = {'7': 7, '14': 14, '21': 21}[str(layers)]
    for idx, layer in enumerate(model.children()):
        x = layer(x)
        if idx == stop:
            return x


def warping_loss(can_images, extrinsic, init_ext, intrinsic, depth, target_images, vgg, layers='14'):
    """warping_loss.py:8-56 after the canonical synthesis call.  Returns (loss, warped image)."""
    if can_images.shape[2] > 256:
        can_images = F.interpolate(can_images, size=(256, 256), mode='area')
    mask = torch.where(depth < depth.mean(), torch.ones_like(depth.mean()), torch.zeros_like(depth.mean()))
    uv, _ = warp_uv(extrinsic, init_ext, intrinsic, depth)
    ft, fs = get_features(target_images, vgg, layers), get_features(can_images, vgg, layers)
    R, fr = depth.shape[-1], ft.shape[-1]
    uv_r = F.interpolate(uv.reshape(1, R, R, 2).permute(0, 3, 1, 2), size=(fr, fr), mode='bilinear').permute(0, 2, 3, 1)
    wf = F.grid_sample(fs, uv_r, mode='bilinear', align_corners=False)
    wi = F.grid_sample(can_images, uv.reshape(1, R, R, 2), mode='bilinear', align_corners=False)
    mask = F.interpolate(mask, size=(fr, fr), mode='bilinear')
    return ((wf - ft) * mask).abs().mean(), wi


def create_samples(N=256, voxel_origin=(0, 0, 0), cube_length=2.0):
    """training/coaches/single_id_coach.py:165-188 restated on CPU (fp32 index arithmetic kept as written there).
    Pinned by make_goldens_stage1.py, which executes the reference function's own source (extracted with ast)."""
    import numpy as np
    origin = np.array(voxel_origin) - cube_length / 2
    size = cube_length / (N - 1)
    idx = torch.arange(0, N ** 3, 1, dtype=torch.long)
    s = torch.zeros(N ** 3, 3)
    s[:, 2] = idx % N
    s[:, 1] = (idx.float() / N) % N
    s[:, 0] = ((idx.float() / N) / N) % N
    s[:, 0] = (s[:, 0] * size) + origin[2]
    s[:, 1] = (s[:, 1] * size) + origin[1]
    s[:, 2] = (s[:, 2] * size) + origin[0]
    return s.unsqueeze(0), origin, size


def noise_regularizer(bufs):
    """w_projector.py:221-237."""
    reg = 0.0
    for v in bufs:
        noise = v[None, None, :, :]
        while True:
            reg = reg + (noise * torch.roll(noise, shifts=1, dims=3)).mean() ** 2
            reg = reg + (noise * torch.roll(noise, shifts=1, dims=2)).mean() ** 2
            if noise.shape[2] <= 8:
                break
            noise = F.avg_pool2d(noise, kernel_size=2)
    return reg


def normalize_noise(bufs):
    """w_projector.py:262-268 (returns new tensors instead of updating in place)."""
    out = []
    for b in bufs:
        b = b - b.mean()
        out.append(b * b.square().mean().rsqrt())
    return out

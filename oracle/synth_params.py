"""Deterministic synthetic generator parameters and inputs (TEST INFRASTRUCTURE ONLY).

The reference's pickles are not available offline, so every parity test runs on a
random generator.  To make the *same* generator reproducible in three places (the
reference imported by make_goldens.py, the CPU oracle, and the CUDA product on the
GPU box) parameters are filled by tensor NAME with a per-name seeded CPU generator
instead of relying on module construction order.
"""
import math
import zlib

import torch

# Values recalled from upstream EG3D's FFHQ 512-128 config.
RENDERING_KWARGS = dict(
    depth_resolution=48, depth_resolution_importance=48, ray_start=2.25, ray_end=3.3, box_warp=1,
    disparity_space_sampling=False, clamp_mode='softplus',
    superresolution_module='training.superresolution.SuperresolutionHybrid8X',
    superresolution_noise_mode='none', sr_antialias=True, decoder_lr_mul=1,
    c_gen_conditioning_zero=False, c_scale=1, avg_camera_radius=2.7, avg_camera_pivot=[0, 0, 0.2],
    white_back=False,
)

G_KWARGS_FULL = dict(
    z_dim=512, c_dim=25, w_dim=512, img_resolution=512, img_channels=3, sr_num_fp16_res=4,
    mapping_kwargs={'num_layers': 2}, channel_base=32768, channel_max=512, num_fp16_res=4, conv_clamp=256,
    fused_modconv_default='inference_only',
    sr_kwargs=dict(channel_base=32768, channel_max=512, fused_modconv_default='inference_only'),
)

# Same topology with thin backbone channels (64,64,64,64,32,16,8): fast on CPU, exercises odd channel counts.
G_KWARGS_TINY = dict(G_KWARGS_FULL, channel_base=2048, channel_max=64)


def rendering_kwargs(**over):
    rk = dict(RENDERING_KWARGS)
    rk.update(over)
    return rk


def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) * 2654435761 + seed) % (2 ** 63))


def fill_params_(named_tensors, seed, noise_strength=0.1):
    """In-place fill of every parameter/buffer of a TriPlaneGenerator-shaped module, keyed by name."""
    with torch.no_grad():
        for name, t in sorted(named_tensors.items()):
            if name.endswith('resample_filter') or name.endswith('w_avg'):
                continue
            g = _gen(name, seed)
            r = torch.randn(t.shape, generator=g, dtype=torch.float32)
            if name.endswith('affine.bias'):
                v = 1.0 + 0.1 * r
            elif name.endswith('noise_strength'):
                v = noise_strength * (1.0 + 0.5 * torch.tanh(r))
            elif name.endswith('bias'):
                v = 0.1 * r
            elif '.mapping.' in name and name.endswith('weight') and 'embed' not in name:
                v = 100.0 * r          # lr_multiplier 0.01 => stored weights are randn / 0.01
            else:
                v = r                   # conv / torgb / affine / decoder weights, const, noise_const
            t.copy_(v.to(t.dtype))


def latent_ws(seed, n=1, num_ws=14, w_dim=512):
    return torch.randn(n, num_ws, w_dim, generator=torch.Generator().manual_seed(seed))


def camera(yaw=0.0, pitch=0.0, radius=2.7, focal=4.2647, n=1):
    """[n,25] label: look-at cam2world (OpenCV convention) on a sphere about the origin + normalised intrinsics.
    yaw = pitch = 0 gives the canonical camera of w_projector.py:79-84."""
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    origin = torch.tensor([radius * sy * cp, radius * sp, radius * cy * cp], dtype=torch.float64)
    fwd = -origin / origin.norm()
    down = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)
    x_axis = torch.linalg.cross(down, fwd)
    x_axis = x_axis / x_axis.norm()
    y_axis = torch.linalg.cross(fwd, x_axis)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x_axis, y_axis, fwd, origin
    # canonical check: yaw=pitch=0 -> origin (0,0,2.7), fwd (0,0,-1), cols: x=(1,0,0), y=(0,-1,0)
    intr = torch.tensor([focal, 0, 0.5, 0, focal, 0.5, 0, 0, 1], dtype=torch.float64)
    c = torch.cat([m.reshape(-1), intr]).to(torch.float32)
    return c[None].repeat(n, 1)


def targets(seed, r_raw):
    g = torch.Generator().manual_seed(seed)
    t512 = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    t_raw = torch.nn.functional.interpolate(t512, size=(r_raw, r_raw), mode='area')
    return t512, t_raw

"""Deterministic synthetic generator parameters and inputs (TEST INFRASTRUCTURE ONLY).

The reference's pickles are not available offline, so every parity test runs on a
random generator.  To make the *same* generator reproducible in three places (the
reference imported by make_goldens.py, the CPU oracle, and the CUDA product on the
GPU box) parameters are filled by tensor NAME with a per-name seeded CPU generator
instead of relying on module construction order.
"""
import contextlib
import math
import zlib

import torch

_TORCH_RANDN = torch.randn          # the real one: SeededNormal replaces the module attribute while it is active

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


G_KWARGS = {'full': G_KWARGS_FULL, 'tiny': G_KWARGS_TINY}

# The golden cases: shared by oracle/make_goldens.py (records them from the real reference) and tests/golden_util.py
# (rebuilds the inputs).  cam = (yaw, pitch); batch entry 1 (if any) uses (-yaw, pitch/2).
GOLDEN_CASES = {
    'tiny_r64_s16': dict(arch='tiny', R=64, S=16, S_imp=16, N=1, cam=(0.0, 0.0), bwd=True),
    'tiny_r32_s8_n2_white': dict(arch='tiny', R=32, S=8, S_imp=8, N=2, cam=(0.25, -0.15), bwd=True, rk={'white_back': True}),
    'tiny_r64_s12_noimp': dict(arch='tiny', R=64, S=12, S_imp=0, N=1, cam=(-0.3, 0.1), bwd=False),
    # stage-2 call shape of base_coach.py:163 (noise_mode='random') + the density-noise branch of renderer.py:201-202
    'tiny_r32_s8_random': dict(arch='tiny', R=32, S=8, S_imp=8, N=2, cam=(0.1, 0.2), bwd=True, noise_mode='random', rk={'density_noise': 0.05}),
    # renderer.py:146-152 ('auto' limits; focal 1.5 makes the outer rays miss the box) and :230-237 (disparity sampling)
    'tiny_r32_s8_auto': dict(arch='tiny', R=32, S=8, S_imp=8, N=2, cam=(0.3, -0.1), bwd=True, focal=1.5, rk={'ray_start': 'auto', 'ray_end': 'auto'}),
    'tiny_r32_s8_disp': dict(arch='tiny', R=32, S=8, S_imp=8, N=1, cam=(-0.2, 0.1), bwd=True, rk={'disparity_space_sampling': True}),
    'full_r64_s16': dict(arch='full', R=64, S=16, S_imp=16, N=1, cam=(0.0, 0.0), bwd=False),                   # BASELINE config 1
    'full_r64_s16_n2': dict(arch='full', R=64, S=16, S_imp=16, N=2, cam=(0.2, 0.1), bwd=True),                  # full architecture, batch 2
    'full_r128_s48': dict(arch='full', R=128, S=48, S_imp=48, N=1, cam=(0.3, -0.2), bwd=True),                  # BASELINE config 2/4 (grads incl. pose)
    'full_r256_s96': dict(arch='full', R=256, S=96, S_imp=96, N=1, cam=(-0.2, 0.15), bwd=True),                 # BASELINE config 5
}


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


def case_camera(cfg):
    yaw, pitch = cfg['cam']
    c = camera(yaw, pitch, n=cfg['N'], focal=cfg.get('focal', 4.2647))
    if cfg['N'] > 1:   # make the batch entries different
        c[1] = camera(-yaw, pitch * 0.5, focal=cfg.get('focal', 4.2647))[0]
    return c


class SeededNormal:
    """Replacement for torch.randn / torch.randn_like drawing from one seeded CPU generator, recording the shapes in call
    order.  Used (a) by make_goldens.py around the reference's synthesis() and (b) by the GPU tests around the product's,
    so that noise_mode='random' and density_noise see identical draws on both sides."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(int(seed))
        self.shapes = []

    def randn(self, *size, device=None, dtype=None, **_):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (list, tuple, torch.Size)) else tuple(size)
        self.shapes.append(list(shape))
        t = _TORCH_RANDN(shape, generator=self.g, dtype=torch.float32)
        return t.to(device=device, dtype=dtype or torch.float32)

    def randn_like(self, x, **_):
        return self.randn(x.shape, device=x.device, dtype=x.dtype)

    @contextlib.contextmanager
    def patched(self, enable=True):
        if not enable:
            yield self
            return
        old = torch.randn, torch.randn_like
        torch.randn, torch.randn_like = self.randn, self.randn_like
        try:
            yield self
        finally:
            torch.randn, torch.randn_like = old


def noise_layer_prefixes(gk=None, img_resolution=256):
    """Backbone SynthesisLayers in execution order (networks_stylegan2.py:503-518): b4.conv1, b8.conv0, b8.conv1, ..."""
    out, res = [], 4
    while res <= img_resolution:
        if res > 4:
            out.append(f'backbone.synthesis.b{res}.conv0.')
        out.append(f'backbone.synthesis.b{res}.conv1.')
        res *= 2
    return out


def replay_normal_draws(seed, shapes, gk=None):
    """Re-draw what SeededNormal produced for `shapes` and sort it into the oracle's arguments:
    ({layer prefix -> [N,1,res,res]} or None, (coarse density draw, fine density draw))."""
    g = torch.Generator().manual_seed(int(seed))
    draws = [_TORCH_RANDN(tuple(s), generator=g, dtype=torch.float32) for s in shapes]
    layer = [d for d in draws if d.ndim == 4 and d.shape[1] == 1 and d.shape[2] == d.shape[3]]
    dens = [d for d in draws if not (d.ndim == 4 and d.shape[1] == 1 and d.shape[2] == d.shape[3])]
    nr = dict(zip(noise_layer_prefixes(gk), layer)) if layer else None
    dd = (dens[0], dens[1] if len(dens) > 1 else None) if dens else (None, None)
    return nr, dd


def grad_slices(g, max_samples=2048):
    """Compact but channel-resolved summary of one gradient tensor (float64 numpy): L2 norm of every dim-0 slice, of every
    dim-1 slice, and a strided sample of <= max_samples elements."""
    g = g.detach().double().cpu()
    if g.ndim == 0:
        return g.abs().reshape(1).numpy(), torch.zeros(0).numpy(), g.reshape(1).numpy()
    oc = g.reshape(g.shape[0], -1).norm(dim=1).numpy()
    ic = g.transpose(0, 1).reshape(g.shape[1], -1).norm(dim=1).numpy() if g.ndim >= 2 else torch.zeros(0).numpy()
    flat = g.reshape(-1)
    stride = max(1, flat.numel() // max_samples)
    return oc, ic, flat[::stride][:max_samples].numpy()

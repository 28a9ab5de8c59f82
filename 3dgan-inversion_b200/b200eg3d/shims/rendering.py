"""Look-alikes of training/volumetric_rendering/{ray_marcher,renderer,ray_sampler}.py on the b200eg3d kernels."""
import torch

from .. import ops
from .._lib import call, device_guard, ptr, stream
from ..generator import ImportanceRenderer, OSGDecoder, RaySampler  # noqa: F401  (re-exported under the reference's paths)


class _Composite(torch.autograd.Function):
    """rgb, depth of MipRayMarcher2.run_forward (ray_marcher.py:25-57) for samples already in marching order."""

    @staticmethod
    @device_guard
    def forward(ctx, colors, densities, depths, white_back):
        c, s, t = ops._f32c(colors), ops._f32c(densities), ops._f32c(depths)
        n, m, S, ch = c.shape
        if ch != 32:
            raise NotImplementedError('b200eg3d MipRayMarcher2 shim: 32 feature channels (the EG3D tri-plane decoder)')
        dev = c.device
        st = stream()
        mm = torch.zeros([2], device=dev, dtype=torch.int32)
        mm[:1].fill_(-1)
        call('b200_depth_minmax', ptr(t), t.numel(), ptr(mm), st)
        feat = torch.empty([n, m, 32], device=dev, dtype=torch.float32)
        depth = torch.empty([n, m, 1], device=dev, dtype=torch.float32)
        wsum = torch.empty([n, m, 1], device=dev, dtype=torch.float32)
        tt, ss = t.reshape(n, m, S), s.reshape(n, m, S)
        call('b200_ray_composite_fwd', ptr(tt), ptr(ss), ptr(c), S, None, None, None, 0, ptr(mm), int(bool(white_back)), n * m, ptr(feat), ptr(depth),
             ptr(wsum), st)
        ctx.cfg = (int(bool(white_back)), S)
        ctx.save_for_backward(tt, ss, c, mm)
        return feat, depth

    @staticmethod
    @device_guard
    def backward(ctx, d_feat, d_depth):
        tt, ss, c, mm = ctx.saved_tensors
        white_back, S = ctx.cfg
        n, m = tt.shape[:2]
        if d_feat is None:
            d_feat = torch.zeros([n, m, 32], device=c.device, dtype=torch.float32)
        d_c, d_s = torch.empty_like(c), torch.empty_like(ss)
        call('b200_ray_composite_bwd', ptr(tt), ptr(ss), ptr(c), S, None, None, None, 0, ptr(mm), white_back, n * m, ptr(ops._f32c(d_feat)),
             ptr(ops._f32c(d_depth)) if d_depth is not None else None, None, ptr(d_c), ptr(d_s), None, None, stream())
        return d_c, d_s.reshape(n, m, S, 1), None, None


class MipRayMarcher2(torch.nn.Module):
    """ray_marcher.py:20-63.  The colours -- 32 channels x S samples per ray, the heavy operand -- go through the fused per-ray
    kernel; the per-interval `weights` (third output; a few floats per ray) are formed with the same statements as the reference so
    that callers which differentiate through them keep working."""

    def run_forward(self, colors, densities, depths, rendering_options):
        if rendering_options.get('clamp_mode', 'softplus') != 'softplus':
            raise AssertionError('MipRayMarcher only supports `clamp_mode`=`softplus`!')
        rgb, depth = _Composite.apply(colors, densities, depths, rendering_options.get('white_back', False))
        deltas = depths[:, :, 1:] - depths[:, :, :-1]
        dens_mid = torch.nn.functional.softplus((densities[:, :, :-1] + densities[:, :, 1:]) / 2 - 1)
        alpha = 1 - torch.exp(-dens_mid * deltas)
        alpha_shifted = torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2)
        weights = alpha * torch.cumprod(alpha_shifted, -2)[:, :, :-1]
        return rgb, depth, weights

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)


def sample_from_planes(plane_axes, plane_features, coordinates, mode='bilinear', padding_mode='zeros', box_warp=None):
    """renderer.py:55-66 is fused with the decoder in b200eg3d (ImportanceRenderer.run_model); the un-fused sampler has no
    look-alike: its [N, 3, M, 32] output is exactly the 300 MB tensor the fused kernel exists to avoid."""
    raise NotImplementedError('b200eg3d fuses sample_from_planes into ImportanceRenderer.run_model; call that instead')

"""Dense density queries for shape extraction.

The reference extracts a mesh by evaluating sigma on a 256^3 / 512^3 voxel grid in chunks of 1e6 points and re-runs the whole
StyleGAN2 backbone for every chunk (training/coaches/single_id_coach.py:118-140, gen_videos.py:140-158: 17 / 135 backbone
passes per shape).  Here the planes are synthesised once and every chunk is one launch of the fused tri-plane sampler +
decoder in its density-only mode (no rgb stores, only column 0 of the second decoder layer).

    create_samples(N, voxel_origin, cube_length)       single_id_coach.py:165-188, same fp32 arithmetic, built on the device
    query_sigma(G, ws, samples, max_batch)             the `while head < samples.shape[1]` loop of create_geometry
    density_grid(G, ws, shape_res)                     samples + query + reshape / flip / border trim (:141-153)
"""
import numpy as np
import torch

from . import ops
from ._lib import call, on_device, ptr, stream


def create_samples(N=256, voxel_origin=[0, 0, 0], cube_length=2.0, device='cuda'):
    """Same statements as the reference, evaluated on `device` (note its float index arithmetic: y and x carry the
    fractional part of the flat index, and indices above 2^24 round in fp32 -- reproduced, not corrected)."""
    voxel_origin = np.array(voxel_origin) - cube_length / 2
    voxel_size = cube_length / (N - 1)
    overall_index = torch.arange(0, N ** 3, 1, dtype=torch.long, device=device)
    samples = torch.zeros(N ** 3, 3, device=device)
    samples[:, 2] = overall_index % N
    samples[:, 1] = (overall_index.float() / N) % N
    samples[:, 0] = ((overall_index.float() / N) / N) % N
    samples[:, 0] = (samples[:, 0] * voxel_size) + voxel_origin[2]
    samples[:, 1] = (samples[:, 1] * voxel_size) + voxel_origin[1]
    samples[:, 2] = (samples[:, 2] * voxel_size) + voxel_origin[0]
    return samples.unsqueeze(0), voxel_origin, voxel_size


@torch.no_grad()
def query_sigma(G, ws, samples, max_batch=1 << 24, planes=None):
    """sigma [N, P, 1] of sample_mixed(samples, ., ws, noise_mode='const') for all P points; the backbone runs once."""
    with on_device(ws):
        return _query_sigma(G, ws, samples, max_batch, planes)


def _query_sigma(G, ws, samples, max_batch, planes):
    if planes is None:
        planes = G.backbone.synthesis(ws, update_emas=False, noise_mode='const')
    pl = G.renderer._planes_nhwc(planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1]))
    pl = pl.detach().to(torch.float32).contiguous()
    n, hp, wp, _ = pl.shape
    co = samples.detach().to(torch.float32)
    if co.shape[0] != n:
        co = co.expand(n, -1, -1)
    P = co.shape[1]
    W1, b1, W2, b2, lr_mul = ops._decoder_params(G.decoder)
    w = [t.detach().to(torch.float32).contiguous() for t in (W1, b1, W2, b2)]
    sigmas = torch.empty([n, P, 1], device=pl.device, dtype=torch.float32)
    box = float(G.rendering_kwargs['box_warp'])
    if n == 1:                                      # chunks are contiguous slices: no copies
        head = 0
        while head < P:
            cnt = min(max_batch, P - head)
            c = co[:, head:head + cnt].contiguous()
            call('b200_triplane_mlp_fwd', ptr(pl), 1, hp, wp, ptr(c), None, None, None, 0, 0, cnt, box, *map(ptr, w), float(lr_mul),
                 None, ptr(sigmas[:, head:head + cnt]), None, stream())
            head += cnt
    else:
        call('b200_triplane_mlp_fwd', ptr(pl), n, hp, wp, ptr(co.contiguous()), None, None, None, 0, 0, P, box, *map(ptr, w), float(lr_mul),
             None, ptr(sigmas), None, stream())
    return sigmas


@torch.no_grad()
def density_grid(G, ws, shape_res=512, pad_value=-1000.0, trim=True, max_batch=1 << 24):
    """create_geometry of single_id_coach.py:118-153 up to (not including) the file export: [res, res, res] fp32 on the device."""
    samples, voxel_origin, voxel_size = create_samples(N=shape_res, voxel_origin=[0, 0, 0],
                                                       cube_length=G.rendering_kwargs['box_warp'] * 1, device=ws.device)
    sig = query_sigma(G, ws, samples, max_batch=max_batch).reshape(shape_res, shape_res, shape_res).flip(0)
    if trim:
        pad = int(30 * shape_res / 256)
        sig[:pad] = pad_value
        sig[-pad:] = pad_value
        sig[:, :pad] = pad_value
        sig[:, -pad:] = pad_value
        sig[:, :, :pad] = pad_value
        sig[:, :, -pad:] = pad_value
    return sig

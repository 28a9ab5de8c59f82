"""Shape-extraction density queries: voxel sample grid and chunked sigma queries with cached planes."""
import os

import numpy as np
import pytest
import torch

import eg3d_oracle as oracle
import stage1_oracle as s1
import synth_params as sp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stage1_warp.npz')


def test_oracle_create_samples_matches_reference_fixture():
    fx = np.load(GOLD)
    a, origin, size = s1.create_samples(N=16, voxel_origin=[0, 0, 0], cube_length=1.0)
    assert np.array_equal(a[0].numpy(), fx['samples_n16'])
    b, _, _ = s1.create_samples(N=256, voxel_origin=[0, 0, 0], cube_length=1.0)
    assert np.array_equal(b[0, ::4099].numpy(), fx['samples_n256'])
    assert abs(size - 1.0 / 15) < 1e-12 and np.allclose(origin, -0.5)


@pytest.fixture(scope='module')
def b2():
    import b200eg3d
    assert torch.cuda.is_available()
    b200eg3d.ops.library_info()
    return b200eg3d


@pytest.mark.gpu
def test_create_samples_on_device_is_bit_identical(b2):
    fx = np.load(GOLD)
    a, _, _ = b2.geometry.create_samples(N=16, voxel_origin=[0, 0, 0], cube_length=1.0)
    assert np.array_equal(a[0].cpu().numpy(), fx['samples_n16'])
    b, _, _ = b2.geometry.create_samples(N=256, voxel_origin=[0, 0, 0], cube_length=1.0)
    assert np.array_equal(b[0, ::4099].cpu().numpy(), fx['samples_n256'])


@pytest.mark.gpu
@pytest.mark.parametrize('max_batch', [1 << 24, 1000, 4096])
def test_query_sigma_matches_sample_mixed_and_oracle(b2, max_batch):
    """Chunked density-only queries == sample_mixed()['sigma'] == the oracle's run_model on the oracle's planes."""
    rk = sp.rendering_kwargs()
    G = b2.TriPlaneGenerator(rendering_kwargs=rk, **sp.G_KWARGS_TINY).eval()
    named = dict(list(G.named_parameters()) + list(G.named_buffers()))
    sp.fill_params_(named, 3)
    P = {k: v.detach().clone() for k, v in named.items()}
    G = G.cuda()
    ws = sp.latent_ws(4)
    samples, _, _ = s1.create_samples(N=24, voxel_origin=[0, 0, 0], cube_length=rk['box_warp'])
    p96 = oracle.backbone_synthesis(P, ws, noise_mode='const')
    planes_ref = p96.reshape(1, 3, 32, p96.shape[-2], p96.shape[-1])
    sig = b2.geometry.query_sigma(G, ws.cuda(), samples.cuda(), max_batch=max_batch)
    full = G.sample_mixed(samples.cuda(), torch.zeros_like(samples).cuda(), ws.cuda(), noise_mode='const')['sigma']
    assert sig.shape == (1, 24 ** 3, 1)
    assert (sig - full).abs().max().item() < 1e-4          # two backbone passes: split-K atomics make the planes reproducible to ~1e-5
    _, sig_ref = oracle.run_model(P, planes_ref, samples, rk)
    assert (sig.cpu() - sig_ref).abs().max().item() < 1e-3


@pytest.mark.gpu
def test_density_grid_layout(b2):
    rk = sp.rendering_kwargs()
    G = b2.TriPlaneGenerator(rendering_kwargs=rk, **sp.G_KWARGS_TINY).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), 3)
    G = G.cuda()
    ws = sp.latent_ws(4).cuda()
    res = 32
    grid = b2.geometry.density_grid(G, ws, shape_res=res)
    assert grid.shape == (res, res, res)
    pad = int(30 * res / 256)
    assert float(grid[:pad].max()) == -1000.0 and float(grid[:, :, -pad:].max()) == -1000.0
    samples, _, _ = b2.geometry.create_samples(N=res, cube_length=rk['box_warp'])
    raw = b2.geometry.query_sigma(G, ws, samples).reshape(res, res, res).flip(0)
    inner = (slice(pad, res - pad),) * 3
    assert (grid[inner] - raw[inner]).abs().max().item() < 1e-4

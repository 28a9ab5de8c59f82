"""Size-independent properties of the CPU oracle (the checker the GPU parity tests rely on): cheap invariants that would
expose a wrong restatement independently of the reference fixtures."""
import math

import pytest
import torch

import eg3d_oracle as oracle
import stage1_oracle as s1


def test_upfirdn_is_linear_and_preserves_dc():
    g = torch.Generator().manual_seed(0)
    f = oracle.fir_1331()
    a, b = torch.randn(1, 3, 9, 9, generator=g), torch.randn(1, 3, 9, 9, generator=g)
    up = lambda t: oracle.upsample2d(t, f)
    assert torch.allclose(up(2.5 * a - b), 2.5 * up(a) - up(b), atol=1e-5)
    c = up(torch.ones(1, 1, 12, 12))
    assert c.shape == (1, 1, 24, 24) and torch.allclose(c[:, :, 2:-2, 2:-2], torch.ones(1, 1, 20, 20), atol=1e-6)   # gain 4 restores DC


def test_ray_march_weights_form_a_sub_probability():
    g = torch.Generator().manual_seed(1)
    n, m, s = 1, 50, 24
    depths = (torch.rand(n, m, s, 1, generator=g) * 0.9 + 2.3).sort(dim=2).values
    sig = torch.randn(n, m, s, 1, generator=g) * 4
    col = torch.rand(n, m, s, 5, generator=g)
    rgb, depth, w = oracle.ray_march(col, sig, depths)
    assert (w >= 0).all() and (w.sum(2) <= 1 + 1e-5).all()
    assert (depth >= depths.min() - 1e-6).all() and (depth <= depths.max() + 1e-6).all()
    assert (rgb >= -1 - 1e-5).all() and (rgb <= 1 + 1e-5).all()                  # colours in [0,1] -> composite*2-1 in [-1,1]
    rgb_wb, _, _ = oracle.ray_march(col, sig, depths, white_back=True)
    assert torch.allclose(rgb_wb, rgb + 2 * (1 - w.sum(2)), atol=1e-5)           # white background adds the missing opacity


def test_importance_depths_stay_inside_the_coarse_interval_and_follow_the_weights():
    g = torch.Generator().manual_seed(2)
    n, m, s, k = 1, 40, 32, 64
    t = oracle.stratified_depths(n, m, s, 2.25, 3.3, torch.rand(n, m, s, 1, generator=g), torch.float32)
    w = torch.zeros(n, m, s - 1, 1)
    w[:, :, 10] = 1.0                                                             # all mass in one interval
    tf = oracle.importance_depths(t, w, k, torch.rand(n * m, k, generator=g))
    assert tf.shape == (n, m, k, 1)
    assert (tf >= t.min(dim=2, keepdim=True).values - 1e-6).all() and (tf <= t.max(dim=2, keepdim=True).values + 1e-6).all()
    mid = 0.5 * (t[:, :, 10] + t[:, :, 11])
    assert ((tf - mid.unsqueeze(2)).abs().median() < 3 * (3.3 - 2.25) / (s - 1))  # samples concentrate around the heavy interval


def test_ray_sampler_geometry():
    import synth_params as sp
    c = sp.camera(0.2, -0.1)
    o, d = oracle.ray_sampler(c[:, :16].reshape(1, 4, 4), c[:, 16:].reshape(1, 3, 3), 16)
    assert torch.allclose(d.norm(dim=-1), torch.ones(1, 256), atol=1e-6)
    assert torch.allclose(o, c[:, :16].reshape(1, 4, 4)[:, :3, 3].unsqueeze(1).expand(-1, 256, -1))
    centre = d.reshape(16, 16, 3)[7:9, 7:9].mean((0, 1))
    assert torch.allclose(torch.nn.functional.normalize(centre, dim=0), -torch.nn.functional.normalize(o[0, 0], dim=0), atol=2e-2)   # looks at the origin


def test_noise_regulariser_is_zero_for_uncorrelated_limits_and_scale_quartic():
    g = torch.Generator().manual_seed(3)
    b = [torch.randn(64, 64, generator=g), torch.randn(16, 16, generator=g)]
    r1 = s1.noise_regularizer(b)
    r2 = s1.noise_regularizer([2 * x for x in b])
    assert torch.allclose(r2, 16 * r1, rtol=1e-5)                                # (mean of products)^2 -> 4th power of the scale
    const = s1.noise_regularizer([torch.ones(32, 32)])
    assert abs(float(const) - 2 * 3) < 1e-5                                      # levels 32,16,8: each contributes 1^2 + 1^2
    nrm = s1.normalize_noise(b)
    for x in nrm:
        assert abs(float(x.mean())) < 1e-6 and abs(float(x.square().mean()) - 1) < 1e-5


def test_calc_loss_terms():
    g = torch.Generator().manual_seed(4)
    real = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
    out = {'image': real.clone(), 'image_raw': torch.nn.functional.interpolate(real, size=(8, 8), mode='area'), 'image_depth': torch.full((1, 1, 8, 8), 2.7)}
    loss, parts = oracle.calc_loss(out, real)
    assert float(loss) < 1e-12                                                   # perfect reconstruction, flat depth
    out['image_depth'] = torch.arange(64.).reshape(1, 1, 8, 8)
    _, parts = oracle.calc_loss(out, real)
    assert abs(float(parts[2]) - (1 + 64)) < 1e-4                                # forward differences: 1 along x, 8 along y

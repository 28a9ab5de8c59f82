"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (where the CPU oracle would take minutes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def b2():
    import b200eg3d
    assert torch.cuda.is_available()
    b200eg3d.ops.library_info()
    return b200eg3d


def test_fir_upsample_is_linear_and_keeps_dc_at_512(b2):
    g = torch.Generator().manual_seed(0)
    f = b2.ops.setup_filter([1, 3, 3, 1]).cuda()
    a, b = torch.randn(1, 64, 256, 256, generator=g).cuda(), torch.randn(1, 64, 256, 256, generator=g).cuda()
    up = lambda t: b2.ops.upsample2d(t, f)
    lhs, rhs = up(1.5 * a - b), 1.5 * up(a) - up(b)
    assert lhs.shape == (1, 64, 512, 512) and (lhs - rhs).abs().max().item() < 1e-5
    c = up(torch.ones(1, 8, 256, 256, device='cuda'))
    assert (c[:, :, 2:-2, 2:-2] - 1).abs().max().item() < 1e-6


@pytest.mark.parametrize('R,S', [(128, 48), (256, 96)])
def test_render_invariants_at_full_size(b2, R, S):
    """Composite weights are a sub-probability, depth stays inside the global sample range, features inside [-1, 1]
    (sigmoid colours), and the render is deterministic for fixed noise draws -- on random planes at the metric's shapes."""
    import synth_params as sp
    g = torch.Generator().manual_seed(R)
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S)
    planes = (torch.randn(1, 256, 256, 96, generator=g) * 0.5).cuda()
    dec = b2.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    with torch.no_grad():
        for p in dec.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.ndim == 1 else 1.0))
    dec = dec.cuda()
    c = sp.camera(0.3, -0.2).cuda()
    ro, rd = b2.RaySampler()(c[:, :16].view(1, 4, 4), c[:, 16:25].view(1, 3, 3), R)
    M = R * R
    u1, u2 = torch.rand(1, M, S, 1, generator=g).cuda(), torch.rand(M, S, generator=g).cuda()
    t_base = torch.linspace(rk['ray_start'], rk['ray_end'], S, device='cuda')
    delta = (rk['ray_end'] - rk['ray_start']) / (S - 1)
    with torch.no_grad():
        feat, depth, wsum = b2.ops.render(planes, dec, ro, rd, rk['box_warp'], t_base, delta, u1, u2)
        feat2, depth2, _ = b2.ops.render(planes, dec, ro, rd, rk['box_warp'], t_base, delta, u1, u2)
    assert feat.shape == (1, M, 32) and depth.shape == (1, M, 1)
    assert torch.isfinite(feat).all() and torch.isfinite(depth).all()
    assert wsum.min().item() >= 0 and wsum.max().item() <= 1 + 1e-5
    assert depth.min().item() >= rk['ray_start'] - 1e-4 and depth.max().item() <= rk['ray_end'] + delta + 1e-4
    assert feat.min().item() >= -1.01 and feat.max().item() <= 1.01
    assert torch.equal(feat, feat2) and torch.equal(depth, depth2)             # no atomics on the forward render path


def test_density_grid_matches_pointwise_queries_at_256(b2):
    import synth_params as sp
    rk = sp.rendering_kwargs()
    G = b2.TriPlaneGenerator(rendering_kwargs=rk, **sp.G_KWARGS_FULL).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), 7)
    G = G.cuda()
    ws = sp.latent_ws(1).cuda()
    with torch.no_grad():
        planes = G.backbone.synthesis(ws, noise_mode='const')
        samples, _, _ = b2.geometry.create_samples(N=256, cube_length=rk['box_warp'])
        a = b2.geometry.query_sigma(G, ws, samples, planes=planes)                # one launch
        b = b2.geometry.query_sigma(G, ws, samples, max_batch=1_000_000, planes=planes)   # the reference's chunking
        sub = samples[:, ::4099]
        c = G.renderer.run_model(planes.view(1, 3, 32, 256, 256), G.decoder, sub, None, rk)['sigma']
    assert a.shape == (1, 256 ** 3, 1) and torch.equal(a, b)
    assert (a[:, ::4099] - c).abs().max().item() < 1e-5                          # density-only mode == full decoder, same planes

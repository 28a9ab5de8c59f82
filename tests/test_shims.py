"""Op-level seam (second row): the look-alikes of torch_utils.ops.* and of the renderer classes.

CPU: install() / uninstall() route the reference's import paths.  GPU: every look-alike against plain ATen statements of what
the reference function computes (file:line in the look-alike's docstring), the way an un-rebuilt pickle's embedded source would
call them -- including both branches of modulated_conv2d (networks_stylegan2.py:58-91) -- and the generator's
fused_modconv=False path against its fused kernels."""
import sys

import pytest
import torch
import torch.nn.functional as F

import eg3d_oracle as oracle


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_install_routes_reference_import_paths():
    import b200eg3d.shims as shims
    had = 'torch_utils' in sys.modules
    paths = shims.install()
    try:
        assert 'torch_utils.ops.conv2d_resample' in paths
        from torch_utils.ops import bias_act, conv2d_gradfix, conv2d_resample, fma, upfirdn2d      # networks_stylegan2.py:17-21
        from training.volumetric_rendering.ray_marcher import MipRayMarcher2                    # renderer.py:20
        from training.volumetric_rendering.renderer import ImportanceRenderer                    # triplane.py:14
        from training.volumetric_rendering.ray_sampler import RaySampler                         # triplane.py:15
        import b200eg3d
        assert ImportanceRenderer is b200eg3d.ImportanceRenderer and RaySampler is b200eg3d.RaySampler
        assert MipRayMarcher2.__module__.startswith('b200eg3d')
        for mod, names in ((bias_act, ['bias_act']), (upfirdn2d, ['upfirdn2d', 'setup_filter', 'filter2d', 'upsample2d', 'downsample2d']),
                           (conv2d_resample, ['conv2d_resample']), (conv2d_gradfix, ['conv2d', 'conv_transpose2d', 'no_weight_gradients']),
                           (fma, ['fma'])):
            for n in names:
                assert getattr(mod, n).__module__.startswith('b200eg3d'), (mod.__name__, n)
        assert shims.install() == paths                                      # idempotent
    finally:
        shims.uninstall()
    if not had:
        assert 'torch_utils' not in sys.modules
    with pytest.raises(RuntimeError, match='no CPU path'):                   # the look-alikes have no reference fallback
        from b200eg3d.shims import ops_modules
        ops_modules.conv2d_resample(torch.zeros(1, 8, 4, 4), torch.zeros(8, 8, 3, 3), padding=1)


@pytest.fixture(scope='module')
def om():
    import b200eg3d
    from b200eg3d.shims import ops_modules
    assert torch.cuda.is_available()
    return ops_modules


def _modulated_conv2d_via(ops_ns, x, weight, styles, noise, up, f, fused):
    """The call pattern of modulated_conv2d (networks_stylegan2.py:34-91) against an operator namespace exposing
    conv2d_resample / fma: fused = grouped convolution with per-sample weights (:81-90), else activation scaling (:70-79)."""
    n, cin = x.shape[0], x.shape[1]
    cout, _, kh, kw = weight.shape
    w = weight.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)
    dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    flip = up == 1
    if not fused:
        y = ops_ns.conv2d_resample(x=x * styles.reshape(n, -1, 1, 1), w=weight, f=f, up=up, padding=kh // 2, flip_weight=flip)
        return ops_ns.fma(y, dcoefs.reshape(n, -1, 1, 1), noise) if noise is not None else y * dcoefs.reshape(n, -1, 1, 1)
    w = w * dcoefs.reshape(n, -1, 1, 1, 1)
    y = ops_ns.conv2d_resample(x=x.reshape(1, -1, *x.shape[2:]), w=w.reshape(-1, cin, kh, kw), f=f, up=up, padding=kh // 2, groups=n, flip_weight=flip)
    y = y.reshape(n, -1, *y.shape[2:])
    return y + noise if noise is not None else y


@pytest.mark.gpu
@pytest.mark.parametrize('fused', [True, False])
@pytest.mark.parametrize('cin,cout,res,up', [(16, 24, 8, 1), (32, 16, 8, 2), (12, 20, 5, 1), (64, 64, 16, 2)])
def test_modulated_conv2d_through_the_shims(om, cin, cout, res, up, fused):
    g = torch.Generator().manual_seed(cin * 100 + res + up)
    n = 2
    x, wgt = torch.randn(n, cin, res, res, generator=g), torch.randn(cout, cin, 3, 3, generator=g)
    st, nz = torch.randn(n, cin, generator=g) + 1, torch.randn(res * up, res * up, generator=g) * 0.1
    dy = torch.randn(n, cout, res * up, res * up, generator=g)
    ref_in = [t.clone().requires_grad_(True) for t in (x, wgt, st)]
    y_ref = oracle.modulated_conv2d(ref_in[0], ref_in[1], ref_in[2], noise=nz, up=up, f=oracle.fir_1331())
    (y_ref * dy).sum().backward()
    dev_in = [t.clone().cuda().requires_grad_(True) for t in (x, wgt, st)]
    y = _modulated_conv2d_via(om, dev_in[0], dev_in[1], dev_in[2], nz.cuda(), up, om.setup_filter([1, 3, 3, 1]).cuda(), fused)
    (y * dy.cuda()).sum().backward()
    assert relerr(y, y_ref) < 2e-5
    for a, b, name in zip(dev_in, ref_in, ('dx', 'dW', 'dstyles')):
        # weight gradients run single-pass bf16 on the tensor cores; in the fused form the style gradient flows through them
        assert relerr(a.grad, b.grad) < (5e-3 if (name == 'dW' or (fused and name == 'dstyles')) else 2e-4), name


@pytest.mark.gpu
def test_torgb_pattern_filter2d_downsample2d_fma(om):
    g = torch.Generator().manual_seed(3)
    x, w = torch.randn(2, 32, 16, 16, generator=g), torch.randn(3, 32, 1, 1, generator=g)
    y = om.conv2d_resample(x.cuda(), w.cuda())                                               # 1x1, networks_stylegan2.py:355 (demodulate=False)
    assert relerr(y, F.conv2d(x, w)) < 2e-5
    y = om.conv2d_resample(x.cuda(), w.cuda(), f=om.setup_filter([1, 3, 3, 1]).cuda(), up=2)   # conv2d_resample.py:99-102
    assert relerr(y, oracle.upfirdn2d(F.conv2d(x, w), oracle.fir_1331(), up=2, pad=(2, 1, 2, 1), gain=4.0)) < 2e-5
    f = om.setup_filter([1, 3, 3, 1])
    assert relerr(om.filter2d(x.cuda(), f.cuda()), oracle.upfirdn2d(x, f, pad=(2, 1, 2, 1))) < 1e-5         # upfirdn2d.py:303-310
    assert relerr(om.downsample2d(x.cuda(), f.cuda()), oracle.upfirdn2d(x, f, down=2, pad=(1, 1, 1, 1))) < 1e-5   # upfirdn2d.py:377-384
    a, b, c = (torch.randn(2, 8, 4, 4, generator=g).cuda().requires_grad_(True), torch.randn(2, 8, 1, 1, generator=g).cuda().requires_grad_(True),
               torch.randn(4, 4, generator=g).cuda().requires_grad_(True))
    out = om.fma(a, b, c)
    assert torch.allclose(out, a * b + c)
    out.square().sum().backward()
    ar, br, cr = (t.detach().clone().requires_grad_(True) for t in (a, b, c))
    (ar * br + cr).square().sum().backward()
    for t, r in ((a, ar), (b, br), (c, cr)):
        assert t.grad.shape == r.grad.shape and torch.allclose(t.grad, r.grad, rtol=1e-5, atol=1e-5)
    with om.no_weight_gradients():
        wq = w.clone().cuda().requires_grad_(True)
        xq = x.clone().cuda().requires_grad_(True)
        om.conv2d(xq, wq).sum().backward()
        assert wq.grad is None and xq.grad is not None                                        # conv2d_gradfix.py:27-34
    with pytest.raises(NotImplementedError):
        om.conv2d_resample(x.cuda(), torch.randn(8, 32, 3, 3).cuda(), f=f.cuda(), down=2)


@pytest.mark.gpu
@pytest.mark.parametrize('white', [False, True])
def test_mip_ray_marcher2_shim(om, white):
    from b200eg3d.shims.rendering import MipRayMarcher2
    g = torch.Generator().manual_seed(7)
    n, m, S = 2, 50, 24
    col, den = torch.rand(n, m, S, 32, generator=g), torch.randn(n, m, S, 1, generator=g) * 2
    dep = (torch.rand(n, m, S, 1, generator=g) + 2).sort(dim=2).values
    opts = {'clamp_mode': 'softplus', 'white_back': white}
    d_rgb, d_dep, d_w = torch.randn(n, m, 32, generator=g), torch.randn(n, m, 1, generator=g), torch.randn(n, m, S - 1, 1, generator=g)
    ref = [t.clone().requires_grad_(True) for t in (col, den)]
    rgb_r, dep_r, w_r = oracle.ray_march(ref[0], ref[1], dep, white)
    ((rgb_r * d_rgb).sum() + (dep_r * d_dep).sum() + (w_r * d_w).sum()).backward()
    dev = [t.clone().cuda().requires_grad_(True) for t in (col, den)]
    rgb, depth, w = MipRayMarcher2()(dev[0], dev[1], dep.cuda(), opts)
    ((rgb * d_rgb.cuda()).sum() + (depth * d_dep.cuda()).sum() + (w * d_w.cuda()).sum()).backward()
    assert relerr(rgb, rgb_r) < 2e-5 and relerr(depth, dep_r) < 2e-6 and relerr(w, w_r) < 2e-5
    assert relerr(dev[0].grad, ref[0].grad) < 1e-4 and relerr(dev[1].grad, ref[1].grad) < 1e-4


@pytest.mark.gpu
def test_generator_unfused_path_matches_fused(golden_dir):
    """fused_modconv=False (networks_stylegan2.py:70-79) runs on the look-alikes and must agree with the fused kernels and the fixture."""
    import numpy as np
    import synth_params as sp
    from golden_util import load_case
    import b200eg3d
    case = load_case(golden_dir, 'tiny_r64_s16')
    G = b200eg3d.TriPlaneGenerator(rendering_kwargs=case.rk, **case.gk).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), case.param_seed)
    G = G.cuda().requires_grad_(True)
    G.neural_rendering_resolution = case.R
    G.renderer.fixed_noise = (case.u_strat, case.u_imp)
    ws = case.ws.cuda().requires_grad_(True)
    outs, grads = [], []
    for fused in (True, False):
        G.zero_grad(set_to_none=True)
        ws.grad = None
        out = G.synthesis(ws, case.c.cuda(), noise_mode='const', force_fp32=True, fused_modconv=fused)
        oracle.pti_loss(out, case.t512.cuda(), case.t_raw.cuda()).backward()
        outs.append({k: v.detach().clone() for k, v in out.items()})
        grads.append((ws.grad.clone(), G.backbone.synthesis.b16.conv1.weight.grad.clone(), G.backbone.synthesis.b32.conv0.affine.weight.grad.clone()))
    for k in ('image', 'image_raw', 'image_depth'):
        assert (outs[0][k] - outs[1][k]).abs().max().item() < 2e-4, k
    assert np.abs(outs[1]['image_raw'].cpu().numpy() - case.fx['image_raw']).max() < 1e-3
    for a, b in zip(*grads):
        assert relerr(a, b) < 5e-3

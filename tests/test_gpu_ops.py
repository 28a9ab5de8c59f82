"""GPU parity of the individual CUDA operators (called through the C-ABI) against the CPU oracle on seeded inputs."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import eg3d_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def b2():
    import b200eg3d
    assert torch.cuda.is_available()
    b200eg3d.ops.library_info()          # raises if the .so is missing
    return b200eg3d


def gen(seed):
    return torch.Generator().manual_seed(seed)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def maxdiff(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def relerr(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ---------------------------------------------------------------------------------------------- modconv layer

@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('cin,cout,res,up', [(8, 16, 8, 1), (16, 8, 8, 2), (64, 64, 16, 1), (64, 32, 16, 2), (12, 20, 5, 1),
                                             (20, 12, 5, 2), (128, 128, 32, 1), (256, 128, 16, 2), (32, 128, 24, 2), (96, 64, 40, 1)])
def test_modconv_layer_fwd_bwd(b2, cin, cout, res, up, tc, monkeypatch):
    """tc=False: exact-fp32 SIMT kernels; tc=True: tcgen05 kernels where the shape allows (3-pass split-bf16 forward and,
    for this test, 3-pass backward so that the tight tolerance applies to both)."""
    monkeypatch.setitem(b2.ops.CONFIG, 'tc', tc)
    monkeypatch.setitem(b2.ops.CONFIG, 'dgrad_passes', 3)
    monkeypatch.setitem(b2.ops.CONFIG, 'wgrad_passes', 3)
    n = 2
    g = gen(cin * 1000 + cout + up)
    x = torch.randn(n, cin, res, res, generator=g)
    W = torch.randn(cout, cin, 3, 3, generator=g)
    s = 1 + 0.3 * torch.randn(n, cin, generator=g)
    b = 0.2 * torch.randn(cout, generator=g)
    ores = res * up
    noise = torch.randn(ores, ores, generator=g)
    strength = torch.tensor(0.13)
    dz = torch.randn(n, cout, ores, ores, generator=g)
    gain, clamp = 1.0, 2.5       # small clamp so that the clamp branch is exercised

    leaves = [t.clone().requires_grad_(True) for t in (x, W, s, b, noise, strength)]
    xr, Wr, sr, br, nr, str_ = leaves
    y = oracle.modulated_conv2d(xr, Wr, sr, noise=nr * str_, up=up, f=oracle.fir_1331())
    z_ref = oracle.bias_act(y, br, act='lrelu', gain=math.sqrt(2) * gain, clamp=clamp * gain)
    z_ref.backward(dz)

    dev = 'cuda'
    cl = [t.detach().to(dev).requires_grad_(True) for t in (nhwc(x), W, s, b, noise, strength)]
    z, _ = b2.ops.modconv_layer(cl[0], cl[1], cl[2], cl[3], cl[4], cl[5], up, math.sqrt(2) * gain, clamp * gain)
    z.backward(nhwc(dz).to(dev))
    assert maxdiff(nchw(z), z_ref) < 2e-4 * max(1.0, z_ref.abs().max().item())
    names = ['dx', 'dW', 'dstyles', 'dbias', 'dnoise', 'dstrength']
    refs = [nhwc(xr.grad), Wr.grad, sr.grad, br.grad, nr.grad, str_.grad]
    for nm, a, r in zip(names, cl, refs):
        # scalar / mask-driven reductions move a little when split-bf16 rounding flips a clamp or lrelu mask
        tol = 2e-3 if (tc and nm in ('dstrength', 'dnoise', 'dbias')) else 2e-4
        assert relerr(a.grad, r) < tol, (nm, relerr(a.grad, r))


def test_modconv_layer_single_pass_backward(b2, monkeypatch):
    """Default PTI numerics: 3-pass forward, single-pass bf16 backward GEMMs -> gradients within 1e-2 relative L2."""
    monkeypatch.setitem(b2.ops.CONFIG, 'tc', True)
    monkeypatch.setitem(b2.ops.CONFIG, 'dgrad_passes', 1)
    monkeypatch.setitem(b2.ops.CONFIG, 'wgrad_passes', 1)
    n, cin, cout, res = 1, 128, 64, 32
    g = gen(77)
    x = torch.randn(n, cin, res, res, generator=g)
    W = torch.randn(cout, cin, 3, 3, generator=g)
    s = 1 + 0.3 * torch.randn(n, cin, generator=g)
    dz = torch.randn(n, cout, res, res, generator=g)
    for up in (1, 2):
        xr, Wr, sr = [t.clone().requires_grad_(True) for t in (x, W, s)]
        y = oracle.modulated_conv2d(xr, Wr, sr, up=up, f=oracle.fir_1331())
        z_ref = oracle.bias_act(y, None, act='lrelu')
        dzz = dz if up == 1 else torch.randn(n, cout, 2 * res, 2 * res, generator=g)
        z_ref.backward(dzz)
        cl = [t.detach().cuda().requires_grad_(True) for t in (nhwc(x), W, s)]
        z, _ = b2.ops.modconv_layer(cl[0], cl[1], cl[2], torch.zeros(cout, device='cuda'), None, None, up, math.sqrt(2), None)
        z.backward(nhwc(dzz).cuda())
        assert maxdiff(nchw(z), z_ref) < 2e-4 * z_ref.abs().max().item()
        for nm, a, r in zip(['dx', 'dW', 'ds'], cl, [nhwc(xr.grad), Wr.grad, sr.grad]):
            assert relerr(a.grad, r) < 1e-2, (up, nm, relerr(a.grad, r))


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('cin,cimg,res,prev', [(16, 96, 8, True), (64, 3, 16, True), (32, 96, 4, False), (10, 3, 6, True)])
def test_torgb_fwd_bwd(b2, cin, cimg, res, prev, tc, monkeypatch):
    monkeypatch.setitem(b2.ops.CONFIG, 'tc', tc)
    monkeypatch.setitem(b2.ops.CONFIG, 'dgrad_passes', 3)
    monkeypatch.setitem(b2.ops.CONFIG, 'wgrad_passes', 3)
    n = 2
    g = gen(cin + cimg)
    x = torch.randn(n, cin, res, res, generator=g)
    W = torch.randn(cimg, cin, 1, 1, generator=g)
    s = (1 + 0.3 * torch.randn(n, cin, generator=g)) / math.sqrt(cin)
    b = 0.2 * torch.randn(cimg, generator=g)
    ip = torch.randn(n, cimg, res // 2, res // 2, generator=g) if prev else None
    dimg = torch.randn(n, cimg, res, res, generator=g)
    clamp = 1.5
    leaves = [t.clone().requires_grad_(True) for t in (x, W, s, b)]
    ipr = ip.clone().requires_grad_(True) if prev else None
    y = oracle.bias_act(oracle.modulated_conv2d(leaves[0], leaves[1], leaves[2], demodulate=False), leaves[3], clamp=clamp)
    img_ref = oracle.upsample2d(ipr, oracle.fir_1331()) + y if prev else y
    img_ref.backward(dimg)

    dev = 'cuda'
    cl = [t.detach().to(dev).requires_grad_(True) for t in (nhwc(x), W, s, b)]
    ipc = nhwc(ip).to(dev).requires_grad_(True) if prev else None
    img = b2.ops.torgb_layer(cl[0], cl[1], cl[2], cl[3], ipc, clamp)
    img.backward(nhwc(dimg).to(dev))
    assert maxdiff(nchw(img), img_ref) < 1e-4
    for nm, a, r in zip(['dx', 'dW', 'ds', 'db'], cl, [nhwc(leaves[0].grad), leaves[1].grad, leaves[2].grad, leaves[3].grad]):
        assert relerr(a.grad, r) < 2e-4, (nm, relerr(a.grad, r))
    if prev:
        assert relerr(ipc.grad, nhwc(ipr.grad)) < 2e-4


# ---------------------------------------------------------------------------------------------- bias_act / upfirdn2d

@pytest.mark.parametrize('act', ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish'])
@pytest.mark.parametrize('clamp', [None, 0.7])
def test_bias_act_all_activations(b2, act, clamp):
    g = gen(3)
    x = torch.randn(2, 5, 7, 3, generator=g)
    b = torch.randn(5, generator=g)
    dy = torch.randn(2, 5, 7, 3, generator=g)
    fn = {'linear': lambda t: t, 'relu': torch.relu, 'lrelu': lambda t: F.leaky_relu(t, 0.2), 'tanh': torch.tanh,
          'sigmoid': torch.sigmoid, 'elu': F.elu, 'selu': F.selu, 'softplus': F.softplus, 'swish': lambda t: torch.sigmoid(t) * t}[act]
    gain = {'relu': math.sqrt(2), 'lrelu': math.sqrt(2), 'swish': math.sqrt(2)}.get(act, 1.0)
    xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = fn(xr + br.reshape(1, -1, 1, 1)) * gain
    if clamp is not None:
        y_ref = y_ref.clamp(-clamp, clamp)
    y_ref.backward(dy)
    xc, bc = x.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = b2.ops.bias_act(xc, bc, dim=1, act=act, clamp=clamp)
    y.backward(dy.cuda())
    assert maxdiff(y, y_ref) < 2e-6
    assert maxdiff(xc.grad, xr.grad) < 5e-6
    assert maxdiff(bc.grad, br.grad) < 5e-5


@pytest.mark.parametrize('up,down,pad,flip', [(1, 1, (1, 1, 1, 1), False), (2, 1, (2, 1, 2, 1), False), (1, 2, (1, 1, 1, 1), True),
                                              (2, 2, (3, 0, 1, 2), False), (1, 1, (-1, 2, 0, -1), False), (3, 1, (2, 2, 2, 2), False)])
def test_upfirdn2d_matches_oracle(b2, up, down, pad, flip):
    g = gen(up * 10 + down)
    x = torch.randn(2, 3, 9, 11, generator=g)
    f = torch.rand(4, 4, generator=g)
    f = f / f.sum()
    dy = None
    xr = x.clone().requires_grad_(True)
    fr = f.flip([0, 1]) if flip else f
    y_ref = oracle.upfirdn2d(xr, fr, up=up, down=down, pad=pad, gain=1.7)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    xc = x.cuda().requires_grad_(True)
    y = b2.ops.upfirdn2d(xc, f.cuda(), up=up, down=down, padding=list(pad), flip_filter=flip, gain=1.7)
    assert y.shape == y_ref.shape
    y.backward(dy.cuda())
    assert maxdiff(y, y_ref) < 1e-5
    assert maxdiff(xc.grad, xr.grad) < 1e-5


def test_upsample2d_and_setup_filter(b2):
    f = b2.ops.setup_filter([1, 3, 3, 1])
    assert maxdiff(f, oracle.fir_1331()) < 1e-7
    x = torch.randn(1, 4, 6, 6, generator=gen(0))
    assert maxdiff(b2.ops.upsample2d(x.cuda(), f.cuda()), oracle.upsample2d(x, oracle.fir_1331())) < 1e-5


def test_empty_and_error_paths(b2):
    from b200eg3d import _lib
    with pytest.raises(RuntimeError):
        b2.ops.bias_act(torch.zeros(2, 3), None)                 # CPU tensor: no CPU path exists
    with pytest.raises(RuntimeError):
        _lib.call('b200_conv_fwd', None, None, None, 1, 4, 4, 8, 8, 5, 1, None)     # bad kernel size
    y = b2.ops.bias_act(torch.zeros(0, 3, device='cuda'), torch.zeros(3, device='cuda'))
    assert y.numel() == 0


# ---------------------------------------------------------------------------------------------- renderer pieces

def _decoder(b2, seed):
    g = gen(seed)
    dec = b2.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    P = {}
    with torch.no_grad():
        for k, v in dec.named_parameters():
            v.copy_(torch.randn(v.shape, generator=g) * (0.3 if k.endswith('bias') else 1.0))
            P['decoder.' + k] = v.detach().clone()
    return dec, P


@pytest.fixture(params=[1, 0], ids=['tcgen05', 'mma_sync'])
def triplane_impl(request):
    """Both generations of the fused sampler + decoder stay under test: the tcgen05 pipeline (default) and the mma.sync kernels."""
    from b200eg3d import _lib
    prev = _lib.load().b200_set_triplane_impl(request.param)
    yield request.param
    _lib.load().b200_set_triplane_impl(prev)


@pytest.mark.parametrize('coord_grad', [True, False], ids=['dcoords', 'nodcoords'])
@pytest.mark.parametrize('npts', [1, 31, 32, 1000, 5000])
def test_run_model_fwd_bwd(b2, npts, triplane_impl, coord_grad):
    n, res = 2, 16
    g = gen(npts)
    planes = torch.randn(n, 3, 32, res, res, generator=g)
    coords = (torch.rand(n, npts, 3, generator=g) - 0.5) * 1.3       # some points fall outside the box (zero padding)
    dec, P = _decoder(b2, 5)
    d_rgb = torch.randn(n, npts, 32, generator=g)
    d_sig = torch.randn(n, npts, 1, generator=g)
    rk = {'box_warp': 1.2}
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    pr, cr = planes.clone().requires_grad_(True), coords.clone().requires_grad_(True)
    rgb_ref, sig_ref = oracle.run_model(Pr, pr, cr, rk)
    (rgb_ref * d_rgb).sum().add((sig_ref * d_sig).sum()).backward()

    dec = dec.cuda()
    for p in dec.parameters():
        p.requires_grad_(True)
    pl = planes.permute(0, 3, 4, 1, 2).reshape(n, res, res, 96).contiguous().cuda().requires_grad_(True)
    cc = coords.cuda().requires_grad_(coord_grad)        # without coordinate gradients the tcgen05 backward is taken (PTI shape)
    out = b2.ImportanceRenderer().run_model(pl, dec, cc, None, rk)
    (out['rgb'] * d_rgb.cuda()).sum().add((out['sigma'] * d_sig.cuda()).sum()).backward()
    assert maxdiff(out['rgb'], rgb_ref) < 2e-5
    assert maxdiff(out['sigma'], sig_ref) < 1e-4
    dpl_ref = pr.grad.permute(0, 3, 4, 1, 2).reshape(n, res, res, 96)
    # the gradient operands of the backward decoder GEMMs (d_out, d_a) are single bf16 / TF32 on the tensor cores (the forward and
    # the recompute keep the split operands): ~2^-9 relative PER POINT, averaging out over the points of a real pass -- with one or
    # two points nothing averages
    assert relerr(pl.grad, dpl_ref) < (3e-3 if npts >= 31 else 6e-3)
    if coord_grad:
        assert relerr(cc.grad, cr.grad) < (3e-3 if npts >= 31 else 6e-3)
    for k, v in dec.named_parameters():
        # dW1 / dW2 are contracted over the points from bf16 operands: rounding averages out over the ~1e6 points of a real
        # pass, with a handful of points it is ~4e-3
        assert relerr(v.grad, Pr['decoder.' + k].grad) < 1e-2, k


@pytest.mark.parametrize('S,S2,white', [(12, 12, False), (16, 0, False), (8, 8, True), (48, 48, False)])
@pytest.mark.parametrize('ray_grad', [True, False], ids=['drays', 'nodrays'])
def test_render_fwd_bwd(b2, S, S2, white, triplane_impl, ray_grad):
    n, res, R = 1, 32, 12
    M = R * R
    g = gen(S * 7 + S2)
    planes = torch.randn(n, 3, 32, res, res, generator=g)
    dec, P = _decoder(b2, 9)
    import synth_params as sp
    c = sp.camera(0.2, -0.1, n=n)
    ro, rd = oracle.ray_sampler(c[:, :16].reshape(-1, 4, 4), c[:, 16:].reshape(-1, 3, 3), R)
    ro, rd = ro.contiguous(), rd.contiguous()
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S2, white_back=white)
    u1 = torch.rand(n, M, S, 1, generator=g)
    u2 = torch.rand(n * M, max(S2, 1), generator=g)
    dfeat = torch.randn(n, M, 32, generator=g)
    ddepth = torch.randn(n, M, 1, generator=g)
    dws = torch.randn(n, M, 1, generator=g)

    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    pr = planes.clone().requires_grad_(True)
    ror, rdr = ro.clone().requires_grad_(True), rd.clone().requires_grad_(True)
    f_ref, d_ref, w_ref = oracle.render(Pr, pr, ror, rdr, rk, u1, u2)
    ((f_ref * dfeat).sum() + (d_ref * ddepth).sum() + (w_ref * dws).sum()).backward()

    dec = dec.cuda()
    for p in dec.parameters():
        p.requires_grad_(True)
    ren = b2.ImportanceRenderer()
    ren.fixed_noise = (u1, u2)
    pl = planes.permute(0, 3, 4, 1, 2).reshape(n, res, res, 96).contiguous().cuda().requires_grad_(True)
    roc, rdc = ro.cuda().requires_grad_(ray_grad), rd.cuda().requires_grad_(ray_grad)
    f, d, w = ren(pl, dec, roc, rdc, rk)
    ((f * dfeat.cuda()).sum() + (d * ddepth.cuda()).sum() + (w * dws.cuda()).sum()).backward()
    assert maxdiff(f, f_ref) < 5e-5
    assert maxdiff(d, d_ref) < 5e-5
    assert maxdiff(w, w_ref) < 5e-5
    assert relerr(pl.grad, pr.grad.permute(0, 3, 4, 1, 2).reshape(n, res, res, 96)) < 5e-3
    if ray_grad:
        assert relerr(roc.grad, ror.grad) < 1e-2
        assert relerr(rdc.grad, rdr.grad) < 1e-2
    for k, v in dec.named_parameters():
        assert relerr(v.grad, Pr['decoder.' + k].grad) < 1e-2, k


@pytest.mark.gpu
@pytest.mark.parametrize('n,h,r', [(1, 512, 128), (2, 16, 8), (1, 12, 12), (3, 24, 4)])
def test_pti_loss_fwd_bwd(b2, n, h, r):
    """Fused calc_loss (base_coach.py:101-126 without LPIPS) against the oracle, on the layouts the generator produces:
    image = NCHW view of NHWC memory, image_raw = 3-channel slice of the NHWC feature image, depth contiguous."""
    g = torch.Generator().manual_seed(n * 1000 + h)
    img_nhwc = torch.randn(n, h, h, 3, generator=g)
    feat_nhwc = torch.randn(n, r, r, 32, generator=g)
    depth = torch.rand(n, 1, r, r, generator=g) + 2.5
    real = torch.rand(n, 3, h, h, generator=g) * 2 - 1
    ref_in = [t.clone().requires_grad_(True) for t in (img_nhwc, feat_nhwc, depth)]
    out_ref = {'image': ref_in[0].permute(0, 3, 1, 2), 'image_raw': ref_in[1].permute(0, 3, 1, 2)[:, :3], 'image_depth': ref_in[2]}
    loss_ref, parts_ref = oracle.calc_loss(out_ref, real, 0.7, 1.3)
    (loss_ref * 1.5).backward()
    dev_in = [t.clone().cuda().requires_grad_(True) for t in (img_nhwc, feat_nhwc, depth)]
    out = {'image': dev_in[0].permute(0, 3, 1, 2), 'image_raw': dev_in[1].permute(0, 3, 1, 2)[:, :3], 'image_depth': dev_in[2]}
    loss, parts = b2.losses.pti_loss(out, real.cuda(), 0.7, 1.3, return_parts=True)
    (loss * 1.5).backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * max(1.0, abs(loss_ref.item()))
    for a, b in zip(parts[1:].tolist(), parts_ref):
        assert abs(a - b.item()) <= 1e-5 * max(1.0, abs(b.item()))
    for a, b in zip(dev_in, ref_in):
        assert relerr(a.grad, b.grad) < 1e-5
    tv = b2.losses.compute_tv_norm(depth.cuda())
    assert abs(tv.item() - oracle.tv_norm(depth).item()) <= 1e-6 + 1e-5 * oracle.tv_norm(depth).item()
    with pytest.raises(RuntimeError, match='divide'):
        b2.losses.pti_loss({'image': dev_in[0].permute(0, 3, 1, 2), 'image_raw': torch.zeros(n, 3, 5, 5, device='cuda'), 'image_depth': None},
                           real.cuda())


@pytest.mark.parametrize('mode', ['sorted', 'unsorted', 'ties'])
@pytest.mark.parametrize('S1,S2', [(48, 48), (7, 13), (16, 0), (96, 96)])
def test_ray_composite_merge_orders(b2, mode, S1, S2):
    """The merge of coarse + fine samples (renderer.py:212-222): sorted coarse depths take the binary-search merge, unsorted
    ones the all-pairs ranking; ties resolve like a stable sort of [coarse..., fine...].  Forward and backward vs the oracle."""
    from b200eg3d._lib import call, ptr, stream
    M = 40
    g = gen(S1 * 100 + S2 + len(mode))
    t_c = torch.rand(1, M, S1, 1, generator=g) + 2.0
    if mode != 'unsorted':
        t_c = t_c.sort(dim=2).values
    t_f = torch.rand(1, M, max(S2, 1), 1, generator=g)[:, :, :S2] + 2.0
    if mode == 'ties' and S2 > 0:
        k = min(S1, S2)
        t_f[:, :, :k:3] = t_c[:, :, :k:3]                                            # fine samples that coincide with coarse ones
        if S2 > 4:
            t_f[:, :, 1] = t_f[:, :, 4]                                              # and a tie inside the fine list
    rgb = torch.rand(1, M, S1 + S2, 32, generator=g)
    sig = torch.randn(1, M, S1 + S2, 1, generator=g) * 3
    d_feat, d_depth = torch.randn(1, M, 32, generator=g), torch.randn(1, M, 1, generator=g)
    # oracle: stable sort of the concatenation, then the mid-point marcher
    rr, sr = rgb.clone().requires_grad_(True), sig.clone().requires_grad_(True)
    t_all = torch.cat([t_c, t_f], 2)
    order = torch.sort(t_all, dim=2, stable=True).indices
    feat_ref, depth_ref, w_ref = oracle.ray_march(torch.gather(rr, 2, order.expand(-1, -1, -1, 32)), torch.gather(sr, 2, order),
                                                  torch.gather(t_all, 2, order))
    ((feat_ref * d_feat).sum() + (depth_ref * d_depth).sum()).backward()
    dev = lambda t: t.contiguous().cuda()
    tc, tf = dev(t_c.reshape(M, S1)), (dev(t_f.reshape(M, S2)) if S2 else None)
    rc, rf = dev(rgb[0, :, :S1]), (dev(rgb[0, :, S1:]) if S2 else None)
    sc, sf = dev(sig[0, :, :S1, 0]), (dev(sig[0, :, S1:, 0]) if S2 else None)
    mm = torch.zeros(2, dtype=torch.int32, device='cuda'); mm[:1].fill_(-1)
    call('b200_depth_minmax', ptr(tc), tc.numel(), ptr(mm), stream())
    if S2:
        call('b200_depth_minmax', ptr(tf), tf.numel(), ptr(mm), stream())
    feat, depth, wsum = torch.empty(M, 32, device='cuda'), torch.empty(M, device='cuda'), torch.empty(M, device='cuda')
    call('b200_ray_composite_fwd', ptr(tc), ptr(sc), ptr(rc), S1, ptr(tf), ptr(sf), ptr(rf), S2, ptr(mm), 0, M, ptr(feat), ptr(depth), ptr(wsum), stream())
    assert maxdiff(feat, feat_ref[0]) < 2e-5 and maxdiff(depth, depth_ref[0, :, 0]) < 2e-5
    assert maxdiff(wsum, w_ref.sum(2)[0, :, 0]) < 2e-5
    d_rc, d_sc = torch.empty_like(rc), torch.empty_like(sc)
    d_rf, d_sf = (torch.empty_like(rf), torch.empty_like(sf)) if S2 else (None, None)
    g_feat, g_depth = dev(d_feat[0]), dev(d_depth[0, :, 0])        # keep the device copies alive across the asynchronous launch
    call('b200_ray_composite_bwd', ptr(tc), ptr(sc), ptr(rc), S1, ptr(tf), ptr(sf), ptr(rf), S2, ptr(mm), 0, M, ptr(g_feat),
         ptr(g_depth), None, ptr(d_rc), ptr(d_sc), ptr(d_rf), ptr(d_sf), stream())
    assert relerr(d_rc, rr.grad[0, :, :S1]) < 1e-4 and relerr(d_sc, sr.grad[0, :, :S1, 0]) < 1e-4
    if S2:
        assert relerr(d_rf, rr.grad[0, :, S1:]) < 1e-4 and relerr(d_sf, sr.grad[0, :, S1:, 0]) < 1e-4


@pytest.mark.parametrize('n,R', [(1, 128), (2, 16), (3, 5)])
def test_ray_sampler_kernel(b2, n, R):
    """RaySampler.forward kernel + its cam2world gradient against the oracle (ray_sampler.py:24-73)."""
    import synth_params as sp
    c = torch.cat([sp.camera(0.3 - 0.2 * i, -0.2 + 0.15 * i) for i in range(n)], 0)
    c[:, 17] = 0.02                                                   # non-zero skew exercises the full unprojection
    g = gen(R)
    u_o, u_d = torch.randn(n, R * R, 3, generator=g), torch.randn(n, R * R, 3, generator=g)
    cr = c[:, :16].reshape(n, 4, 4).clone().requires_grad_(True)
    o_ref, d_ref = oracle.ray_sampler(cr, c[:, 16:].reshape(n, 3, 3), R)
    ((o_ref * u_o).sum() + (d_ref * u_d).sum()).backward()
    cd = c[:, :16].reshape(n, 4, 4).clone().cuda().requires_grad_(True)
    o, d = b2.RaySampler()(cd, c[:, 16:].reshape(n, 3, 3).cuda(), R)
    ((o * u_o.cuda()).sum() + (d * u_d.cuda()).sum()).backward()
    assert maxdiff(o, o_ref) == 0.0 and maxdiff(d, d_ref) < 1e-6
    assert relerr(cd.grad[:, :3], cr.grad[:, :3]) < 1e-4 and float(cd.grad[:, 3].abs().max()) == 0.0
    o2, d2 = b2.RaySampler().forward_torch(cd.detach(), c[:, 16:].reshape(n, 3, 3).cuda(), R)
    assert maxdiff(o2, o) == 0.0 and maxdiff(d2, d) < 1e-6
    _, dcam, uv = b2.RaySampler()(cd.detach(), c[:, 16:].reshape(n, 3, 3).cuda(), R, need_cam_space=True)
    assert dcam.shape == (n, R * R, 3) and uv.shape == (n, R * R, 2)


@pytest.mark.parametrize('noise_mode', ['const', 'batch', 'none'])
def test_conv_fwd_tc_fused_epilogue_matches_two_kernel_path(b2, noise_mode):
    """b200_conv_fwd_tc_act (layer epilogue applied in the TMEM drain) == b200_conv_fwd_tc followed by b200_layer_act_fwd."""
    from b200eg3d._lib import call, ptr, stream, load
    n, h, w, cin, cout = 2, 96, 128, 256, 64
    assert load().b200_conv_tc_act_fusable(n, h, w, cin, cout, 3) == 1
    assert load().b200_conv_tc_act_fusable(n, 8, 8, cin, cout, 3) == 0          # split-K layer: not fusable
    g = gen(17)
    x = torch.randn(n, h, w, cin, generator=g).cuda()
    wm = (torch.randn(n, 9, cout, cin, generator=g) * 0.05).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    strength = torch.full([], 0.3).cuda()
    noise, nbs = None, 0
    if noise_mode == 'const':
        noise = torch.randn(h, w, generator=g).cuda()
    elif noise_mode == 'batch':
        noise, nbs = torch.randn(n, 1, h, w, generator=g).cuda(), h * w
    xh, xl = b2.ops._split(x, True)
    wh, wl = b2.ops._split(wm, True)
    y = torch.empty(n, h, w, cout, device='cuda')
    z0, z1 = torch.empty_like(y), torch.empty_like(y)
    h0, l0, h1, l1 = [torch.empty(n, h, w, cout, device='cuda', dtype=torch.bfloat16) for _ in range(4)]
    call('b200_conv_fwd_tc', ptr(xh), ptr(xl), ptr(wh), ptr(wl), ptr(y), n, h, w, cin, cout, 3, 1, 3, 0, stream())
    call('b200_layer_act_fwd', ptr(y), ptr(z0), ptr(h0), ptr(l0), ptr(bias), ptr(noise), ptr(strength) if noise is not None else None, nbs,
         n, h * w, cout, 1, 0.2, math.sqrt(2), 1.5, stream())
    call('b200_conv_fwd_tc_act', ptr(xh), ptr(xl), ptr(wh), ptr(wl), ptr(z1), ptr(h1), ptr(l1), ptr(bias), ptr(noise),
         ptr(strength) if noise is not None else None, nbs, n, h, w, cin, cout, 3, 3, 0.2, math.sqrt(2), 1.5, stream())
    assert float(z0.abs().max()) == 1.5                                          # the clamp is active
    assert maxdiff(z1, z0) < 1e-6 and maxdiff(h1.float(), h0.float()) < 1e-2 and maxdiff((h1.float() + l1.float()), z0) < 1e-4


# ---------------------------------------------------------------------------------------------- optimiser

@pytest.mark.parametrize('lr_on_device', [False, True])
def test_adam_matches_torch_adam(b2, lr_on_device):
    """b200eg3d.optim.Adam (one b200_adam_step launch for the whole list) against torch.optim.Adam: ragged sizes (vector path,
    scalar tail, unaligned views), a parameter without gradient, five steps with changing gradients."""
    from b200eg3d.optim import Adam
    g = gen(5)
    base = torch.randn(70001, generator=g).cuda()
    shapes = [(512, 512, 3, 3), (3,), (17, 5), (1,), (8192,), (8193,)]
    pa = [torch.randn(*s, generator=g).cuda().requires_grad_(True) for s in shapes]
    pa.append(base[1:40000].detach().clone().requires_grad_(True))
    pa.append(torch.randn(64, generator=g).cuda().requires_grad_(True))            # never receives a gradient
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    lr = 3e-3
    oa = Adam(pa, lr=torch.tensor(lr, device='cuda') if lr_on_device else lr, betas=(0.9, 0.999), eps=1e-8)
    ob = torch.optim.Adam(pb, lr=lr, betas=(0.9, 0.999), eps=1e-8)
    for it in range(5):
        for a, b in zip(pa[:-1], pb[:-1]):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (it - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    assert float(oa.step_count) == 5.0
    for a, b in zip(pa, pb):
        assert relerr(a, b) < 2e-6, (a.shape, relerr(a, b))
    assert maxdiff(pa[-1], pb[-1]) == 0.0
    sd = oa.state_dict()
    oa.load_state_dict(sd)
    assert float(oa.state[pa[0]]['step']) == 5.0


def test_wgrad_pool_and_stream_match_serial_path(b2, monkeypatch, golden_dir):
    """The weight gradients of a synthesis network are accumulated into one zero-filled pool by wgrad kernels running on their own
    stream (ops.CONFIG['wgrad_stream']); with the stream off the same kernels run in line.  Both must give the same gradients."""
    import synth_params as sp
    from golden_util import load_case
    case = load_case(golden_dir, 'tiny_r64_s16')
    G = b2.TriPlaneGenerator(rendering_kwargs=case.rk, **case.gk).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), case.param_seed)
    G = G.cuda().float()
    G.neural_rendering_resolution = case.R
    G.renderer.fixed_noise = (case.u_strat.cuda(), case.u_imp.cuda())
    ws, c = case.ws.cuda(), case.c.cuda()
    named = [(n, p) for n, p in G.named_parameters() if '.mapping.' not in n]
    grads = []
    for on in (True, False, True):
        monkeypatch.setitem(b2.ops.CONFIG, 'wgrad_stream', on)
        for _, p in named:
            p.grad = None
        out = G.synthesis(ws, c, noise_mode='const')
        (out['image'].square().mean() + out['image_raw'].square().mean()).backward()
        torch.cuda.synchronize()
        grads.append([p.grad.clone() if p.grad is not None else None for _, p in named])
    assert sum(g is not None and float(g.abs().max()) > 0 for g in grads[0]) > 20
    for (n, _), a, b, a2 in zip(named, *grads):
        if a is None:
            assert b is None and a2 is None, n
            continue
        if a.ndim < 2:
            continue        # scalar / vector gradients (noise strengths, biases) are cancellation-dominated: the atomics' order moves them
        # run-to-run spread of the weight gradients (atomic order in the tri-plane scatter) is ~1e-4 .. 3e-3 in either mode
        assert relerr(a, b) < 1e-2 and relerr(a2, b) < 1e-2, (n, relerr(a, b), relerr(a2, b))


# ---------------------------------------------------------------------------------------------- CTA-pair convolution tiles

@pytest.mark.parametrize('kind,n,h,w,cin,cout,k,up', [
    ('fwd', 1, 128, 128, 128, 128, 3, 1),      # N = 128 pair tiles
    ('fwd', 1, 64, 64, 512, 512, 3, 1),        # b64.conv1: 16 pair tiles x 4 N tiles
    ('fwd', 1, 128, 128, 256, 256, 3, 1),      # N = 256 pair tiles (all 512 tensor-memory columns)
    ('fwd', 1, 120, 136, 64, 64, 3, 1),        # odd tile count (dummy second tile), N = 64
    ('fwd', 2, 96, 96, 128, 96, 1, 1),         # 1x1 (ToRGB shape), ragged N, batch 2
    ('fwd', 1, 64, 64, 512, 256, 3, 2),        # transposed conv: four output-parity classes, 65 x 65 grids
    ('fwd', 1, 128, 128, 128, 64, 3, 2),
    ('dgrad', 1, 128, 128, 128, 128, 3, 1),
    ('dgrad', 1, 128, 128, 256, 256, 3, 1),    # MN-major B, N = 256
    ('dgrad', 1, 64, 64, 512, 512, 3, 1),
    ('dgrad', 1, 72, 72, 256, 128, 3, 2),      # stride-2 gather
    ('dgrad', 2, 100, 60, 192, 96, 1, 1),      # ragged N tile (192 = 128 + 64), batch 2
])
def test_conv_tc_pair_tiles_match_single_cta(b2, kind, n, h, w, cin, cout, k, up):
    """cta_group::2 (M = 256 over two pixel tiles, B shared by the pair) against the single-CTA tcgen05 kernel on the same operands,
    and against an fp64 ATen convolution."""
    from b200eg3d._lib import call, ptr, stream, load
    lib = load()
    g = gen(h * 7 + cin + cout + up)
    taps = k * k
    wm = (torch.randn(n, taps, cout, cin, generator=g) / math.sqrt(cin * taps)).cuda()
    wh, wl = b2.ops._split(wm, True)
    hs, ws_ = (h, w) if up == 1 else (2 * h + 1, 2 * w + 1)
    if kind == 'fwd':
        x = torch.randn(n, h, w, cin, generator=g).cuda()
        xh, xl = b2.ops._split(x, True)
        outs = []
        for pair in (1, 0):
            prev = lib.b200_set_conv_pair(pair)
            y = torch.full([n, hs, ws_, cout], float('nan'), device='cuda')
            call('b200_conv_fwd_tc', ptr(xh), ptr(xl), ptr(wh), ptr(wl), ptr(y), n, h, w, cin, cout, k, up, 3, 0, stream())
            lib.b200_set_conv_pair(prev)
            outs.append(y)
        wt = wm.double().reshape(n, k, k, cout, cin)
        ref = []
        for i in range(n):
            xi = x[i:i + 1].double().permute(0, 3, 1, 2)
            if up == 1:
                ref.append(F.conv2d(xi, wt[i].permute(2, 3, 0, 1), padding=k // 2))
            else:      # the (2h+1) x (2w+1) grid of the stride-2 transposed convolution: out[2i + kh, 2j + kw] += x[i, j] * W[kh, kw]
                ref.append(F.conv_transpose2d(xi, wt[i].permute(3, 2, 0, 1), stride=2))
        ref = torch.cat(ref).permute(0, 2, 3, 1)
    else:
        dy = torch.randn(n, hs, ws_, cout, generator=g).cuda()
        dh, dl = b2.ops._split(dy, True)
        outs = []
        for pair in (1, 0):
            prev = lib.b200_set_conv_pair(pair)
            dx = torch.full([n, h, w, cin], float('nan'), device='cuda')
            call('b200_conv_dgrad_tc', ptr(dh), ptr(dl), ptr(wh), ptr(wl), ptr(dx), n, h, w, cin, cout, k, up, 3, 0, stream())
            lib.b200_set_conv_pair(prev)
            outs.append(dx)
        ref = None
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all() and torch.isfinite(outs[1]).all()
    assert maxdiff(outs[0], outs[1]) <= 1e-6 * float(outs[1].abs().max()), maxdiff(outs[0], outs[1])
    if ref is not None:
        assert relerr(outs[0], ref) < 2e-5, relerr(outs[0], ref)


# ---------------------------------------------------------------------------------------------- lean activations (no fp32 copy)

@pytest.mark.parametrize('c', [64, 128, 256, 20])
def test_layer_act_bwd_from_split_output_matches_fp32_output(b2, c):
    """b200_layer_act_bwd with the saved output given as its split-bf16 pair (hi decides the sign, hi + lo the clamp test, read only
    where hi reaches the clamp) == the same call with the fp32 output, including values AT the clamp and just below it."""
    from b200eg3d._lib import call, ptr, stream
    g = gen(c)
    n, hw, clamp, gain = 2, 37 * 23, 2.0, math.sqrt(2)
    z = (torch.randn(n, hw, c, generator=g) * 1.5).clamp(-clamp, clamp)
    flat = z.view(-1)
    flat[::7] = clamp                                   # clamped from above
    flat[1::11] = -clamp
    flat[2::13] = clamp - 1e-3                          # rounds UP to the clamp in bf16, but is not clamped
    flat[3::17] = -(clamp - 3e-4)
    flat[4::19] = 0.0
    z = z.cuda()
    dz = torch.randn(n, hw, c, generator=g).cuda()
    noise = torch.randn(hw, generator=g).cuda()
    strength = torch.full([], 0.4).cuda()
    zh, zl = b2.ops._split(z, True) if c % 4 == 0 else (z.to(torch.bfloat16), (z - z.to(torch.bfloat16).float()).to(torch.bfloat16))
    res = []
    for zr in ((ptr(z), None, None), (None, ptr(zh), ptr(zl))):
        dy = torch.empty_like(dz)
        dbias, dstr, dnoise = torch.zeros(c, device='cuda'), torch.zeros([], device='cuda'), torch.zeros_like(noise)
        call('b200_layer_act_bwd', ptr(dz), *zr, ptr(dy), None, None, ptr(dbias), ptr(noise), ptr(strength), 0, ptr(dstr), ptr(dnoise),
             n, hw, c, 1, 0.2, gain, clamp, stream())
        res.append((dy, dbias, dstr, dnoise))
    assert maxdiff(res[0][0], res[1][0]) == 0.0
    for a, b in zip(res[0][1:], res[1][1:]):
        assert relerr(a, b) < 1e-5
    ref = dz * gain * torch.where(z > 0, 1.0, 0.2) * (z.abs() < clamp)
    assert maxdiff(res[1][0], ref) < 1e-6


def test_conv1x1_wgrad_from_split_input(b2):
    from b200eg3d._lib import call, ptr, stream
    g = gen(3)
    n, npix, cin, cout = 2, 4099, 64, 3
    x = torch.randn(n, npix, cin, generator=g).cuda()
    dy = torch.randn(n, npix, cout, generator=g).cuda()
    xh, xl = b2.ops._split(x, True)
    dw = torch.empty(n, 1, cout, cin, device='cuda')
    call('b200_conv1x1_wgrad_split', ptr(xh), ptr(xl), ptr(dy), ptr(dw), n, npix, cin, cout, stream())
    ref = torch.einsum('npo,npi->noi', dy.double(), (xh.double() + xl.double()))
    assert relerr(dw[:, 0], ref) < 1e-5
    assert relerr(dw[:, 0], torch.einsum('npo,npi->noi', dy.double(), x.double())) < 1e-4


def test_lean_activations_match_fp32_copies(b2, monkeypatch, golden_dir):
    """ops.CONFIG['lean_acts']: activations between tensor-core layers exist only as their split-bf16 pair (the layer's fp32 output is a
    shape-only placeholder).  Image, raw image and every gradient must equal the path that keeps the fp32 copies."""
    import synth_params as sp
    from golden_util import load_case
    case = load_case(golden_dir, 'full_r64_s16')
    G = b2.TriPlaneGenerator(rendering_kwargs=case.rk, **case.gk).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), case.param_seed)
    G = G.cuda().float()
    G.neural_rendering_resolution = case.R
    G.renderer.fixed_noise = (case.u_strat.cuda(), case.u_imp.cuda())
    ws, c = case.ws.cuda(), case.c.cuda()
    named = [(n, p) for n, p in G.named_parameters() if '.mapping.' not in n]
    outs = []
    for lean in (True, False):
        monkeypatch.setitem(b2.ops.CONFIG, 'lean_acts', lean)
        for _, p in named:
            p.grad = None
        out = G.synthesis(ws, c, noise_mode='const')
        (out['image'].square().mean() + out['image_raw'].square().mean()).backward()
        torch.cuda.synchronize()
        outs.append((out['image'].detach().clone(), out['image_raw'].detach().clone(), [p.grad.clone() if p.grad is not None else None for _, p in named]))
    # two forward passes differ by the order of the split-K atomics of the 4x4 .. 32x32 blocks (~1e-7), which the importance
    # sampling amplifies: a few 1e-5 between ANY two runs
    assert maxdiff(outs[0][0], outs[1][0]) < 2e-4 and maxdiff(outs[0][1], outs[1][1]) < 2e-4
    for (n, _), a, b in zip(named, outs[0][2], outs[1][2]):
        if a is None or a.ndim < 2:
            continue        # scalar / vector gradients are cancellation-dominated (atomic order), see test_wgrad_pool_and_stream_match_serial_path
        assert relerr(a, b) < 1e-2, (n, relerr(a, b))


@pytest.mark.parametrize('c,split_out', [(64, True), (128, True), (256, True), (512, True), (128, False)])
def test_layer_act_bwd_two_addends_equal_presummed_gradient(b2, c, split_out):
    """b200_layer_act_bwd_sum2(dz, dz2) == b200_layer_act_bwd(dz + dz2), bit for bit in dy (the kernel forms the same fp32 sum autograd's
    accumulation pass would), for the split and the fp32 form of the saved output; shapes outside the specialised kernel report 0."""
    from b200eg3d._lib import call, ptr, stream
    lib = b2._lib.load()
    g = gen(c + 5)
    n, hw, clamp, gain = 2, 29 * 31, 2.0, math.sqrt(2)
    z = (torch.randn(n, hw, c, generator=g) * 1.5).clamp(-clamp, clamp).cuda()
    dz, dz2 = torch.randn(n, hw, c, generator=g).cuda(), torch.randn(n, hw, c, generator=g).cuda()
    noise = torch.randn(hw, generator=g).cuda()
    strength = torch.full([], 0.4).cuda()
    zh, zl = b2.ops._split(z, True)
    assert lib.b200_layer_act_bwd_sum2_supported(n, hw, c, 1, 0) == 1
    assert lib.b200_layer_act_bwd_sum2_supported(n, hw, 20, 1, 0) == 0 and lib.b200_layer_act_bwd_sum2_supported(n, hw, 20, 0, 0) == 1   # 5 lanes: no shuffle sum
    assert lib.b200_layer_act_bwd_sum2_supported(n, hw, 3, 0, 0) == 0
    for zr in ((ptr(z), None, None), (None, ptr(zh), ptr(zl))):
        res = []
        for two in (True, False):
            dy = torch.empty_like(dz) if not split_out else None
            dyh = torch.empty(n, hw, c, device='cuda', dtype=torch.bfloat16) if split_out else None
            dyl = torch.empty_like(dyh) if split_out else None
            dbias, dstr, dnoise = torch.zeros(c, device='cuda'), torch.zeros([], device='cuda'), torch.zeros_like(noise)
            tail = (ptr(dy), ptr(dyh), ptr(dyl), ptr(dbias), ptr(noise), ptr(strength), 0, ptr(dstr), ptr(dnoise), n, hw, c, 1, 0.2, gain, clamp,
                    stream())
            if two:
                call('b200_layer_act_bwd_sum2', ptr(dz), ptr(dz2), *zr, *tail)
            else:
                dsum = dz + dz2
                call('b200_layer_act_bwd', ptr(dsum), *zr, *tail)
            res.append((dy if not split_out else dyh.float() + dyl.float(), dbias, dstr, dnoise))
        torch.cuda.synchronize()
        assert maxdiff(res[0][0], res[1][0]) == 0.0
        for a, b in zip(res[0][1:], res[1][1:]):
            assert relerr(a, b) < 1e-5
        ref = (dz + dz2) * gain * torch.where(z > 0, 1.0, 0.2) * (z.abs() < clamp)
        assert maxdiff(res[0][0], ref) < (1e-6 if not split_out else 2e-4)


@pytest.mark.parametrize('key', ['fork_grads', 'early_sr_bank'])
def test_forked_gradients_match_autograd_accumulation(b2, monkeypatch, golden_dir, key):
    """ops.CONFIG['fork_grads']: conv1's output reaches the next block and the block's ToRGB through two handles, and the two gradients
    are summed inside the activation backward.  Every gradient must equal the path where autograd adds them in a pass of its own.
    ops.CONFIG['early_sr_bank']: the super-resolution module's styles / modulated weights (and their backward) on the second stream
    beside the renderer must equal the stream-ordered placement in front of the first SR convolution."""
    import synth_params as sp
    from golden_util import load_case
    case = load_case(golden_dir, 'full_r64_s16')
    G = b2.TriPlaneGenerator(rendering_kwargs=case.rk, **case.gk).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), case.param_seed)
    G = G.cuda().float()
    G.neural_rendering_resolution = case.R
    G.renderer.fixed_noise = (case.u_strat.cuda(), case.u_imp.cuda())
    c = case.c.cuda()
    named = [(n, p) for n, p in G.named_parameters() if '.mapping.' not in n]
    outs = []
    for fork in (True, False):
        monkeypatch.setitem(b2.ops.CONFIG, key, fork)
        for _, p in named:
            p.grad = None
        ws = case.ws.cuda().requires_grad_(True)
        out = G.synthesis(ws, c, noise_mode='const')
        (out['image'].square().mean() + out['image_raw'].square().mean()).backward()
        torch.cuda.synchronize()
        outs.append((out['image'].detach().clone(), ws.grad.clone(), [p.grad.clone() if p.grad is not None else None for _, p in named]))
    assert maxdiff(outs[0][0], outs[1][0]) < 2e-4
    assert relerr(outs[0][1], outs[1][1]) < 1e-3, relerr(outs[0][1], outs[1][1])
    for (n, _), a, b in zip(named, outs[0][2], outs[1][2]):
        if a is None or a.ndim < 2:
            continue
        assert relerr(a, b) < 1e-2, (n, relerr(a, b))


@pytest.mark.parametrize('h,w,c,pad,flip,act', [(513, 513, 64, 1, 0, True), (257, 257, 128, 1, 0, True), (256, 256, 128, 2, 1, False),
                                                (130, 203, 64, 2, 1, False), (131, 201, 128, 1, 0, True)])
def test_fir_column_window_matches_patch_kernel(b2, h, w, c, pad, flip, act):
    """b200_upfirdn2d_fused with the separable hint (sliding column window, 8 MACs per element) == the 2 x 4 patch kernel and the
    oracle's upfirdn2d on the same input: forward shape (pad 1, layer epilogue, split outputs) and adjoint shape (pad 2, flipped)."""
    from b200eg3d._lib import call, ptr, stream
    g = gen(h + w + c)
    x = torch.randn(1, h, w, c, generator=g).cuda()
    f = b2.ops.fir_filter(torch.device('cuda'))
    oh, ow = h + 2 * pad - 3, w + 2 * pad - 3
    bias = torch.randn(c, generator=g).cuda()
    noise = torch.randn(oh, ow, generator=g).cuda()
    strength = torch.full([], 0.3).cuda()
    outs = []
    for sep in (1, 0):
        y = torch.full([1, oh, ow, c], float('nan'), device='cuda')
        yh = torch.empty(1, oh, ow, c, device='cuda', dtype=torch.bfloat16)
        yl = torch.empty_like(yh)
        a = (1, ptr(bias), ptr(noise), ptr(strength), 0, 1, 0.2, math.sqrt(2), 1.5) if act else (0, None, None, None, 0, 0, 0.0, 1.0, -1.0)
        call('b200_upfirdn2d_fused', ptr(x), ptr(f), None, ptr(y), ptr(yh), ptr(yl), 1, h, w, c, 4, 4, 1, 1, pad, pad, pad, pad, flip, 4.0,
             *a, sep, stream())
        outs.append((y, yh.float() + yl.float()))
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0][0]).all()
    assert maxdiff(outs[0][0], outs[1][0]) < 2e-5 and maxdiff(outs[0][1], outs[1][1]) < 2e-4
    if not act:
        ff = f.cpu().flip([0, 1]) if flip else f.cpu()          # the filter is symmetric: flipping is exercised, not distinguished
        ref = oracle.upfirdn2d(nchw(x).cpu(), ff, pad=(pad, pad, pad, pad), gain=4.0)
        assert maxdiff(nchw(outs[0][0]), ref) < 2e-5


def test_zero_pools_match_per_launch_memsets(b2, monkeypatch, golden_dir):
    """ops.CONFIG['zero_pools']: the outputs of all split-K convolutions of a network come from one zero-filled pool (the launches are told
    `prezeroed`); switched off, every such launch clears its own output.  Same image and gradients either way."""
    import synth_params as sp
    from golden_util import load_case
    from b200eg3d._lib import load
    assert load().b200_conv_tc_ksplit(0, 1, 16, 16, 512, 512, 3, 1) > 1
    case = load_case(golden_dir, 'full_r64_s16')
    G = b2.TriPlaneGenerator(rendering_kwargs=case.rk, **case.gk).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), case.param_seed)
    G = G.cuda().float()
    G.neural_rendering_resolution = case.R
    G.renderer.fixed_noise = (case.u_strat.cuda(), case.u_imp.cuda())
    ws, c = case.ws.cuda(), case.c.cuda()
    named = [(n, p) for n, p in G.named_parameters() if '.mapping.' not in n]
    outs = []
    for on in (True, False):
        monkeypatch.setitem(b2.ops.CONFIG, 'zero_pools', on)
        for _, p in named:
            p.grad = None
        out = G.synthesis(ws, c, noise_mode='const')
        (out['image'].square().mean() + out['image_raw'].square().mean()).backward()
        torch.cuda.synchronize()
        outs.append((out['image'].detach().clone(), [p.grad.clone() if p.grad is not None else None for _, p in named]))
    assert maxdiff(outs[0][0], outs[1][0]) < 2e-4
    for (n, _), a, b in zip(named, outs[0][1], outs[1][1]):
        if a is None or a.ndim < 2:
            continue
        assert relerr(a, b) < 1e-2, (n, relerr(a, b))


@pytest.mark.parametrize('M,S', [(1000, 48), (37, 7), (4096, 96)])
def test_depths_coarse_range_and_fine_depths_inside(b2, M, S):
    """b200_ray_depths_coarse also gathers the global depth range (the two order-preserving words of b200_depth_minmax) -- and the
    fine depths of b200_ray_importance are interpolations between mid-points of their ray's coarse depths (renderer.py:297-307),
    so they never leave that ray's coarse range: the range of the merged samples that ray_marcher.py:50 clamps to IS the coarse
    range, which is why ops._Render makes no second pass over the fine depths."""
    import struct
    from b200eg3d._lib import call, ptr, stream
    g = gen(M + S)
    t_base = torch.linspace(2.25, 3.3, S).cuda()
    u = torch.rand(M, S, generator=g).cuda()
    delta = (3.3 - 2.25) / (S - 1)
    t_c = torch.empty(M, S, device='cuda')
    mm = torch.tensor([-1, 0], dtype=torch.int32, device='cuda')
    call('b200_ray_depths_coarse', ptr(t_base), ptr(u), ptr(t_c), M, S, float(delta), ptr(mm), stream())
    ref = t_base[None, :] + u * delta
    assert (t_c - ref).abs().max().item() < 1e-6        # (the kernel may contract the multiply-add)

    def dec(word):
        w = int(word) & 0xffffffff
        bits = (w & 0x7fffffff) if (w & 0x80000000) else (~w & 0xffffffff)
        return struct.unpack('<f', struct.pack('<I', bits))[0]
    lo, hi = (dec(v) for v in mm.cpu().tolist())
    assert lo == t_c.min().item() and hi == t_c.max().item()
    mm2 = torch.tensor([-1, 0], dtype=torch.int32, device='cuda')
    call('b200_depth_minmax', ptr(t_c), t_c.numel(), ptr(mm2), stream())
    assert torch.equal(mm, mm2)
    if S >= 4:
        sig = (torch.randn(M, S, generator=g) * 4).cuda()
        sig[: M // 8] = -30.0                                # empty rays: flat weights, the degenerate branch of sample_pdf
        u2 = torch.rand(M, S, generator=g).cuda()
        u2[0, 0], u2[0, 1] = 0.0, 1.0 - 2.0 ** -24           # the ends of the unit interval
        t_f = torch.empty(M, S, device='cuda')
        call('b200_ray_importance', ptr(t_c), ptr(sig), ptr(u2), ptr(t_f), M, S, S, stream())
        assert (t_f >= t_c.min(dim=1, keepdim=True).values).all() and (t_f <= t_c.max(dim=1, keepdim=True).values).all()


@pytest.mark.parametrize('n', [1, 2])
@pytest.mark.parametrize('table', ['eg3d', 'odd'])
def test_bank_weight_kernels_match_per_layer_prep(b2, table, n):
    """The grouped ('bank') weight kernels -- every layer of a synthesis network in one launch (b200_bank_weights_fwd / _bwd,
    networks_stylegan2.py:58-67 and its backward) -- against the per-layer entry points b200_modconv_weight_prep(_bwd) on the same
    styles: modulated weights as fp32 and as the split-bf16 pair, demodulation coefficients, d W and d styles.  'eg3d' = shapes the
    vectorised kernels take (every layer of the 512-128 generator has one of them), 'odd' adds a channel count that sends the whole
    table down the generic kernels."""
    import ctypes
    from b200eg3d import _lib
    from b200eg3d._lib import call, ptr, stream
    shapes = [(512, 512, 9, 1), (256, 512, 9, 1), (128, 256, 9, 1), (64, 128, 9, 1), (64, 64, 9, 1), (128, 32, 9, 1), (3, 512, 1, 0), (3, 64, 1, 0),
              (96, 256, 9, 1), (32, 16, 9, 1), (3, 8, 1, 0)]
    if table == 'odd':
        shapes = shapes[:3] + [(40, 24, 9, 1), (3, 24, 1, 0)]
    g = gen(17 + n)
    arr = (_lib.BankLayer * len(shapes))()
    keep = []
    for l, (cout, cin, taps, demod) in enumerate(shapes):
        W = torch.randn(cout, cin, taps, generator=g).cuda()
        s = (torch.randn(n, cin, generator=g) * 0.5 + 1.0).cuda()
        t = dict(W=W, s=s, dcoef=torch.zeros(n, cout, device='cuda'), wmod=torch.zeros(n, taps, cout, cin, device='cuda'),
                 w_hi=torch.zeros(n, taps, cout, cin, device='cuda', dtype=torch.bfloat16), w_lo=torch.zeros(n, taps, cout, cin, device='cuda', dtype=torch.bfloat16),
                 dwmod=(torch.randn(n, taps, cout, cin, generator=g)).cuda(), dW=torch.zeros(cout, cin, taps, device='cuda'), ds=torch.zeros(n, cin, device='cuda'))
        keep.append(t)
        e = arr[l]
        e.weight, e.styles, e.dcoef, e.wmod, e.w_hi, e.w_lo = ptr(W), ptr(s), ptr(t['dcoef']) if demod else None, ptr(t['wmod']), ptr(t['w_hi']), ptr(t['w_lo'])
        e.dwmod, e.d_weight, e.d_styles = ptr(t['dwmod']), ptr(t['dW']), ptr(t['ds'])
        e.widx, e.cin, e.cout, e.taps, e.demod, e.post_scale = 0, cin, cout, taps, demod, 1.0
    call('b200_bank_weights_fwd', ctypes.addressof(arr), len(shapes), n, stream())
    call('b200_bank_weights_bwd', ctypes.addressof(arr), len(shapes), n, stream())
    for (cout, cin, taps, demod), t in zip(shapes, keep):
        wmod, hi, lo = torch.empty_like(t['wmod']), torch.empty_like(t['w_hi']), torch.empty_like(t['w_lo'])
        dcoef = torch.ones(n, cout, device='cuda')
        call('b200_modconv_weight_prep', ptr(t['W']), ptr(t['s']), ptr(wmod), ptr(hi), ptr(lo), ptr(dcoef) if demod else None, n, cout, cin, taps, demod, stream())
        dW, ds = torch.empty_like(t['dW']), torch.empty_like(t['ds'])
        call('b200_modconv_weight_prep_bwd', ptr(t['W']), ptr(t['s']), ptr(dcoef) if demod else None, ptr(t['dwmod']), ptr(dW), ptr(ds), n, cout, cin, taps, demod, stream())
        tag = (cout, cin, taps)
        scale = wmod.abs().max().item()
        assert maxdiff(t['wmod'], wmod) <= 2e-6 * scale, tag
        if demod:
            assert relerr(t['dcoef'], dcoef) < 1e-6, tag
        # the pair reconstructs the fp32 weight to ~2^-17 (two bf16 roundings); both paths split the same way
        assert maxdiff(t['w_hi'].float() + t['w_lo'].float(), wmod) <= 2e-5 * scale, tag
        assert maxdiff(t['w_hi'].float(), hi.float()) <= 2 ** -7 * scale, tag            # (a rounding boundary may flip one bf16 ulp)
        assert relerr(t['dW'], dW) < 2e-5 and relerr(t['ds'], ds) < 2e-5, tag


@pytest.mark.parametrize('cin,cout,npix,n,clamp', [(64, 3, 4097, 1, 256.0), (128, 3, 1000, 2, 0.5), (8, 1, 33, 1, -1.0), (512, 4, 260, 1, 1.0),
                                                    (32, 3, 7, 3, -1.0)])
def test_conv1x1_fwd_thin(b2, cin, cout, npix, n, clamp):
    """b200_conv1x1_fwd_thin (the super-resolution ToRGB layers: modulated 1x1 convolution to <= 4 channels from the split-bf16 pair,
    bias and clamp applied on the way out; networks_stylegan2.py:353-357) against the same sum in fp64."""
    from b200eg3d._lib import call, ptr, stream
    g = gen(cin * 7 + cout + npix)
    x = torch.randn(n, npix, cin, generator=g).cuda()
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    wmod = (torch.randn(n, cout, cin, generator=g) / cin ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    y = torch.empty(n, npix, cout, device='cuda')
    assert b2._lib.load().b200_conv1x1_fwd_thin_supported(cin, cout) == 1
    call('b200_conv1x1_fwd_thin', ptr(hi), ptr(lo), ptr(wmod), ptr(bias), ptr(y), n, npix, cin, cout, float(clamp), stream())
    ref = torch.einsum('npc,noc->npo', (hi.double() + lo.double()), wmod.double()) + bias.double()
    if clamp >= 0:
        ref = ref.clamp(-clamp, clamp)
    assert maxdiff(y, ref) < 2e-5
    y2 = torch.empty_like(y)
    call('b200_conv1x1_fwd_thin', ptr(hi), ptr(lo), ptr(wmod), None, ptr(y2), n, npix, cin, cout, -1.0, stream())
    assert maxdiff(y2, torch.einsum('npc,noc->npo', (hi.double() + lo.double()), wmod.double())) < 2e-5
    assert b2._lib.load().b200_conv1x1_fwd_thin_supported(24, 3) == 0 and b2._lib.load().b200_conv1x1_fwd_thin_supported(64, 5) == 0

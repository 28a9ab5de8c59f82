"""Rebuild the inputs of a golden case (parameters by name-seed, ws, camera, noise draws, targets)."""
import os
import types

import numpy as np
import torch

import eg3d_oracle as oracle
import synth_params as sp

CASE_CFG = sp.GOLDEN_CASES          # shared with oracle/make_goldens.py


def manifest(arch, golden_dir=None):
    """name -> shape of every parameter and buffer of the REAL reference TriPlaneGenerator (recorded by make_goldens.py)."""
    import json
    golden_dir = golden_dir or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    with open(os.path.join(golden_dir, f'manifest_{arch}.json')) as f:
        return json.load(f)


def param_shapes(gk):
    """name -> shape of every parameter / buffer of the reference TriPlaneGenerator architecture
    (training/triplane.py:19-46, networks_stylegan2.py:276-518, superresolution.py:29-43)."""
    cb, cm, w_dim = gk['channel_base'], gk['channel_max'], gk['w_dim']
    shapes = {}

    def layer(pre, cin, cout, res, k=3, noise=True):
        shapes[pre + 'affine.weight'] = (cin, w_dim)
        shapes[pre + 'affine.bias'] = (cin,)
        shapes[pre + 'weight'] = (cout, cin, k, k)
        shapes[pre + 'bias'] = (cout,)
        if noise:
            shapes[pre + 'noise_const'] = (res, res)
            shapes[pre + 'noise_strength'] = ()

    def block(pre, cin, cout, res, img_ch):
        if cin == 0:
            shapes[pre + 'const'] = (cout, res, res)
        else:
            layer(pre + 'conv0.', cin, cout, res)
        layer(pre + 'conv1.', cout, cout, res)
        layer(pre + 'torgb.', cout, img_ch, res, k=1, noise=False)

    ch = {r: min(cb // r, cm) for r in [4, 8, 16, 32, 64, 128, 256]}
    for r in ch:
        block(f'backbone.synthesis.b{r}.', ch[r // 2] if r > 4 else 0, ch[r], r, 96)
    block('superresolution.block0.', 32, 128, 256, 3)
    block('superresolution.block1.', 128, 64, 512, 3)
    shapes['decoder.net.0.weight'] = (64, 32)
    shapes['decoder.net.0.bias'] = (64,)
    shapes['decoder.net.2.weight'] = (33, 64)
    shapes['decoder.net.2.bias'] = (33,)
    return shapes


def load_case(golden_dir, name):
    fx = dict(np.load(os.path.join(golden_dir, name + '.npz'), allow_pickle=False))
    R, S, S_imp, N, yaw, pitch, pseed, wseed, nseed, tseed = fx['meta']
    R, S, S_imp, N = int(R), int(S), int(S_imp), int(N)
    cfg = CASE_CFG[name]
    gk, over = sp.G_KWARGS[cfg['arch']], cfg.get('rk', {})
    assert (R, S, S_imp, N) == (cfg['R'], cfg['S'], cfg['S_imp'], cfg['N']), 'fixture and case table disagree'
    case = types.SimpleNamespace(name=name, fx=fx, R=R, S=S, S_imp=S_imp, N=N, gk=gk, arch=cfg['arch'], param_seed=int(pseed),
                                 noise_mode=cfg.get('noise_mode', 'const'), has_grads=bool(cfg['bwd']))
    case.rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S_imp, **over)
    case.ws = sp.latent_ws(int(wseed), n=N)
    case.c = sp.case_camera(cfg)
    case.u_strat, case.u_imp = oracle.draw_depth_noise(int(nseed), N, R * R, S, max(S_imp, 1),
                                                       tensor_limits=(case.rk['ray_start'] == 'auto'))
    t512, t_raw = sp.targets(int(tseed), R)
    case.t512, case.t_raw = t512.expand(N, -1, -1, -1), t_raw.expand(N, -1, -1, -1)
    import json
    case.randn_seed = int(fx['randn_seed'][0])
    case.randn_shapes = json.loads(str(fx['randn_shapes']))
    return case


def oracle_noise_args(case):
    """(noise_random, density_draws) for oracle.synthesis on this case (None, (None, None) for the const-noise cases)."""
    return sp.replay_normal_draws(case.randn_seed, case.randn_shapes, case.gk)


def check_image(img, fx, tol):
    """Compare a full [N,3,512,512] image with the fixture: every second pixel at `tol` and ALL pixels through 16x16 tile sums."""
    img = img.detach().double().cpu()
    d_sub = float((img[..., ::2, ::2].float().numpy() - fx['image_sub2']).__abs__().max())
    n, c, h, w = img.shape
    tiles = img.reshape(n, c, h // 16, 16, w // 16, 16).sum(dim=(3, 5)).numpy()
    d_tile = float(np.abs(tiles - fx['image_tile16']).max()) / 256.0          # mean abs deviation per pixel of the worst tile
    return d_sub, d_tile


def check_param_grads(named_grads, fx, tol, scalar_floor=0.2, do_assert=True):
    """Per-parameter comparison against the fixture's channel-resolved summaries.  For every gradient tensor:
      'oc' / 'ic'  relative L2 of the vector of per-output-channel (per-input-channel) norms            <= tol
      'chan'       worst single channel norm, relative to max(its reference, 20 % of the typical one)   <= 10 tol
                   (structural check: a permuted / dropped channel block deviates by O(1))
      'samp'       relative L2 over the strided element sample (element-wise agreement)                  <= 2 tol
    0-dim gradients (noise_strength: single cancellation-dominated sums over a whole activation map) are compared on the
    scale of the largest such scalar in the network.  Returns {metric: (worst value, parameter name)}."""
    names = [str(n) for n in fx['grad_names']]
    off = fx['grad_off']
    o0 = i0 = s0 = 0
    worst = {'oc': (0.0, ''), 'ic': (0.0, ''), 'chan': (0.0, ''), 'samp': (0.0, '')}
    limit = {'oc': tol, 'ic': tol, 'chan': 10 * tol, 'samp': 2 * tol}
    fails = []

    def note(key, dev, n):
        if dev > worst[key][0]:
            worst[key] = (dev, n)
        if dev > limit[key]:
            fails.append((n, key, dev))

    scal = max([float(np.sqrt(fx['grad_mom'][k][1])) for k, n in enumerate(names) if off[k][0] == 1 and off[k][1] == 0 and off[k][2] == 1] + [0.0])
    for k, n in enumerate(names):
        no, ni, ns = (int(v) for v in off[k])
        ref_oc, ref_ic, ref_s = fx['grad_oc'][o0:o0 + no], fx['grad_ic'][i0:i0 + ni], fx['grad_samp'][s0:s0 + ns].astype(np.float64)
        o0, i0, s0 = o0 + no, i0 + ni, s0 + ns
        g = named_grads[n]
        assert g is not None, f'{n}: no gradient'
        oc, ic, smp = sp.grad_slices(g)
        assert oc.shape == ref_oc.shape and ic.shape == ref_ic.shape and smp.shape == ref_s.shape, n
        if g.ndim == 0:
            if abs(oc[0] - ref_oc[0]) > tol * max(ref_oc[0], scalar_floor * scal):
                fails.append((n, 'scalar', abs(oc[0] - ref_oc[0]) / max(ref_oc[0], scalar_floor * scal)))
            continue
        for key, a, b in (('oc', oc, ref_oc), ('ic', ic, ref_ic)):
            if b.size == 0:
                continue
            if g.ndim >= 2:                     # 1-D tensors: a "channel norm" is a single element, covered by 'samp'
                note(key, float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)), n)
            floor = 0.2 * float(np.sqrt((b ** 2).mean()))
            note('chan', float(np.max(np.abs(a - b) / np.maximum(b, max(floor, 1e-30)))), n)
        note('samp', float(np.linalg.norm(smp - ref_s) / max(np.linalg.norm(ref_s), 1e-30)), n)
    if do_assert:
        assert not fails, sorted(fails, key=lambda f: -f[2])[:8]
    return worst


def build_param_dict(case, requires_grad=False):
    P = {k: torch.zeros(v) for k, v in param_shapes(case.gk).items()}
    sp.fill_params_(P, case.param_seed)
    if requires_grad:
        for v in P.values():
            v.requires_grad_(True)
    return P


# ----------------------------------------------------------------------------------------------------------------------
# stage-1 (w-projection) fixtures: shared by oracle/make_goldens_stage1.py and the tests

def stage1_feature_net(seed=11):
    """Seeded stand-in for torchvision VGG16.features (same child layout: conv/relu pairs, max-pools after children 3, 8, 15;
    pretrained weights are not available offline).  get_features(..., '14') returns the activation after child 14."""
    import torch.nn as nn
    chans = [(3, 8), (8, 8), None, (8, 16), (16, 16), None, (16, 32), (32, 32), (32, 32), None, (32, 32), (32, 32), (32, 32)]
    layers = []
    for c in chans:
        if c is None:
            layers.append(nn.MaxPool2d(2))
        else:
            layers += [nn.Conv2d(c[0], c[1], 3, padding=1), nn.ReLU(inplace=False)]
    net = nn.Sequential(*layers)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.ndim > 1 else 0.05))
    return net.eval().requires_grad_(False)


def stage1_inputs(name, R, H):
    seed = {'a': 21, 'b': 22}[name]
    g = torch.Generator().manual_seed(seed)
    c0, c1 = sp.camera(0.0, 0.0), sp.camera(0.25, -0.15)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, R), torch.linspace(-1, 1, R), indexing='ij')
    depth = 2.75 - 0.35 * torch.exp(-(xx ** 2 + yy ** 2) * 3) + 0.01 * torch.randn(R, R, generator=g)
    can = torch.nn.functional.interpolate(torch.rand(1, 3, H // 8, H // 8, generator=g) * 2 - 1, size=(H, H), mode='bilinear')
    return {'init_ext': c0[:, :16].reshape(1, 4, 4).clone(), 'extrinsic': c1[:, :16].reshape(1, 4, 4).clone(),
            'intrinsic': c0[0, 16:25].clone(), 'depth': depth.reshape(1, 1, R, R).contiguous(), 'can_image': can.contiguous(),
            'target': torch.rand(1, 3, 256, 256, generator=g) * 2 - 1}


def stage1_pose_net(seed=13):
    """Seeded stand-in for the camera encoder (cam_predictor of w_projector.py:160-171, use_6d branch): target image [1,3,256,256]
    in 0..255 -> 6-D rotation representation near the canonical pose."""
    import torch.nn as nn
    net = nn.Sequential(nn.AvgPool2d(32), nn.Flatten(), nn.Linear(3 * 8 * 8, 6))
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        net[2].weight.copy_(torch.randn(net[2].weight.shape, generator=g) * 2e-5)
        net[2].bias.copy_(torch.tensor([1.0, 0.05, 0.0, 0.02, -1.0, 0.03]))
    return net


def stage1_feature_fn(net, scale=5e-5):
    """The feature-distance network of the fixture: the stand-in stack on the 0..255 image scaled to 0..1, features scaled so that the
    squared feature distance is O(1e3) like the warping and regulariser terms (call signature of the reference's vgg16(...))."""
    return lambda img, resize_images=False, return_lpips=True: net(img * (1.0 / 255.0)) * scale


def stage1_noise_init(named_bufs, seed=17):
    """Start values of the 17 noise buffers (w_projector.py:128-133 draws them with randn_like): one seeded generator, buffer order."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, buf in named_bufs:
            buf.copy_(torch.randn(buf.shape, generator=gen).to(buf.device))


def stage1_iter_setup(R=32, S=8, seed=31):
    """Inputs of the one-iteration w-projection fixture (tests/golden/stage1_iter.npz): tiny generator, stand-in networks, start state."""
    g = torch.Generator().manual_seed(seed)
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S)
    return {'R': R, 'S': S, 'rk': rk, 'gk': sp.G_KWARGS_TINY, 'param_seed': 7,
            'target': torch.rand(3, 512, 512, generator=g) * 2 - 1,
            'w_start': sp.latent_ws(5)[:, :1].clone(),
            'translation': torch.tensor([[0.01, -0.02, 0.015]]),
            'w_noise': torch.randn(1, 1, 512, generator=g) * 0.02,
            'step': 60, 'num_steps': 400, 'w_std': 1.0}

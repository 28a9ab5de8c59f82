"""Rebuild the inputs of a golden case (parameters by name-seed, ws, camera, noise draws, targets)."""
import os
import types

import numpy as np
import torch

import eg3d_oracle as oracle
import synth_params as sp

# must mirror oracle/make_goldens.py CASES (G kwargs, rendering overrides)
CASE_CFG = {
    'tiny_r64_s16': (sp.G_KWARGS_TINY, {}),
    'tiny_r32_s8_n2_white': (sp.G_KWARGS_TINY, {'white_back': True}),
    'tiny_r64_s12_noimp': (sp.G_KWARGS_TINY, {}),
    'full_r64_s16': (sp.G_KWARGS_FULL, {}),
    'full_r128_s48': (sp.G_KWARGS_FULL, {}),
    'full_r256_s96': (sp.G_KWARGS_FULL, {}),
}


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
    gk, over = CASE_CFG[name]
    case = types.SimpleNamespace(name=name, fx=fx, R=R, S=S, S_imp=S_imp, N=N, gk=gk, param_seed=int(pseed))
    case.rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S_imp, **over)
    case.ws = sp.latent_ws(int(wseed), n=N)
    case.c = sp.camera(yaw, pitch, n=N)
    if N > 1:
        case.c[1] = sp.camera(-yaw, pitch * 0.5)[0]
    case.u_strat, case.u_imp = oracle.draw_depth_noise(int(nseed), N, R * R, S, max(S_imp, 1))
    t512, t_raw = sp.targets(int(tseed), R)
    case.t512, case.t_raw = t512.expand(N, -1, -1, -1), t_raw.expand(N, -1, -1, -1)
    return case


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

"""Generate tests/golden/*.npz (+ manifest_*.json, mapping.npz) from the REAL reference (TEST INFRASTRUCTURE ONLY; runs
in the build container).

    python oracle/make_goldens.py [--only NAME]

Imports the unmodified reference from /root/reference (read-only), builds a TriPlaneGenerator with the architecture of the
ffhq512-128 pickles, fills it with the name-seeded synthetic parameters of oracle/synth_params.py, and records what
``G.synthesis(ws, c, noise_mode=..., force_fp32=True)`` returns on CPU (the reference falls back to its own
_bias_act_ref/_upfirdn2d_ref there) plus gradients of the PTI stand-in loss.  /root/reference does not exist on the GPU box,
so the outputs are committed as fixtures.

Fixture content (v2, after the round-1 review called the first version porous):
  image       every second pixel ([::2, ::2], fp32) + float64 sums over all 16x16 tiles of the FULL image + global moments
  image_raw / image_depth   complete
  planes      [::8, ::8] + moments
  gradients   grad_ws, grad_c complete; per parameter: moments, L2 norm of every dim-0 slice (output channel), of every
              dim-1 slice (input channel), and a strided sample of <= 2048 elements -- a wrong channel block cannot hide
  manifest    name -> shape of every parameter / buffer of the real reference class (manifest_{full,tiny}.json)
  mapping     G.mapping outputs (mapping.npz)

Shims applied to the reference while it runs here:
  * ray_sampler.py:38 calls .cuda() unconditionally; Tensor.cuda is made the identity;
  * for the noise_mode='random' / density_noise case torch.randn / torch.randn_like are replaced by a seeded recorder
    (synth_params.SeededNormal) so that the product can be fed the identical draws on the GPU box.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import eg3d_oracle as oracle  # noqa: E402
import synth_params as sp  # noqa: E402

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

CASES = sp.GOLDEN_CASES
PARAM_SEED, WS_SEED, NOISE_SEED, TARGET_SEED, RANDN_SEED = 7, 1, 11, 2, 5


def moments(t):
    t = t.detach().double()
    return np.array([t.sum().item(), t.square().sum().item(), t.min().item(), t.max().item()], dtype=np.float64)


def sub(t, stride):
    return t.detach()[..., ::stride, ::stride].contiguous().numpy()


def tile_sums(t, tile=16):
    n, c, h, w = t.shape
    return t.detach().double().reshape(n, c, h // tile, tile, w // tile, tile).sum(dim=(3, 5)).numpy()


def build_reference(gk, rk):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # ray_sampler.py:38 shim
    from training.triplane import TriPlaneGenerator        # noqa: E402  (reference import)
    G = TriPlaneGenerator(rendering_kwargs=rk, **gk).eval().float()
    named = dict(list(G.named_parameters()) + list(G.named_buffers()))
    sp.fill_params_(named, PARAM_SEED)
    return G, named


def grad_summaries(G):
    params = dict(G.named_parameters())
    names = sorted(n for n, p in params.items() if p.grad is not None and '.mapping.' not in n)
    fx = {'grad_names': np.array(names), 'grad_mom': np.stack([moments(params[n].grad)[:2] for n in names])}
    oc, ic, smp, off = [], [], [], []
    for n in names:
        g = params[n].grad.detach().double()
        o, i, s = sp.grad_slices(g)
        off.append([len(o), len(i), len(s)])
        oc.append(o); ic.append(i); smp.append(s)
    fx['grad_off'] = np.array(off, dtype=np.int64)
    fx['grad_oc'] = np.concatenate(oc).astype(np.float64)
    fx['grad_ic'] = np.concatenate(ic).astype(np.float64) if sum(len(v) for v in ic) else np.zeros([0])
    fx['grad_samp'] = np.concatenate(smp).astype(np.float32)
    return fx


def run_case(name):
    cfg = CASES[name]
    gk, R, S, S_imp, N = sp.G_KWARGS[cfg['arch']], cfg['R'], cfg['S'], cfg['S_imp'], cfg['N']
    yaw, pitch = cfg['cam']
    do_bwd, over, focal = cfg['bwd'], cfg.get('rk', {}), cfg.get('focal', 4.2647)
    noise_mode = cfg.get('noise_mode', 'const')
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S_imp, **over)
    G, named = build_reference(gk, rk)
    G.neural_rendering_resolution = R
    ws = sp.latent_ws(WS_SEED, n=N)
    c = sp.case_camera(cfg)
    t512, t_raw = sp.targets(TARGET_SEED, R)
    t512, t_raw = t512.expand(N, -1, -1, -1), t_raw.expand(N, -1, -1, -1)

    for p in G.parameters():
        p.requires_grad_(do_bwd)
    ws.requires_grad_(do_bwd)
    c.requires_grad_(do_bwd)

    t0 = time.time()
    torch.manual_seed(NOISE_SEED)
    ctx = torch.enable_grad() if do_bwd else torch.no_grad()
    normal = sp.SeededNormal(RANDN_SEED)
    with ctx, normal.patched(enable=(noise_mode == 'random' or rk.get('density_noise', 0) > 0)):
        planes = {}
        h = G.backbone.synthesis.register_forward_hook(lambda m, i, o: planes.__setitem__('p', o))
        out = G.synthesis(ws, c, noise_mode=noise_mode, force_fp32=True)
        h.remove()
        loss = oracle.pti_loss(out, t512, t_raw)
    t_fwd = time.time() - t0
    fx = {
        'meta': np.array([R, S, S_imp, N, yaw, pitch, PARAM_SEED, WS_SEED, NOISE_SEED, TARGET_SEED], dtype=np.float64),
        'randn_seed': np.array([RANDN_SEED]), 'randn_shapes': np.array(json.dumps(normal.shapes)),
        'image_sub2': sub(out['image'], 2), 'image_tile16': tile_sums(out['image']), 'image_mom': moments(out['image']),
        'image_raw': out['image_raw'].detach().numpy(), 'image_depth': out['image_depth'].detach().numpy(),
        'planes_sub8': sub(planes['p'], 8), 'planes_mom': moments(planes['p']),
        'loss': np.array([loss.item()], dtype=np.float64),
    }
    if do_bwd:
        t0 = time.time()
        loss.backward()
        print(f'  bwd {time.time() - t0:.1f}s')
        fx['grad_ws'] = ws.grad.numpy()
        fx['grad_c'] = c.grad.numpy()
        fx.update(grad_summaries(G))

    # Pin the oracle restatement against the reference on the spot (same parameters, same draws).
    P = {k: v.detach() for k, v in named.items()}
    u1, u2 = oracle.draw_depth_noise(NOISE_SEED, N, R * R, S, max(S_imp, 1), tensor_limits=(rk['ray_start'] == 'auto'))
    nr, dd = sp.replay_normal_draws(RANDN_SEED, normal.shapes, gk)
    with torch.no_grad():
        o = oracle.synthesis(P, ws.detach(), c.detach(), rk, R, u1, u2, return_planes=True, noise_mode=noise_mode,
                             noise_random=nr, density_draws=dd)
    for k in ('image', 'image_raw', 'image_depth'):
        d = (o[k] - out[k].detach()).abs().max().item()
        print(f'  oracle vs reference {k}: max-abs {d:.3e}')
        assert d < 2e-4, (name, k, d)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **fx)
    print(f'{name}: fwd {t_fwd:.1f}s loss {loss.item():.6f} -> {os.path.getsize(os.path.join(OUT, name + ".npz")) / 1e3:.0f} kB')


def run_manifest():
    for arch in ('full', 'tiny'):
        G, named = build_reference(sp.G_KWARGS[arch], sp.rendering_kwargs())
        man = {'parameters': {n: list(p.shape) for n, p in G.named_parameters()},
               'buffers': {n: list(b.shape) for n, b in G.named_buffers()},
               'num_ws': int(G.backbone.num_ws), 'source': 'training.triplane.TriPlaneGenerator of /root/reference'}
        with open(os.path.join(OUT, f'manifest_{arch}.json'), 'w') as f:
            json.dump(man, f, indent=0, sort_keys=True)
        print(f'manifest_{arch}: {len(man["parameters"])} parameters, {len(man["buffers"])} buffers')


def run_mapping():
    """G.mapping (triplane.py:48-51 -> networks_stylegan2.py:213-262): plain, truncated, truncated with cutoff."""
    G, named = build_reference(sp.G_KWARGS['full'], sp.rendering_kwargs())
    with torch.no_grad():
        G.backbone.mapping.w_avg.copy_(torch.randn(512, generator=torch.Generator().manual_seed(9)) * 0.5)
    z = torch.randn(2, 512, generator=torch.Generator().manual_seed(3))
    c = torch.cat([sp.camera(0.2, 0.1), sp.camera(-0.3, 0.0)])
    with torch.no_grad():
        fx = {'z': z.numpy(), 'c': c.numpy(), 'w_avg': G.backbone.mapping.w_avg.numpy().copy(), 'param_seed': np.array([PARAM_SEED]),
              'ws_plain': G.mapping(z, c).numpy(), 'ws_psi07': G.mapping(z, c, truncation_psi=0.7).numpy(),
              'ws_psi05_cut8': G.mapping(z, c, truncation_psi=0.5, truncation_cutoff=8).numpy()}
    np.savez_compressed(os.path.join(OUT, 'mapping.npz'), **fx)
    print('mapping: ws', fx['ws_plain'].shape)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    a = ap.parse_args()
    for nm in CASES:
        if a.only in (None, nm):
            print('case', nm)
            run_case(nm)
    if a.only in (None, 'manifest'):
        run_manifest()
    if a.only in (None, 'mapping'):
        run_mapping()

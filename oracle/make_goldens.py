"""Generate tests/golden/*.npz from the REAL reference (TEST INFRASTRUCTURE ONLY; runs in the build container).

    python oracle/make_goldens.py [--only NAME]

Imports the unmodified reference from /root/reference (read-only), builds a
TriPlaneGenerator with the architecture of the ffhq512-128 pickles, fills it with the
name-seeded synthetic parameters of oracle/synth_params.py, and records what
``G.synthesis(ws, c, noise_mode='const', force_fp32=True)`` returns on CPU (the reference
falls back to its own _bias_act_ref/_upfirdn2d_ref there) plus gradients of the PTI
stand-in loss.  /root/reference does not exist on the GPU box, so the outputs are
committed as compact fixtures: full small tensors, strided sub-samples + float64
moments of big ones.

The one shim applied: ray_sampler.py:38 calls .cuda() unconditionally; Tensor.cuda is
made the identity for the duration of this script.
"""
import argparse
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

CASES = {
    # name: (G kwargs, R, S, S_imp, batch, (yaw,pitch), backward?, rendering overrides)
    'tiny_r64_s16': (sp.G_KWARGS_TINY, 64, 16, 16, 1, (0.0, 0.0), True, {}),
    'tiny_r32_s8_n2_white': (sp.G_KWARGS_TINY, 32, 8, 8, 2, (0.25, -0.15), True, {'white_back': True}),
    'tiny_r64_s12_noimp': (sp.G_KWARGS_TINY, 64, 12, 0, 1, (-0.3, 0.1), False, {}),
    'full_r64_s16': (sp.G_KWARGS_FULL, 64, 16, 16, 1, (0.0, 0.0), False, {}),            # BASELINE config 1
    'full_r128_s48': (sp.G_KWARGS_FULL, 128, 48, 48, 1, (0.3, -0.2), True, {}),           # BASELINE config 2/4 (grads incl. pose)
    'full_r256_s96': (sp.G_KWARGS_FULL, 256, 96, 96, 1, (-0.2, 0.15), False, {}),         # BASELINE config 5
}
PARAM_SEED, WS_SEED, NOISE_SEED, TARGET_SEED = 7, 1, 11, 2


def moments(t):
    t = t.detach().double()
    return np.array([t.sum().item(), t.square().sum().item(), t.min().item(), t.max().item()], dtype=np.float64)


def sub(t, stride):
    return t.detach()[..., ::stride, ::stride].contiguous().numpy()


def build_reference(gk, rk):
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # ray_sampler.py:38 shim
    from training.triplane import TriPlaneGenerator        # noqa: E402  (reference import)
    G = TriPlaneGenerator(rendering_kwargs=rk, **gk).eval().float()
    named = dict(list(G.named_parameters()) + list(G.named_buffers()))
    sp.fill_params_(named, PARAM_SEED)
    return G, named


def run_case(name):
    gk, R, S, S_imp, N, (yaw, pitch), do_bwd, over = CASES[name]
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S_imp, **over)
    G, named = build_reference(gk, rk)
    G.neural_rendering_resolution = R
    ws = sp.latent_ws(WS_SEED, n=N)
    c = sp.camera(yaw, pitch, n=N)
    if N > 1:   # make the batch entries different
        c[1] = sp.camera(-yaw, pitch * 0.5)[0]
    t512, t_raw = sp.targets(TARGET_SEED, R)
    t512, t_raw = t512.expand(N, -1, -1, -1), t_raw.expand(N, -1, -1, -1)

    for p in G.parameters():
        p.requires_grad_(do_bwd)
    ws.requires_grad_(do_bwd)
    c.requires_grad_(do_bwd)

    t0 = time.time()
    torch.manual_seed(NOISE_SEED)
    ctx = torch.enable_grad() if do_bwd else torch.no_grad()
    with ctx:
        planes = {}
        h = G.backbone.synthesis.register_forward_hook(lambda m, i, o: planes.__setitem__('p', o))
        out = G.synthesis(ws, c, noise_mode='const', force_fp32=True)
        h.remove()
        loss = oracle.pti_loss(out, t512, t_raw)
    t_fwd = time.time() - t0
    fx = {
        'meta': np.array([R, S, S_imp, N, yaw, pitch, PARAM_SEED, WS_SEED, NOISE_SEED, TARGET_SEED], dtype=np.float64),
        'image_sub4': sub(out['image'], 4), 'image_mom': moments(out['image']),
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
        names = sorted(n for n, p in G.named_parameters() if p.grad is not None and '.mapping.' not in n)
        fx['grad_names'] = np.array(names)
        fx['grad_mom'] = np.stack([moments(dict(G.named_parameters())[n].grad)[:2] for n in names])
        fx['grad_head'] = np.stack([np.resize(dict(G.named_parameters())[n].grad.reshape(-1)[:16].numpy(), 16) for n in names])

    # Pin the oracle restatement against the reference on the spot (same parameters, same draws).
    P = {k: v.detach() for k, v in named.items()}
    u1, u2 = oracle.draw_depth_noise(NOISE_SEED, N, R * R, S, max(S_imp, 1))
    with torch.no_grad():
        o = oracle.synthesis(P, ws.detach(), c.detach(), rk, R, u1, u2, return_planes=True)
    for k in ('image', 'image_raw', 'image_depth'):
        d = (o[k] - out[k].detach()).abs().max().item()
        print(f'  oracle vs reference {k}: max-abs {d:.3e}')
        assert d < 2e-4, (name, k, d)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **fx)
    print(f'{name}: fwd {t_fwd:.1f}s loss {loss.item():.6f} -> {os.path.getsize(os.path.join(OUT, name + ".npz")) / 1e3:.0f} kB')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    a = ap.parse_args()
    for nm in CASES:
        if a.only in (None, nm):
            print('case', nm)
            run_case(nm)

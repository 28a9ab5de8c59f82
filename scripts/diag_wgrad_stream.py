"""Run-to-run variation of the parameter gradients with the wgrad stream on / off (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ('3dgan-inversion_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, d))
import torch
import b200eg3d
import synth_params as sp
from golden_util import load_case

case = load_case(os.path.join(ROOT, 'tests', 'golden'), sys.argv[1] if len(sys.argv) > 1 else 'tiny_r64_s16')
G = b200eg3d.TriPlaneGenerator(rendering_kwargs=case.rk, **case.gk).eval()
sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), case.param_seed)
G = G.cuda().float()
G.neural_rendering_resolution = case.R
G.renderer.fixed_noise = (case.u_strat.cuda(), case.u_imp.cuda())
ws, c = case.ws.cuda(), case.c.cuda()
named = [(n, p) for n, p in G.named_parameters() if '.mapping.' not in n]


def run(on):
    b200eg3d.ops.CONFIG['wgrad_stream'] = on
    for _, p in named:
        p.grad = None
    out = G.synthesis(ws, c, noise_mode='const')
    (out['image'].square().mean() + out['image_raw'].square().mean()).backward()
    torch.cuda.synchronize()
    return [p.grad.clone() if p.grad is not None else None for _, p in named]


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


seq = [True, True, True, False, False, False, True, False]
res = [run(o) for o in seq]
for i in range(1, len(seq)):
    worst = max((rel(a, b), n) for (n, _), a, b in zip(named, res[i], res[i - 1]) if a is not None)
    print(f'run {i} ({seq[i]}) vs run {i - 1} ({seq[i - 1]}): worst rel-L2 {worst[0]:.3e} at {worst[1]}')

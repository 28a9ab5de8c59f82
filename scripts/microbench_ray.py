"""Time the per-ray kernels (importance sampling, compositing forward / backward) alone on random inputs of the metric's shape
(16384 rays, 48 + 48 samples) and of BASELINE config 5 (65536 rays, 96 + 96): 10 back-to-back launches, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200'))
import torch
from b200eg3d._lib import call, ptr, stream

dev = 'cuda'
torch.manual_seed(0)


def timeit(fn, name, nbytes, iters=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f'{name:44s} {ms * 1e3:8.1f} us   {nbytes / ms / 1e9:6.2f} TB/s (compulsory bytes)')


for M, S in ((128 * 128, 48), (256 * 256, 96)):
    t_c = (torch.rand(M, S, device=dev) * 0.02 + torch.linspace(2.25, 3.3, S, device=dev)).contiguous()
    t_f = (torch.rand(M, S, device=dev) * 1.05 + 2.25).contiguous()
    sc, sf = torch.randn(M, S, device=dev) * 3, torch.randn(M, S, device=dev) * 3
    rc, rf = torch.rand(M, S, 32, device=dev), torch.rand(M, S, 32, device=dev)
    u = torch.rand(M, S, device=dev)
    mm = torch.zeros(2, dtype=torch.int32, device=dev); mm[:1].fill_(-1)
    call('b200_depth_minmax', ptr(t_c), t_c.numel(), ptr(mm), stream())
    call('b200_depth_minmax', ptr(t_f), t_f.numel(), ptr(mm), stream())
    feat, depth, wsum = torch.empty(M, 32, device=dev), torch.empty(M, device=dev), torch.empty(M, device=dev)
    g_feat, g_depth = torch.randn(M, 32, device=dev), torch.randn(M, device=dev)
    d_rc, d_sc, d_rf, d_sf = torch.empty_like(rc), torch.empty_like(sc), torch.empty_like(rf), torch.empty_like(sf)
    t_out = torch.empty(M, S, device=dev)
    tag = f'{M} rays x {S}+{S}'
    timeit(lambda: call('b200_ray_importance', ptr(t_c), ptr(sc), ptr(u), ptr(t_out), M, S, S, stream()), f'ray_importance      {tag}', M * S * 16)
    timeit(lambda: call('b200_ray_composite_fwd', ptr(t_c), ptr(sc), ptr(rc), S, ptr(t_f), ptr(sf), ptr(rf), S, ptr(mm), 0, M, ptr(feat),
                        ptr(depth), ptr(wsum), stream()), f'ray_composite_fwd   {tag}', M * (2 * S * 136 + 136))
    timeit(lambda: call('b200_ray_composite_bwd', ptr(t_c), ptr(sc), ptr(rc), S, ptr(t_f), ptr(sf), ptr(rf), S, ptr(mm), 0, M, ptr(g_feat),
                        ptr(g_depth), None, ptr(d_rc), ptr(d_sc), ptr(d_rf), ptr(d_sf), stream()), f'ray_composite_bwd   {tag}',
           M * (2 * S * (136 + 132) + 136))

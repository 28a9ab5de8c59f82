"""Time the fused tri-plane sample+MLP kernels in isolation (both generations, variants of the backward) -- optimisation aid.

    python scripts/microbench_triplane.py [--fine]     --fine: unsorted importance-style depths instead of stratified ones
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import synth_params as sp
import eg3d_oracle as oracle
from b200eg3d import _lib
from b200eg3d._lib import call, ptr, stream

dev = 'cuda'
torch.manual_seed(0)
n, R, S = 1, 128, 48
M = R * R
planes = torch.randn(n, 256, 256, 96, device=dev)
c = sp.camera(0.3, -0.2)
ro, rd = oracle.ray_sampler(c[:, :16].reshape(-1, 4, 4), c[:, 16:].reshape(-1, 3, 3), R)
ro, rd = ro.contiguous().to(dev), rd.contiguous().to(dev)
if '--fine' in sys.argv:
    t = (2.25 + torch.rand(n, M, S, device=dev) * 1.05).contiguous()
else:
    t = (torch.linspace(2.25, 3.3, S, device=dev)[None, None] + torch.rand(n, M, S, device=dev) * (1.05 / 47)).contiguous()
W1, b1, W2, b2 = [torch.randn(*s, device=dev) for s in [(64, 32), (64,), (33, 64), (33,)]]
rgb = torch.empty(n, M * S, 32, device=dev); sig = torch.empty(n, M * S, device=dev)
d_rgb = torch.randn_like(rgb); d_sig = torch.randn_like(sig)
dpl = torch.zeros_like(planes); dpts = torch.empty(n, M * S, 3, device=dev)
dro, drd = torch.zeros_like(ro), torch.zeros_like(rd)
dW = [torch.zeros_like(x) for x in (W1, b1, W2, b2)]
work = torch.empty(_lib.load().b200_triplane_bwd_workspace_bytes(n, M * S), device=dev, dtype=torch.uint8)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)          # > L2: evict between timed launches


def timeit(fn, name, iters=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f'{name:48s} {tot / iters:8.3f} ms', flush=True)


fsave = torch.empty(_lib.load().b200_triplane_fsave_bytes(n, M * S), device=dev, dtype=torch.uint8)


def fwd(rw, save=False):
    return lambda: call('b200_triplane_mlp_fwd', ptr(planes), n, 256, 256, None, ptr(ro), ptr(rd), ptr(t), S, rw, M * S, 1.0, ptr(W1), ptr(b1),
                        ptr(W2), ptr(b2), 1.0, ptr(rgb), ptr(sig), ptr(fsave) if save else None, stream())


def bwd(dp, dc, wg, rays=False, rw=0):
    return lambda: call('b200_triplane_mlp_bwd', ptr(planes), n, 256, 256, None, ptr(ro), ptr(rd), ptr(t), S, rw, M * S, 1.0, ptr(W1), ptr(b1), ptr(W2),
                        ptr(b2), 1.0, ptr(d_rgb), ptr(d_sig), ptr(fsave), ptr(dpl) if dp else None, ptr(dpts) if dc else None,
                        ptr(dro) if rays else None, ptr(drd) if rays else None, *[(ptr(x) if wg else None) for x in dW], ptr(work), work.numel(), stream())


for impl, name in ((1, 'tcgen05'), (0, 'mma.sync')):
    _lib.load().b200_set_triplane_impl(impl)
    timeit(fwd(R), f'[{name}] fwd 786k pts, 8x16 patch order')
    timeit(fwd(0), f'[{name}] fwd 786k pts, linear ray order')
    timeit(fwd(0, save=True), f'[{name}] fwd 786k pts, linear, saving features')
    timeit(bwd(True, False, True), f'[{name}] bwd planes+wgrad (PTI)')
    timeit(bwd(True, False, False), f'[{name}] bwd planes only')
    timeit(bwd(False, False, True), f'[{name}] bwd wgrad only')
    timeit(bwd(False, False, False), f'[{name}] bwd neither (recompute + chain)')
    timeit(bwd(True, False, False, rays=True), f'[{name}] bwd planes+rays (w-projection)')

"""Time the fused tri-plane sample+MLP kernels in isolation (variants of the backward) -- optimisation aid."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import synth_params as sp
import eg3d_oracle as oracle
from b200eg3d._lib import call, ptr, stream

dev = 'cuda'
torch.manual_seed(0)
n, R, S = 1, 128, 48
M = R * R
planes = torch.randn(n, 256, 256, 96, device=dev)
c = sp.camera(0.3, -0.2)
ro, rd = oracle.ray_sampler(c[:, :16].reshape(-1, 4, 4), c[:, 16:].reshape(-1, 3, 3), R)
ro, rd = ro.contiguous().to(dev), rd.contiguous().to(dev)
t = (torch.linspace(2.25, 3.3, S, device=dev)[None, None] + torch.rand(n, M, S, device=dev) * (1.05 / 47)).contiguous()
W1, b1, W2, b2 = [torch.randn(*s, device=dev) for s in [(64, 32), (64,), (33, 64), (33,)]]
rgb = torch.empty(n, M * S, 32, device=dev); sig = torch.empty(n, M * S, device=dev)
d_rgb = torch.randn_like(rgb); d_sig = torch.randn_like(sig)
dpl = torch.zeros_like(planes); dpts = torch.empty(n, M * S, 3, device=dev)
dW = [torch.zeros_like(x) for x in (W1, b1, W2, b2)]



def timeit(fn, name, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:40s} {e0.elapsed_time(e1) / iters:8.3f} ms')

fwd = lambda: call('b200_triplane_mlp_fwd', ptr(planes), n, 256, 256, None, ptr(ro), ptr(rd), ptr(t), S, M * S, 1.0, ptr(W1), ptr(b1), ptr(W2), ptr(b2), 1.0, ptr(rgb), ptr(sig), stream())
def bwd(dp, dc, wg):
    return lambda: call('b200_triplane_mlp_bwd', ptr(planes), n, 256, 256, None, ptr(ro), ptr(rd), ptr(t), S, M * S, 1.0, ptr(W1), ptr(b1), ptr(W2), ptr(b2), 1.0,
                        ptr(d_rgb), ptr(d_sig), ptr(dpl) if dp else None, ptr(dpts) if dc else None, *[(ptr(x) if wg else None) for x in dW], None, 0, stream())
timeit(fwd, 'fwd 786k pts')
timeit(bwd(True, False, True), 'bwd planes+wgrad (PTI)')
timeit(bwd(True, False, False), 'bwd planes only')
timeit(bwd(False, False, True), 'bwd wgrad only')
timeit(bwd(False, False, False), 'bwd neither (recompute + chain)')
timeit(bwd(True, True, False), 'bwd planes+coords (w-projection)')

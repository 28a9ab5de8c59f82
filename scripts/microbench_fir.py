"""Time the fused 4x4 FIR (+ layer epilogue + bf16 split) after every up=2 convolution, forward and adjoint shapes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200'))
import torch
from b200eg3d._lib import call, ptr, stream
from b200eg3d import ops
dev = 'cuda'
SEP = 0 if '--patch' in sys.argv else 1          # --patch: the 2 x 4 patch kernel (non-separable path)
f = ops.fir_filter(torch.device(dev))
for (res, c) in [(8, 512), (16, 512), (32, 512), (64, 512), (128, 256), (256, 128), (256, 128), (512, 64)]:
    x = torch.randn(1, res + 1, res + 1, c, device=dev)
    z = torch.empty(1, res, res, c, device=dev); zh = torch.empty(1, res, res, c, device=dev, dtype=torch.bfloat16); zl = torch.empty_like(zh)
    b = torch.randn(c, device=dev); nz = torch.randn(res, res, device=dev); st = torch.full([], 0.1, device=dev)
    dy = torch.randn(1, res, res, c, device=dev)
    gh = torch.empty(1, res + 1, res + 1, c, device=dev, dtype=torch.bfloat16); gl = torch.empty_like(gh)
    fwd = lambda: call('b200_upfirdn2d_fused', ptr(x), ptr(f), None, ptr(z), ptr(zh), ptr(zl), 1, res + 1, res + 1, c, 4, 4, 1, 1, 1, 1, 1, 1, 0, 4.0,
                       1, ptr(b), ptr(nz), ptr(st), 0, 1, 0.2, 1.414, 256.0, SEP, stream())
    bwd = lambda: call('b200_upfirdn2d_fused', ptr(dy), ptr(f), None, None, ptr(gh), ptr(gl), 1, res, res, c, 4, 4, 1, 1, 2, 2, 2, 2, 1, 4.0,
                       0, None, None, None, 0, 0, 0.0, 1.0, -1.0, SEP, stream())
    out = []
    for fn in (fwd, bwd):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 20 * 1e3)
    mb = (x.numel() * 4 + z.numel() * 8) / 1e6
    print(f'res {res:4d} c {c:4d}: fwd {out[0]:7.1f} us  bwd {out[1]:7.1f} us   ({mb:6.1f} MB -> {mb / 6.5e3 * 1e3:5.1f} us at 6.5 TB/s)')

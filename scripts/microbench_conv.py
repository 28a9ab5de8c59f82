"""Time (and let ncu profile) the tensor-core conv kernels on two representative layers -- optimisation / evidence aid."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200'))
import torch
from b200eg3d._lib import call, ptr, stream

dev = 'cuda'
torch.manual_seed(0)


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16).contiguous()


ONCE = '--once' in sys.argv          # ncu captures: launch every kernel exactly once


def timeit(fn, name, flops, iters=10):
    fn(); torch.cuda.synchronize()
    if ONCE:
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f'{name:46s} {ms * 1e3:8.1f} us   {flops / ms / 1e9:8.1f} TFLOP/s (algorithmic)')


for (name, cin, cout, res) in [('sr.block1.conv1  64->64  @512^2', 64, 64, 512), ('b64.conv1  512->512 @64^2', 512, 512, 64),
                               ('b256.conv1 128->128 @256^2', 128, 128, 256)]:
    n, h, w, k = 1, res, res, 3
    xh, xl = bf(n, h, w, cin), bf(n, h, w, cin)
    wh, wl = bf(n, 9, cout, cin), bf(n, 9, cout, cin)
    dh, dl = bf(n, h, w, cout), bf(n, h, w, cout)
    y = torch.empty(n, h, w, cout, device=dev)
    dx = torch.empty(n, h, w, cin, device=dev)
    dw = torch.empty(n, 9, cout, cin, device=dev)
    fl = 2.0 * h * w * 9 * cin * cout
    timeit(lambda: call('b200_conv_fwd_tc', ptr(xh), ptr(xl), ptr(wh), ptr(wl), ptr(y), n, h, w, cin, cout, k, 1, 3, stream()), name + ' fwd x3', fl)
    timeit(lambda: call('b200_conv_dgrad_tc', ptr(dh), ptr(dl), ptr(wh), ptr(wl), ptr(dx), n, h, w, cin, cout, k, 1, 3, stream()), name + ' dgrad x3', fl)
    timeit(lambda: call('b200_conv_wgrad_tc', ptr(xh), None, ptr(dh), None, ptr(dw), n, h, w, cin, cout, k, 1, 1, 0, stream()), name + ' wgrad x1', fl)

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


COLD = '--cold' in sys.argv          # flush the L2 (write 512 MB) before every timed launch: operands come from HBM as in the step
_flush = torch.empty(128 * 1024 * 1024, device=dev, dtype=torch.float32) if COLD else None


def timeit(fn, name, flops, iters=10):
    fn(); torch.cuda.synchronize()
    if ONCE:
        return
    if COLD:
        ts = []
        for _ in range(5):
            _flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        print(f'{name:46s} {ms * 1e3:8.1f} us   {flops / ms / 1e9:8.1f} TFLOP/s (algorithmic, cold L2)')
        return ms
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f'{name:46s} {ms * 1e3:8.1f} us   {flops / ms / 1e9:8.1f} TFLOP/s (algorithmic)')
    return ms


LAYERS = [('b64.conv0   512->512 up  32->64', 512, 512, 32, 2), ('b64.conv1   512->512 @64^2', 512, 512, 64, 1),
          ('b128.conv0  512->256 up  64->128', 512, 256, 64, 2), ('b128.conv1  256->256 @128^2', 256, 256, 128, 1),
          ('b256.conv0  256->128 up 128->256', 256, 128, 128, 2), ('b256.conv1  128->128 @256^2', 128, 128, 256, 1),
          ('sr0.conv0    32->128 up 128->256', 32, 128, 128, 2), ('sr0.conv1   128->128 @256^2', 128, 128, 256, 1),
          ('sr1.conv0   128->64  up 256->512', 128, 64, 256, 2), ('sr1.conv1    64->64  @512^2', 64, 64, 512, 1)]
tot = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
for (name, cin, cout, res, up) in LAYERS:
    n, h, w, k = 1, res, res, 3
    hs = res if up == 1 else 2 * res + 1
    xh, xl = bf(n, h, w, cin), bf(n, h, w, cin)
    wh, wl = bf(n, 9, cout, cin), bf(n, 9, cout, cin)
    dh, dl = bf(n, hs, hs, cout), bf(n, hs, hs, cout)
    y = torch.empty(n, hs, hs, cout, device=dev)
    dx = torch.empty(n, h, w, cin, device=dev)
    dw = torch.empty(n, 9, cout, cin, device=dev)
    fl = 2.0 * h * w * 9 * cin * cout
    for kind, fn in (('fwd', lambda: call('b200_conv_fwd_tc', ptr(xh), ptr(xl), ptr(wh), ptr(wl), ptr(y), n, h, w, cin, cout, k, up, 3, 0, stream())),
                     ('dgrad', lambda: call('b200_conv_dgrad_tc', ptr(dh), ptr(dl), ptr(wh), ptr(wl), ptr(dx), n, h, w, cin, cout, k, up, 3, 0, stream())),
                     ('wgrad', lambda: call('b200_conv_wgrad_tc', ptr(xh), None, ptr(dh), None, ptr(dw), n, h, w, cin, cout, k, up, 1, int('--wgrad-accumulate' in sys.argv), stream()))):
        if kind == 'wgrad' and '--only-fwd-dgrad' in sys.argv:
            continue
        if kind != 'wgrad' and '--only-wgrad' in sys.argv:
            continue
        ms = timeit(fn, f'{name} {kind}', fl)
        if ms:
            tot[kind] += ms
print('totals (us):', {k_: round(v * 1e3, 1) for k_, v in tot.items()})

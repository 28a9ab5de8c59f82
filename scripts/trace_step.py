"""Per-kernel device times of the GRAPH-REPLAYED PTI step from a CUPTI trace (torch.profiler): what each kernel really costs
inside the step (warm caches, two-branch graph), next to ncu's serialised cold-cache launch list."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import bench
from b200eg3d.graphs import GraphedStep
from b200eg3d import losses

dev = torch.device('cuda', 0)
G, ws, c, t512, _ = bench.make_problem(0, dev)
params = [p for n, p in G.named_parameters() if '.mapping.' not in n]
from b200eg3d.optim import Adam
opt = Adam(params, lr=3e-4)
res = [t.to(dev) for t in (ws, c, t512)]

def step(ws, c, t512):
    out = G.synthesis(ws, c, noise_mode='const', force_fp32=True)
    loss = losses.pti_loss(out, t512)
    loss.backward(); opt.step()
    return loss

g = GraphedStep(step, res, optimizer=opt, warmup=3)
for _ in range(5): g(*res)
torch.cuda.synchronize()
N = 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(N): g(*res)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
for e in evs:
    k = e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:70]
    agg[k][0] += 1; agg[k][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
tot = sum(v[1] for v in agg.values())
out = [f'# graph-replayed PTI step, CUPTI kernel trace over {N} replays (B200EG3D_PDL={os.environ.get("B200EG3D_PDL", "1")}, B200EG3D_OVERLAP={os.environ.get("B200EG3D_OVERLAP", "1")}): '
       f'busy time {tot / N / 1e3:.3f} ms per step (sum over both branches; with PDL a kernel\'s time includes its griddepcontrol.wait)']
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    out.append(f'{v[1] / N / 1e3:9.3f} ms {v[0] / N:7.1f} launches {100 * v[1] / tot:6.2f}%  {k}')
print('\n'.join(out))
open(os.path.join(ROOT, 'gpurun_out', 'trace_step_pdl%s_ov%s.txt' % (os.environ.get('B200EG3D_PDL', '1'), os.environ.get('B200EG3D_OVERLAP', '1'))), 'w').write('\n'.join(out) + '\n')

# ---- timeline of the LAST replay: start, duration, stream of every kernel; idle gaps of the union; per-stream busy time
try:
    kev = [e for e in prof.profiler.kineto_results.events() if e.device_type() == torch.autograd.DeviceType.CUDA and e.duration_ns() > 0]
    kev.sort(key=lambda e: e.start_ns())
    per = len(kev) // N
    last = kev[-per:]
    t0 = last[0].start_ns()
    lines, busy_end, idle, streams = [], t0, 0.0, collections.defaultdict(float)
    for e in last:
        s, d = e.start_ns() - t0, e.duration_ns()
        gap = e.start_ns() - busy_end
        if gap > 0:
            idle += gap
        busy_end = max(busy_end, e.start_ns() + d)
        streams[e.device_resource_id()] += d
        nm = e.name().replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:60]
        lines.append(f'{s / 1e3:9.1f} us  +{d / 1e3:7.1f} us  stream {e.device_resource_id():3d}  {"gap %.1f" % (gap / 1e3) if gap > 1500 else "":10s} {nm}')
    span = busy_end - t0
    head = [f'# timeline of one graph replay: span {span / 1e6:.3f} ms, no kernel running for {idle / 1e6:.3f} ms of it',
            '# busy time per stream (ms): ' + ', '.join(f'{k}: {v / 1e6:.3f}' for k, v in sorted(streams.items()))]
    open(os.path.join(ROOT, 'gpurun_out', 'timeline_step.txt'), 'w').write('\n'.join(head + lines) + '\n')
    print('\n'.join(head))
except Exception as ex:                      # profiler internals differ between torch versions
    print('timeline unavailable:', ex)

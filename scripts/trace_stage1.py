"""Per-kernel device times of the graph-replayed stage-1 (w-projection) iteration from a CUPTI trace -- where its 8.5 ms go."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
st, noise = bench.make_stage1(dev, graphed=True)
for _ in range(5):
    st.step(noise)
torch.cuda.synchronize()
N = 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        st.step(noise)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
for e in evs:
    k = e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:80]
    agg[k][0] += 1; agg[k][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
tot = sum(v[1] for v in agg.values())
span = (max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)) / N
out = [f'# graph-replayed stage-1 iteration, CUPTI kernel trace over {N} replays: busy time {tot / N / 1e3:.3f} ms per iteration (sum over all streams), wall span {span / 1e3:.3f} ms']
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:50]:
    out.append(f'{v[1] / N / 1e3:9.3f} ms {v[0] / N:7.1f} launches {100 * v[1] / tot:6.2f}%  {k}')
print('\n'.join(out))
open(os.path.join(ROOT, 'gpurun_out', 'trace_stage1.txt'), 'w').write('\n'.join(out) + '\n')

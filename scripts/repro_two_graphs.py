import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, bench
from b200eg3d.coach import PTIStep
dev = torch.device('cuda', 0)
order = sys.argv[1] if len(sys.argv) > 1 else 'ab'
for tag in order:
    r, s = (128, 48) if tag == 'a' else (256, 96)
    G, ws, c, t512, _ = bench.make_problem(0, dev, r, s)
    inputs = [t.to(dev) for t in (ws, c, t512)]
    try:
        st = PTIStep(G, graphed=True, example=inputs)
        print(tag, 'captured; loss', float(st.step(*inputs)), flush=True)
    except Exception as e:
        print(tag, 'FAILED', type(e).__name__, str(e)[:300], flush=True)
        break

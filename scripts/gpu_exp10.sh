timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4 | cut -c1-200
timeout 300 python bench.py --steps 100 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], 'final_loss', d['final_loss'])"
timeout 300 python bench.py --stage w_projection --steps 50 2>/dev/null | tail -1 | cut -c150-330
B200EG3D_OVERLAP=0 timeout 300 python bench.py --stage w_projection --steps 50 2>/dev/null | tail -1 | cut -c150-330
timeout 300 python bench.py --steps 5 --eager --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('eager', d['value'], d['ms_per_step'], 'final_loss', d['final_loss'])"

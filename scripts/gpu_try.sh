timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
timeout 300 python bench.py --no-cpu --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('long_run',{}).get('median_ms'), d['final_loss']); print({k:v for k,v in d['kernel_ms'].items() if 'upfirdn' in k or 'fir4' in k or 'act' in k})"

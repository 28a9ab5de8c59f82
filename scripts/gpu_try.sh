timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('long_run',{}).get('median_ms')); print(d['extra'])"

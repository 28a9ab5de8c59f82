mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5 | cut -c1-250
timeout 300 python bench.py --steps 50 --no-cpu > gpurun_out/bench_x.log 2>&1; tail -1 gpurun_out/bench_x.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print({k:v for k,v in d['per_call_ms'].items()})"

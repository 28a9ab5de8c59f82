timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
for zp in 1 0; do echo "== ZERO_POOLS=$zp"; B200EG3D_ZERO_POOLS=$zp timeout 300 python bench.py --no-cpu --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('long_run',{}).get('median_ms'), d['final_loss'], d['kernel_ms'].get('Memset '), d['gpu_launches'])"; done

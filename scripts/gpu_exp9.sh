for ov in 0 1; do
echo "=== OVERLAP=$ov"
B200EG3D_OVERLAP=$ov timeout 300 python -m pytest tests/test_gpu_golden.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
B200EG3D_OVERLAP=$ov timeout 300 python bench.py --steps 50 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], 'final_loss', d['final_loss'])"
done

for dp in 3 1; do
echo "=== DGRAD_PASSES=$dp"
B200EG3D_DGRAD_PASSES=$dp timeout 300 python -m pytest tests/test_gpu_golden.py -m gpu -q -s --tb=short -p no:cacheprovider -k "gradients" 2>&1 | grep -E "rel-L2|worst|passed|failed|Error" | cut -c1-200
B200EG3D_DGRAD_PASSES=$dp timeout 300 python bench.py --steps 50 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['per_call_ms']['b200_conv_dgrad_tc'])"
done

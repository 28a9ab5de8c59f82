timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "ray_composite or render" 2>&1 | tail -15 | cut -c1-250
bash scripts/gpu_exp6.sh

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "run_model or render" 2>&1 | tail -15 | cut -c1-250
python scripts/microbench_triplane.py

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "pti_loss" 2>&1 | tail -15 | cut -c1-250
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5 | cut -c1-250
timeout 300 python bench.py --steps 50 --no-cpu > gpurun_out/bench_x.log 2>&1; tail -1 gpurun_out/bench_x.log | cut -c1-330

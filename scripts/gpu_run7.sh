# conv-focused check: op tests first (fast fail), conv microbench, then the full GPU suite and a bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "modconv or torgb" > gpurun_out/pytest7a.log 2>&1
echo "ops rc=$?"; tail -3 gpurun_out/pytest7a.log | cut -c1-200
timeout 200 python scripts/microbench_conv.py 2>&1 | grep -v "^$" | cut -c1-120
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest7.log 2>&1
echo "all rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest7.log | cut -c1-220 | head -20
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench7.log 2>&1
tail -1 gpurun_out/bench7.log | cut -c1-300

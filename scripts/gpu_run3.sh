mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "modconv or torgb" -p no:cacheprovider > gpurun_out/pytest3a.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/pytest3a.log | cut -c1-250
timeout 600 python -m pytest tests -m gpu -q --tb=short -rP -p no:cacheprovider --deselect tests/test_gpu_ops.py > gpurun_out/pytest3b.log 2>&1
echo "rc=$?"; grep -E "max-abs|rel-L2|worst|passed|failed" gpurun_out/pytest3b.log | cut -c1-200
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench3.log 2>&1
tail -1 gpurun_out/bench3.log | cut -c1-2500

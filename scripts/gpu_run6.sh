mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -rP -p no:cacheprovider > gpurun_out/pytest6.log 2>&1
echo "rc=$?"; grep -E "max-abs|rel-L2|worst|passed|failed|^FAILED|^E  " gpurun_out/pytest6.log | cut -c1-220 | head -40
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench6.log 2>&1
tail -1 gpurun_out/bench6.log | cut -c1-2800

mkdir -p gpurun_out
python scripts/microbench_triplane.py 2>&1 | tail -7
timeout 900 python -m pytest tests -m gpu -q --tb=short -rP -p no:cacheprovider > gpurun_out/pytest4.log 2>&1
echo "rc=$?"; grep -E "max-abs|rel-L2|worst|passed|failed|^FAILED|^E  " gpurun_out/pytest4.log | cut -c1-220
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench4.log 2>&1
tail -1 gpurun_out/bench4.log | cut -c1-2600

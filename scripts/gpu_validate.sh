# usage: bash scripts/gpu_validate.sh <tag>   full GPU suite, default bench line, ncu launch list of one eager step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_$1.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_$1.log | cut -c1-220 | head -20
timeout 600 python bench.py > gpurun_out/bench_$1.log 2>&1
tail -1 gpurun_out/bench_$1.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 1 --warmup 3 --no-cpu --eager --ncu-step > gpurun_out/ncu_l_$1.log 2>&1
ls -la gpurun_out | tail -4
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$1.log 2>&1; tail -2 gpurun_out/smoke_$1.log

# usage: bash scripts/gpu_final.sh <tag>    the round's closing evidence (one GPU, ~5 min): full GPU suite twice (the second pass looks for
#   order-of-atomics flakiness), smoke(), the default bench line (all legs), ncu launch list of ONE eager step, CUPTI timeline of one graph replay
mkdir -p gpurun_out
T=${1:-r2z5}
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_${T}.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_${T}.log | cut -c1-220 | head -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${T}.log 2>&1; tail -2 gpurun_out/smoke_${T}.log
timeout 900 python bench.py > gpurun_out/bench_${T}.log 2>&1
tail -1 gpurun_out/bench_${T}.log > gpurun_out/${T}_bench.json; cut -c1-400 gpurun_out/${T}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --eager --ncu-step > gpurun_out/${T}_launches.log 2>&1
python scripts/summarize_ncu.py gpurun_out/${T}_launches.csv gpurun_out/${T}_launches_one_step.txt
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1
cp gpurun_out/timeline_step.txt gpurun_out/${T}_timeline_graph_step.txt; cp gpurun_out/trace_step_pdl1_ov1.txt gpurun_out/${T}_trace_graph_step.txt
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_${T}_pass2.log 2>&1
echo "pytest pass 2 rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_${T}_pass2.log | cut -c1-220 | head -20
head -12 gpurun_out/${T}_launches_one_step.txt; tail -3 gpurun_out/${T}_trace.log

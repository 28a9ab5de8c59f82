# usage: bash scripts/gpu_quick2.sh <tag>   render / golden / graphed tests, the short bench line, the graph timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "render or golden or graphed" > gpurun_out/pytest_$1.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_$1.log | cut -c1-220 | head
timeout 300 python bench.py --no-cpu --no-extra --steps 200 --warmup 10 > gpurun_out/bench_$1.log 2>&1
tail -1 gpurun_out/bench_$1.log > gpurun_out/$1_bench_noextra.json; cut -c90-200 gpurun_out/$1_bench_noextra.json
timeout 200 python scripts/trace_step.py > gpurun_out/$1_trace.log 2>&1; cp gpurun_out/timeline_step.txt gpurun_out/$1_timeline_graph_step.txt
grep -n "ray_composite_bwd" -B3 -A8 gpurun_out/$1_timeline_graph_step.txt | cut -c1-120

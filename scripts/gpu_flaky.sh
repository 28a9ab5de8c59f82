# usage: bash scripts/gpu_flaky.sh [n]   repeat the graph-vs-eager tests and collect the printed deviations (is a failure a race or tolerance noise?)
mkdir -p gpurun_out
: > gpurun_out/flaky.log
for i in $(seq 1 ${1:-6}); do
echo "== run $i" >> gpurun_out/flaky.log
timeout 200 python -m pytest tests/test_gpu_graphed.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|^E  |loss eager|worst|stage-1" | cut -c1-500 >> gpurun_out/flaky.log
done
grep -E "==|passed|failed" gpurun_out/flaky.log | paste - - | cut -c1-120

# usage: bash scripts/gpu_flaky.sh    repeat the graph-vs-eager tests under both gradient-fork settings (is a failure a race or tolerance noise?)
mkdir -p gpurun_out
: > gpurun_out/flaky.log
for fork in 1 0 1 0; do for i in 1 2 3; do
echo "== fork=$fork run $i" >> gpurun_out/flaky.log
B200EG3D_FORK_GRADS=$fork timeout 200 python -m pytest tests/test_gpu_graphed.py -m gpu -q --tb=short -p no:cacheprovider -s -k "graphed_pti" 2>&1 | grep -E "passed|failed|^E  |loss eager|worst" | cut -c1-400 >> gpurun_out/flaky.log
done; done
grep -E "==|passed|failed" gpurun_out/flaky.log | paste - - | cut -c1-120

# usage: bash scripts/gpu_suite_repeat.sh [n]   the full GPU suite n times (tolerances that sit on atomics-order noise show up as occasional failures)
mkdir -p gpurun_out
for i in $(seq 1 ${1:-4}); do
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/suite_rep_$i.log 2>&1
echo "pass $i rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/suite_rep_$i.log | cut -c1-250 | head -12
done

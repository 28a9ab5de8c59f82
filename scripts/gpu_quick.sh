# usage: bash scripts/gpu_quick.sh <tag> [pytest -k expression] [ENV=0 switch for the comparison run]
#   quick check of a change: selected GPU tests, then the bench line without the secondary legs, with and without the switch
mkdir -p gpurun_out
K=${2:-"two_addends or forked or lean or graphed or golden"}
OFF=${3:-B200EG3D_FORK_GRADS=0}
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "$K" > gpurun_out/pytest_$1.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_$1.log | cut -c1-220 | head -20
timeout 300 python bench.py --no-cpu --no-extra --steps 200 --warmup 10 > gpurun_out/bench_$1.log 2>&1
tail -1 gpurun_out/bench_$1.log | cut -c90-200
env $OFF timeout 300 python bench.py --no-cpu --no-extra --steps 200 --warmup 10 > gpurun_out/bench_$1_off.log 2>&1
echo "$OFF"; tail -1 gpurun_out/bench_$1_off.log | cut -c90-200

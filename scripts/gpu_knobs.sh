mkdir -p gpurun_out
for kb in 18 9; do
B200EG3D_ACT_FUSE_MIN_KBLOCKS=$kb timeout 300 python bench.py --no-cpu --no-extra --steps 200 --warmup 10 > gpurun_out/bench_knob_kb$kb.log 2>&1
echo "kb=$kb"; tail -1 gpurun_out/bench_knob_kb$kb.log | cut -c90-200
done
timeout 300 python bench.py --no-cpu --no-extra --steps 200 --warmup 10 > gpurun_out/bench_knob_base.log 2>&1
echo base; tail -1 gpurun_out/bench_knob_base.log | cut -c90-200
timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "graphed" 2>&1 | tail -3

# usage: bash scripts/gpu_final_prof.sh <tag>    (default bench line + ncu launch list of exactly one eager step)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_$1.log 2>&1
tail -1 gpurun_out/bench_$1.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 1 --warmup 3 --no-cpu --eager --ncu-step > gpurun_out/ncu_l.log 2>&1
ls -la gpurun_out | tail -4

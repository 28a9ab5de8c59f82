# usage: bash scripts/gpu_final_prof.sh <tag>    (bench + launch list of one step + full ncu metrics of the top kernels)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_$1.log 2>&1
tail -1 gpurun_out/bench_$1.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 1 --warmup 3 --no-cpu --eager --ncu-step > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc_pix_kernel|conv_tc_wgrad|triplane_mlp_(fwd|bwd)_mma" -o gpurun_out/prof_$1 python bench.py --steps 1 --warmup 3 --no-cpu --eager --ncu-step > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -5

mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 1 --warmup 3 --no-cpu --eager --ncu-step > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-300

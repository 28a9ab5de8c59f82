mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench5.log 2>&1
tail -2 gpurun_out/bench5.log | cut -c1-900
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --eager > gpurun_out/bench5e.log 2>&1
tail -1 gpurun_out/bench5e.log | cut -c1-400

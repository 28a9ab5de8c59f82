mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest1.log 2>&1
tail -5 gpurun_out/pytest1.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench1.log 2>&1
tail -3 gpurun_out/bench1.log

# usage: bash scripts/gpu_mg2_quick.sh   (2 GPUs) tensors on a non-current device + one short 2-rank bench line
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_golden.py -m gpu -q --tb=short -p no:cacheprovider -k "non_current or retained" 2>&1 | tail -3
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-extra > gpurun_out/mg2_quick.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/mg2_quick.log | cut -c90-230)"
grep -n -i "Traceback\|Error" gpurun_out/mg2_quick.log | head -5

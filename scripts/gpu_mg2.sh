# usage: bash scripts/gpu_mg2.sh [runs]    repeat the 2-GPU bench to catch intermittent failures; full logs under gpurun_out/
for i in $(seq 1 ${1:-3}); do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520 + i)) bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/mg2_$i.log 2>&1
  echo "run $i rc=$? $(tail -1 gpurun_out/mg2_$i.log | cut -c1-160)"
  grep -n -i "Traceback\|Error\|error:" gpurun_out/mg2_$i.log | head -5
done

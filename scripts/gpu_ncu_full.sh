# usage: bash scripts/gpu_ncu_full.sh <tag>   ncu --set full of the tri-plane kernels (fwd, bwd with parameter gradients), of the
# forward / dgrad / wgrad tensor-core conv kernel on three layers, and of the ray compositing kernels (one launch each)
mkdir -p gpurun_out
T=$1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:triplane_mlp_fwd -s 1 -c 1 -o gpurun_out/tpfwd_$T -f python scripts/microbench_triplane.py > gpurun_out/ncu_a_$T.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:triplane_mlp_bwd -s 1 -c 1 -o gpurun_out/tpbwd_$T -f python scripts/microbench_triplane.py > gpurun_out/ncu_b_$T.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 0 -c 9 -o gpurun_out/conv_$T -f python scripts/microbench_conv.py --once > gpurun_out/ncu_c_$T.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ray_composite -c 2 -o gpurun_out/ray_$T -f python bench.py --steps 1 --warmup 3 --no-cpu --eager --ncu-step > gpurun_out/ncu_d_$T.log 2>&1
for k in tpfwd tpbwd conv ray; do ncu -i gpurun_out/${k}_$T.ncu-rep --page details > gpurun_out/${k}_${T}_details.txt 2>&1; done
ls -la gpurun_out | tail -8

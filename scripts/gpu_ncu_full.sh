# usage: bash scripts/gpu_ncu_full.sh <tag>   ncu --set full captures (one launch each): tri-plane kernels (fwd, bwd with parameter
# gradients), tensor-core conv kernels (CTA-pair forward / dgrad, wgrad) on the large layers, the FIR / activation-backward
# elementwise kernels and the ray compositing kernels of one eager step
mkdir -p gpurun_out
T=$1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:triplane_fwd_tc -s 1 -c 1 -o gpurun_out/tpfwd_$T -f python scripts/microbench_triplane.py > gpurun_out/ncu_a_$T.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:triplane_bwd_tc -s 1 -c 1 -o gpurun_out/tpbwd_$T -f python scripts/microbench_triplane.py > gpurun_out/ncu_b_$T.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 15 -c 6 -o gpurun_out/conv_$T -f python scripts/microbench_conv.py --once > gpurun_out/ncu_c_$T.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:ray_composite|fir4_strip|layer_act_bwd_vec" -c 12 -o gpurun_out/elem_$T -f python bench.py --steps 1 --warmup 3 --no-cpu --eager --ncu-step > gpurun_out/ncu_d_$T.log 2>&1
for k in tpfwd tpbwd conv elem; do ncu -i gpurun_out/${k}_$T.ncu-rep --page details > gpurun_out/${k}_${T}_details.txt 2>&1; done
ls -la gpurun_out | tail -8

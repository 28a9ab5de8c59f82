# usage: bash scripts/gpu_ncu_full.sh <tag>   ncu --set full of the tri-plane kernels and of three conv layers (one launch each)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triplane_mlp -s 2 -c 2 -o gpurun_out/tp_$1 -f python scripts/microbench_triplane.py > gpurun_out/ncu_tp_$1.log 2>&1
ncu -i gpurun_out/tp_$1.ncu-rep --page details > gpurun_out/tp_$1_details.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 0 -c 9 -o gpurun_out/conv_$1 -f python scripts/microbench_conv.py > gpurun_out/ncu_conv_$1.log 2>&1
ncu -i gpurun_out/conv_$1.ncu-rep --page details > gpurun_out/conv_$1_details.txt 2>&1
python scripts/microbench_triplane.py; python scripts/microbench_conv.py
ls -la gpurun_out | tail -6

# usage: bash scripts/gpu_prof_k.sh <tag>   bench line + one-step launch list + ncu --set full of the tri-plane backward (1 launch)
bash scripts/gpu_final_prof.sh $1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triplane_mlp_bwd -s 1 -c 1 -o gpurun_out/tpbwd_$1 -f python scripts/microbench_triplane.py > gpurun_out/ncu_tp.log 2>&1
ncu -i gpurun_out/tpbwd_$1.ncu-rep --page details > gpurun_out/tpbwd_$1_details.txt 2>&1
tail -8 gpurun_out/ncu_tp.log

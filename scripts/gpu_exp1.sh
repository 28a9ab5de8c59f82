mkdir -p gpurun_out
V=3dgan-inversion_b200/b200eg3d/variants
for v in bm12 bm8 fw12; do echo "== $v"; B200EG3D_LIB=$PWD/$V/lib_$v.so python scripts/microbench_triplane.py; done
echo "== base"; python scripts/microbench_triplane.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triplane_mlp_bwd -s 1 -c 1 -o gpurun_out/tpbwd_r1l -f python scripts/microbench_triplane.py > gpurun_out/ncu_tpb.log 2>&1
ncu -i gpurun_out/tpbwd_r1l.ncu-rep --page details > gpurun_out/tpbwd_r1l_details.txt 2>&1

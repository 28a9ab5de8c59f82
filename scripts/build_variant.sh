# usage: bash scripts/build_variant.sh <name> "<extra nvcc flags>"  ->  3dgan-inversion_b200/b200eg3d/variants/lib_<name>.so
# (kernel-tuning aid: select at run time with B200EG3D_LIB=<path>)
set -e
D=3dgan-inversion_b200/b200eg3d
mkdir -p $D/variants /tmp/b200_variant_$1
for src in $D/csrc/*.cu; do
  f=$(basename $src .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden $2 -c $src -o /tmp/b200_variant_$1/$f.o &
done
wait
nvcc -shared -o $D/variants/lib_$1.so /tmp/b200_variant_$1/*.o -lcudart
echo built $D/variants/lib_$1.so

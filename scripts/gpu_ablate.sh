# usage: bash scripts/gpu_ablate.sh   conv kernel ablations (variants built by scripts/build_variant.sh with -DB200_DBG_*)
D=3dgan-inversion_b200/b200eg3d/variants
for pair in 1 0; do
for v in "" NO_TMA NO_MMA; do
  echo "=== pair=$pair variant=${v:-full}"
  if [ -z "$v" ]; then LIB=""; else LIB="$PWD/$D/lib_$v.so"; fi
  B200EG3D_CONV_PAIR=$pair B200EG3D_LIB=$LIB timeout 200 python scripts/microbench_conv.py --only-fwd-dgrad 2>&1 | grep -E "b128.conv1|b256.conv1|sr1.conv1|b64.conv1" | grep -v wgrad
done
done

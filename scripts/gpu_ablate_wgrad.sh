# usage: bash scripts/gpu_ablate_wgrad.sh   weight-gradient kernel ablations (variants built by scripts/build_variant.sh with -DB200_DBG_*)
D=3dgan-inversion_b200/b200eg3d/variants
for v in "" NO_TMA NO_MMA NO_STORE; do
  echo "=== variant=${v:-full}"
  if [ -z "$v" ]; then LIB=""; else LIB="$PWD/$D/lib_$v.so"; fi
  B200EG3D_LIB=$LIB timeout 200 python scripts/microbench_conv.py --only-wgrad --wgrad-accumulate 2>&1 | grep -E "wgrad|totals"
done

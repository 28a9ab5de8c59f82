V=3dgan-inversion_b200/b200eg3d/variants
for v in bm8 bm10 bm16; do echo "== $v"; B200EG3D_LIB=$PWD/$V/lib_$v.so python scripts/microbench_triplane.py; done

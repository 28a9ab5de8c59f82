for pair in 1 0; do echo "== cold pair=$pair"; B200EG3D_CONV_PAIR=$pair timeout 300 python scripts/microbench_conv.py --cold 2>&1 | tail -32; done

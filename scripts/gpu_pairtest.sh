# usage: bash scripts/gpu_pairtest.sh   CTA-pair conv tiles: op tests, then per-layer timings pair vs single CTA
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "pair_tiles or modconv or torgb or wgrad_pool" -p no:cacheprovider 2>&1 | tail -8
B200EG3D_CONV_PAIR=1 timeout 300 python scripts/microbench_conv.py 2>&1 | tail -40

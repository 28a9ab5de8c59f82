timeout 900 python -m pytest tests/test_gpu_ops.py -q -x --tb=short -k "split_output or split_input or lean_activ" -p no:cacheprovider 2>&1 | tail -15

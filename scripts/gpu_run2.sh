mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rP -p no:cacheprovider > gpurun_out/pytest2.log 2>&1
tail -3 gpurun_out/pytest2.log
python __graft_entry__.py smoke > gpurun_out/smoke2.log 2>&1; tail -2 gpurun_out/smoke2.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.log 2>&1
tail -1 gpurun_out/bench2.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_pix_kernel|triplane_mlp_bwd" -s 60 -c 4 -o gpurun_out/prof_r1a python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out

# usage: bash scripts/gpu_profile_r2z3.sh <tag>     evidence for the per-ray kernel rewrite (one GPU, ~4 min):
#   per-ray microbenchmarks, ncu launch list of ONE eager PTI step (+ per-kernel summary), CUPTI timeline of one graph replay,
#   ncu --set full of the compositing / importance kernels inside one eager step, compute-sanitizer (memcheck + racecheck) over the
#   per-ray op tests (all merge orders, all shapes) and the fused render test
mkdir -p gpurun_out
T=${1:-r2z3}
timeout 100 python scripts/microbench_ray.py > gpurun_out/${T}_microbench_ray.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --eager --ncu-step > gpurun_out/${T}_launches.log 2>&1
python scripts/summarize_ncu.py gpurun_out/${T}_launches.csv gpurun_out/${T}_launches_one_step.txt
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1
cp gpurun_out/timeline_step.txt gpurun_out/${T}_timeline_graph_step.txt; cp gpurun_out/trace_step_pdl1_ov1.txt gpurun_out/${T}_trace_graph_step.txt
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:ray_composite|ray_importance" -c 3 -o gpurun_out/${T}_ray -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --eager --ncu-step > gpurun_out/${T}_ncu_ray.log 2>&1
ncu -i gpurun_out/${T}_ray.ncu-rep --page details > gpurun_out/${T}_ray_ncu_full.txt 2>&1
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "ray_composite_merge_orders or (render_fwd and 12-12 and tcgen05)" \
      > gpurun_out/sanitize_${T}_${tool}.log 2>&1
  echo "exit $?" >> gpurun_out/sanitize_${T}_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit |Error:|hazard" gpurun_out/sanitize_${T}_${tool}.log | sort | uniq -c | head -20 > gpurun_out/${T}_sanitize_${tool}.txt
  cat gpurun_out/${T}_sanitize_${tool}.txt
done
head -24 gpurun_out/${T}_launches_one_step.txt; tail -3 gpurun_out/${T}_trace.log

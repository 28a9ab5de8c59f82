# usage: bash scripts/gpu_sanitize.sh <tag>
# compute-sanitizer over the kernels whose correctness rests on hand-rolled mbarrier phases, TMEM hand-offs and shared-memory tile
# reuse: memcheck (out-of-bounds / misaligned global + shared accesses) and racecheck (shared-memory hazards) on small cases of
# the tcgen05 tri-plane kernels, the tcgen05 convolution kernels and the per-ray kernels.  Summaries -> gpurun_out/sanitize_<tag>_*.txt
mkdir -p gpurun_out
T=${1:-run}
# round 2 additions: the CTA-pair (cta_group::2) convolution tiles (cluster barriers, remote mbarrier arrivals, multicast commits),
# the sliding-window FIR, the specialised activation backward, the one-launch Adam and the split-input thin weight gradient
SEL='(run_model and (32 or 1000) and tcgen05) or (render_fwd and 12-12 and tcgen05) or (modconv_layer_fwd_bwd and 64) or (ray_composite_merge_orders and 7-13) or (pair_tiles and (fwd-1-120 or fwd-2-96 or dgrad-2-100 or dgrad-1-72)) or (fir_column and 130) or (split_output and 64) or split_input or adam_matches'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL" \
      > gpurun_out/sanitize_${T}_${tool}.log 2>&1
  echo "exit $?" >> gpurun_out/sanitize_${T}_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit |Error:|hazard" gpurun_out/sanitize_${T}_${tool}.log | sort | uniq -c | head -20 > gpurun_out/sanitize_${T}_${tool}.txt
  cat gpurun_out/sanitize_${T}_${tool}.txt
done

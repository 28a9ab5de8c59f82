# usage: bash scripts/gpu_profile_final.sh <tag>     evidence set of the round's final state (one GPU, ~5 min):
#   ncu launch list of ONE eager PTI step (+ per-kernel summary), CUPTI timeline of one graph replay, warm / cold per-layer conv
#   timings pair vs single CTA, FIR and tri-plane microbenchmarks, ncu --set full of the CTA-pair conv kernels / wgrad and of the
#   elementwise + ray kernels of one eager step (details pages as text)
mkdir -p gpurun_out
T=${1:-final}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --eager --ncu-step > gpurun_out/${T}_launches.log 2>&1
python scripts/summarize_ncu.py gpurun_out/${T}_launches.csv gpurun_out/${T}_launches_one_step.txt
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1
cp gpurun_out/timeline_step.txt gpurun_out/${T}_timeline_graph_step.txt; cp gpurun_out/trace_step_pdl1_ov1.txt gpurun_out/${T}_trace_graph_step.txt
{ echo "# per-layer tensor-core conv timings (scripts/microbench_conv.py): 10 back-to-back launches, operands L2-warm; fwd / dgrad 3-pass, wgrad 1-pass"
  echo "## CTA pairs (default)"; B200EG3D_CONV_PAIR=1 timeout 300 python scripts/microbench_conv.py 2>&1 | tail -31
  echo "## single CTA (B200EG3D_CONV_PAIR=0)"; B200EG3D_CONV_PAIR=0 timeout 300 python scripts/microbench_conv.py 2>&1 | tail -31
  echo "## CTA pairs, L2 flushed before every launch (--cold)"; B200EG3D_CONV_PAIR=1 timeout 300 python scripts/microbench_conv.py --cold 2>&1 | tail -31
} > gpurun_out/${T}_microbench_conv.txt
{ echo "# 4x4 FIR + layer epilogue (scripts/microbench_fir.py): column-window kernel, then the 2x4 patch kernel (--patch)"
  timeout 100 python scripts/microbench_fir.py; echo "## --patch"; timeout 100 python scripts/microbench_fir.py --patch; } > gpurun_out/${T}_microbench_fir.txt 2>&1
timeout 300 python scripts/microbench_triplane.py > gpurun_out/${T}_microbench_triplane.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 15 -c 6 -o gpurun_out/${T}_conv -f python scripts/microbench_conv.py --once > gpurun_out/${T}_ncu_c.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:ray_composite|fir4_col|layer_act_bwd_fast|adam_multi" -c 14 -o gpurun_out/${T}_elem -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --eager --ncu-step > gpurun_out/${T}_ncu_d.log 2>&1
for k in conv elem; do ncu -i gpurun_out/${T}_$k.ncu-rep --page details > gpurun_out/${T}_${k}_ncu_full.txt 2>&1; done
head -30 gpurun_out/${T}_launches_one_step.txt; tail -3 gpurun_out/${T}_trace.log

# usage: bash scripts/gpu_profile_r2.sh <tag>     round-2 evidence set (one GPU):
#   launch list of ONE eager PTI step, ncu --set full of the tcgen05 tri-plane kernels, dram bytes per launch
mkdir -p gpurun_out
T=${1:-r2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --eager --ncu-step > gpurun_out/${T}_launches.log 2>&1
python scripts/summarize_ncu.py gpurun_out/${T}_launches.csv gpurun_out/${T}_launches_one_step.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:triplane_fwd_tc -s 13 -c 1 -o gpurun_out/${T}_tpfwd -f python scripts/microbench_triplane.py > gpurun_out/${T}_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:triplane_bwd_tc -s 1 -c 1 -o gpurun_out/${T}_tpbwd -f python scripts/microbench_triplane.py > gpurun_out/${T}_ncu_b.log 2>&1
for k in tpfwd tpbwd; do
  ncu -i gpurun_out/${T}_${k}.ncu-rep --page details > gpurun_out/${T}_${k}_ncu_full.txt 2>&1
  ncu -i gpurun_out/${T}_${k}.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]
for r in rows[2:3]:
    d = dict(zip(h, r))
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'lts__t_bytes.sum', 'sm__inst_executed.sum', 'launch__registers_per_thread'):
        print(k, d.get(k))
" > gpurun_out/${T}_${k}_raw.txt
  cat gpurun_out/${T}_${k}_raw.txt
done
head -25 gpurun_out/${T}_launches_one_step.txt

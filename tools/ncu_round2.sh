#!/bin/sh
# ncu evidence of round 2 (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are copied to profiles/):
#  1. launch list of the bench command (eager, no CUDA graph so that every launch is visible), per-launch durations
#  2. one --set full capture of the dominant launch (conv 128->128 @32x32 fprop, grouped batch 250) -> raw metrics csv
set -x
export TGAN_NO_PDL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches.csv 40 > gpurun_out/r2_ncu_launch_list_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 4 -c 1 -o gpurun_out/r2_igemm250 -f \
    python tools/prof_conv250.py > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/r2_igemm250.ncu-rep --page raw --csv > gpurun_out/r2_igemm250_raw.csv 2>/dev/null
#  3. --set full of the weight-packing launches (pack_multi_kernel: classifier, discriminator, generator) of one eager step
ncu --set full --clock-control none --import-source on -k regex:pack_multi -c 4 -o gpurun_out/r2_pack_multi -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_pack.log 2>&1
ncu -i gpurun_out/r2_pack_multi.ncu-rep --page raw --csv > gpurun_out/r2_pack_multi_raw.csv 2>/dev/null

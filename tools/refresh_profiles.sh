#!/bin/sh
# copy the evidence of one validation run (gpurun_out/, tag = run number) into profiles/ -- usage: sh tools/refresh_profiles.sh 39 r2k
set -e
n=$1; tag=$2
cp gpurun_out/bench$n.json profiles/r2_bench_1gpu.json
cp gpurun_out/bench${n}_ref.json profiles/r2_bench_reference_arm.json
for w in svhn mnist; do
  grep '^{' gpurun_out/bench${n}_$w.json | tail -1 > profiles/r2_bench_$w.json
  cp gpurun_out/prof_step_agg_${tag}_$w.txt profiles/r2_step_kernels_agg_$w.txt
done
cp gpurun_out/prof_step_agg_$tag.txt profiles/r2_step_kernels_agg.txt
cp gpurun_out/prof_step_seq_$tag.txt profiles/r2_step_kernels_seq.txt
cp gpurun_out/r2_launches.csv profiles/r2_ncu_launch_list_bench_nograph.csv
cp gpurun_out/r2_ncu_launch_list_summary.txt profiles/r2_ncu_launch_list_summary.txt
python tools/refresh_ncu_summaries.py

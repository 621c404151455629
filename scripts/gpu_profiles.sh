#!/bin/bash
# ncu evidence for the bench command (launch list) and the dominant kernels (full captures); round 2 file names
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2_bench_2p24.csv \
   python bench.py --steps 2 --warmup 3 --no-verify --no-extras > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 2 -c 1 -o gpurun_out/prof_accumulate_r2 \
   python scripts/one_commit.py 24 0 4 0 > gpurun_out/ncu_full_acc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:sort_tile_scatter|sort_tile_hist|sort_group_local|msm_recode' -s 7 -c 7 -o gpurun_out/prof_sort_r2 \
   python scripts/one_commit.py 24 0 3 0 > gpurun_out/ncu_full_sort.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:msm_bucket_chunks|msm_bucket_reduce|msm_merge_level|xyzz_tree_reduce' -s 9 -c 9 -o gpurun_out/prof_reduce_r2 \
   python scripts/one_commit.py 24 0 3 0 > gpurun_out/ncu_full_reduce.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw metric pages (CSV) of the large captures, not the reports
for r in sort reduce; do
  ncu -i gpurun_out/prof_${r}_r2.ncu-rep --page raw --csv > gpurun_out/prof_${r}_r2.raw.csv 2>/dev/null && rm -f gpurun_out/prof_${r}_r2.ncu-rep
done
ncu -i gpurun_out/prof_accumulate_r2.ncu-rep --page raw --csv > gpurun_out/prof_accumulate_r2.raw.csv 2>/dev/null
ls -la gpurun_out | tail -8

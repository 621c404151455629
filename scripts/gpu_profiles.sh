#!/bin/bash
# ncu evidence for the bench command (launch list) and the dominant kernel (full capture)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-verify --no-extras > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 2 -c 1 -o gpurun_out/prof_accumulate_r1 \
   python bench.py --steps 1 --warmup 3 --no-verify --no-extras > gpurun_out/ncu_full_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_tile_scatter -s 6 -c 1 -o gpurun_out/prof_scatter_r1 \
   python bench.py --steps 1 --warmup 3 --no-verify --no-extras > gpurun_out/ncu_full_scatter.log 2>&1
ls -la gpurun_out | tail -8

#!/bin/bash
# second GPU session: full test suite, bench N=1, ncu launch list + full capture of msm_accumulate
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 2 --warmup 3 --no-verify --no-extras > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 2 -c 2 -o gpurun_out/prof_accumulate_r1 \
   python bench.py --steps 1 --warmup 3 --no-verify --no-extras > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out/

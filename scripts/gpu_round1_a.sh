#!/bin/bash
# first GPU session: IMAD peak, parity tests, rough timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
./bench/imad_peak > gpurun_out/imad_peak.json 2> gpurun_out/imad_peak.err; cat gpurun_out/imad_peak.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python scripts/quick_time.py 16 20 22 2>&1 | tail -5

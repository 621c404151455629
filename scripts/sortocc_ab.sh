#!/bin/bash
# A/B of the 256-way scatter pass: blocks per SM (2 = register prefetch, 3 / 4 = occupancy instead)
mkdir -p gpurun_out
: > gpurun_out/sortocc_ab.jsonl
for v in 2 3 4; do
  echo "{\"MZ_SORT_OCC\": $v}" >> gpurun_out/sortocc_ab.jsonl
  MZ_SORT_OCC=$v python scripts/phase_sweep.py ${SIZES:-21:20 24:22} 2>&1 | grep -E "log2n|rror" >> gpurun_out/sortocc_ab.jsonl
done
cut -c1-240 gpurun_out/sortocc_ab.jsonl

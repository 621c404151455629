#!/bin/bash
# sort / accumulate pipeline of the resident MSM: chunk count and growth ratio (development aid)
mkdir -p gpurun_out
: > gpurun_out/pipe_sweep.jsonl
for v in ${VARIANTS:-"1 4" "2 4" "3 4" "4 4" "2 8" "3 8" "3 2"}; do
  set -- $v
  echo "{\"MZ_PIPE_CHUNKS\": $1, \"MZ_PIPE_RATIO\": $2}" >> gpurun_out/pipe_sweep.jsonl
  MZ_PIPE_CHUNKS=$1 MZ_PIPE_RATIO=$2 python scripts/phase_sweep.py ${SIZES:-21:20 22:20 24:22} 2>&1 | grep -E "log2n|rror" >> gpurun_out/pipe_sweep.jsonl
done
cat gpurun_out/pipe_sweep.jsonl

#!/bin/bash
# sweep of the two-level bucket reduce: level-1 chunk length, blocks per SM (development aid)
mkdir -p gpurun_out
: > gpurun_out/reduce_sweep.jsonl
for minb in ${MINBS:-2 3}; do for lb in ${LBS:-0 1 16 32}; do
  echo "{\"MZ_REDUCE_LB\": $lb, \"MZ_REDUCE_MINB\": $minb}" >> gpurun_out/reduce_sweep.jsonl
  MZ_REDUCE_LB=$lb MZ_REDUCE_MINB=$minb python scripts/phase_sweep.py 21:20 24:22 2>&1 | grep log2n >> gpurun_out/reduce_sweep.jsonl
done; done

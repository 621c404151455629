#!/bin/bash
# A/B of the MSD group-local sort against the two-pass LSD form (development aid)
mkdir -p gpurun_out
: > gpurun_out/sort_ab.jsonl
for v in msd lsd; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/sort_ab.jsonl
  if [ $v = lsd ]; then export MZ_SORT_LSD=1; else unset MZ_SORT_LSD; fi
  python scripts/phase_sweep.py ${SIZES:-20:20 21:20 22:20 24:22} 2>&1 | grep -E "log2n|rror" >> gpurun_out/sort_ab.jsonl
done
cat gpurun_out/sort_ab.jsonl

#!/bin/bash
# A/B of the MSD sort's group size: fill = expected group size / capacity above which r drops by one (development aid)
mkdir -p gpurun_out
: > gpurun_out/sortfill_ab.jsonl
for v in 95 60 30; do
  echo "{\"MZ_SORT_GROUP_FILL\": $v}" >> gpurun_out/sortfill_ab.jsonl
  MZ_SORT_GROUP_FILL=$v python scripts/phase_sweep.py ${SIZES:-20:20 21:20 22:20 24:22} 2>&1 | grep -E "log2n|rror" >> gpurun_out/sortfill_ab.jsonl
done
cut -c1-150 gpurun_out/sortfill_ab.jsonl

#!/bin/bash
# A/B of the head merge's fan (heads folded serially per lane and level); development aid
mkdir -p gpurun_out
: > gpurun_out/mergefan_ab.jsonl
for v in 16 8 4 32; do
  echo "{\"MZ_MERGE_FAN\": $v}" >> gpurun_out/mergefan_ab.jsonl
  MZ_MERGE_FAN=$v python scripts/phase_sweep.py ${SIZES:-21:20 24:22} 2>&1 | grep -E "log2n|rror" >> gpurun_out/mergefan_ab.jsonl
done
cut -c1-60,120-260 gpurun_out/mergefan_ab.jsonl

#!/bin/bash
# multi-GPU session: sharded commit/open parity over NCCL, then bench at N GPUs through torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   scripts/dist_check.py 2>&1 | grep -E "rank|Error|error" | tee gpurun_out/dist_check_n$N.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json

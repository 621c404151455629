#!/bin/bash
# multi-GPU session: sharded commit/open parity over NCCL and over the peer-memory exchange, then
# bench at N GPUs through torchrun (peer exchange; NCCL route beside it when $2 = both)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   scripts/dist_check.py 2>&1 | grep -E "rank|Error|error" | tee gpurun_out/dist_check_n$N.txt
if [ "$2" = "both" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus $N --steps 10 --warmup 3 --exchange nccl > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
grep -o '"ms_per_step": [0-9.]*\|"exchange_ms": [0-9.]*' gpurun_out/bench_n${N}_nccl.json
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json

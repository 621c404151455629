#!/bin/bash
# e2e of the range-sharded 2^24 commit at N GPUs for several upload chunkings (count, growth ratio); development aid
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/upload_ratio_n$N.jsonl
port=29520
for v in "4 1" "3 2" "4 2" "4 1.5" "5 1.5" "2 3"; do
  set -- $v
  port=$((port+1))
  MZ_UPLOAD_CHUNKS=$1 MZ_UPLOAD_RATIO=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-verify --no-extras --skip-configs > gpurun_out/_u.json 2> gpurun_out/_u.err
  python - "$1" "$2" <<'PY' >> gpurun_out/upload_ratio_n$N.jsonl
import json, sys
try:
    d = json.loads(open("gpurun_out/_u.json").read().strip().splitlines()[-1])
    print(json.dumps({"chunks": sys.argv[1], "ratio": sys.argv[2], "ms": d["ms_per_step"], "e2e_ms": d["e2e"]["ms_per_step"]}))
except Exception as e:
    print(json.dumps({"chunks": sys.argv[1], "ratio": sys.argv[2], "error": str(e)}))
PY
done
cat gpurun_out/upload_ratio_n$N.jsonl

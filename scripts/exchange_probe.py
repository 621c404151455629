"""Run under torchrun: latency of the peer-memory exchange kernel against NCCL all_gather + sum (development aid)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myzkp_b200 as mz
from myzkp_b200 import synth
from myzkp_b200.dist import DeviceOps, ShardedKZG, shard_range

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = mz.Context(local)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
lo, hi = shard_range(n, rank, world)
ctx.srs_generate(synth.random_scalar(synth.SEED_ALPHA), hi - lo, first=lo)
coefs = synth.random_scalars(n, 99)
d = torch.from_numpy(coefs[lo:hi].view(np.int64).reshape(-1).copy()).to(dev)
out = torch.zeros(64, dtype=torch.uint8, device=dev)
ops = DeviceOps(ctx, dev)
ok = ops.attach_peers(rank, world)
prover = ShardedKZG(ops, rank, world, n)
res = {"world": world, "n": n, "attached": ok}


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ops.msm_partial(d.data_ptr(), hi - lo)
res["exchange_kernel_only_ms"] = timed(lambda: ops.exchange_sum(ops.partial, out), 200)
gathered = torch.zeros(world * 128, dtype=torch.uint8, device=dev)


def nccl_route():
    dist.all_gather_into_tensor(gathered, ops.partial)
    ops.sum_partials(gathered, world, out)


res["nccl_gather_sum_ms"] = timed(nccl_route, 200)
res["local_msm_ms"] = timed(lambda: ops.msm_partial(d.data_ptr(), hi - lo), 20)
res["commit_peer_ms"] = timed(lambda: prover.commit(d.data_ptr(), out), 20)
ops.fused = False
res["commit_nccl_ms"] = timed(lambda: prover.commit(d.data_ptr(), out), 20)
ops.fused = ok
ctx.sync()
if rank == 0:
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"exchange_probe_n{world}.json"), "w"))
dist.barrier()
ctx.peer_detach()
dist.destroy_process_group()

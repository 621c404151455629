"""BASELINE config C3: KZG commit + open of a degree-2^20 polynomial, range-sharded over the ranks of a torchrun
launch (or one GPU when run directly).  Device-resident coefficients, peer-memory exchange, CUDA events, max over
ranks; results checked against the oracle's expected values.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import myzkp_b200 as mz
from myzkp_b200 import synth
from myzkp_b200.dist import DeviceOps, ShardedKZG, shard_range
import myzkp_oracle as orc  # checker

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = 10
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = mz.Context(local)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n = 1 << log2n
alpha, u = synth.random_scalar(synth.SEED_ALPHA), synth.random_scalar(synth.SEED_OPEN)
coefs = synth.random_scalars(n, synth.SEED_SCALARS + log2n)
lo, hi = shard_range(n, rank, world)
ctx.srs_generate(alpha, hi - lo, first=lo)
d = torch.from_numpy(coefs[lo:hi].view(np.int64).reshape(-1).copy()).to(dev)
ops = DeviceOps(ctx, dev)
fused = world > 1 and ops.attach_peers(rank, world)
prover = ShardedKZG(ops, rank, world, n)
out = torch.zeros(64, dtype=torch.uint8, device=dev)
y = torch.zeros(32, dtype=torch.uint8, device=dev)
w = torch.zeros(64, dtype=torch.uint8, device=dev)
scratch = torch.zeros(64, dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn):
    for _ in range(3):
        fn()
    barrier()
    if fused:
        ops.exchange_sum(ops.partial, scratch)  # device-side rendezvous: the timed regions start together
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms_commit = timed(lambda: prover.commit(d.data_ptr(), out))
ms_open = timed(lambda: prover.open(d.data_ptr(), u, y, w))
barrier()
ints = synth.limbs_to_ints(coefs) if rank == 0 else None
if rank == 0:
    fa = synth.horner_mod_r(coefs, alpha)
    c = mz.context.point_from_bytes(out.cpu().numpy().tobytes())
    yy = int.from_bytes(y.cpu().numpy().tobytes(), "little")
    ww = mz.context.point_from_bytes(w.cpu().numpy().tobytes())
    ok = c == orc.fast_mul(fa) and (yy, ww) == orc.expected_open(ints, u, alpha)
    print(json.dumps({"config": f"C3: KZG commit+open of a degree-2^{log2n} polynomial", "n_gpus": world,
                      "exchange": "peer memory kernel" if fused else ("nccl all_gather" if world > 1 else "none"),
                      "commit_ms": round(ms_commit, 4), "open_ms": round(ms_open, 4),
                      "commit_plus_open_ms": round(ms_commit + ms_open, 4),
                      "commits_per_sec": round(1e3 / ms_commit, 1), "verified_vs_oracle": bool(ok)}), flush=True)
if world > 1:
    ctx.peer_detach()
    dist.barrier()
    dist.destroy_process_group()

"""Run under torchrun: sharded commit + open over NCCL on real GPUs, checked against the oracle's expected values."""
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

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = mz.Context(local)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ok = True
for n in (1 << 16, (1 << 14) + 3, 5):
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    u = synth.random_scalar(synth.SEED_OPEN)
    coefs = synth.random_scalars(n, 1234 + n)
    lo, hi = shard_range(n, rank, world)
    ctx.srs_generate(alpha, hi - lo, first=lo)
    d = torch.from_numpy(coefs[lo:hi].view(np.int64).reshape(-1).copy()).to(dev) if hi > lo else torch.zeros(4, dtype=torch.int64, device=dev)
    ints = synth.limbs_to_ints(coefs)
    exp_c, exp_o = orc.expected_commit(ints, alpha), orc.expected_open(ints, u, alpha)
    for route in ("nccl", "peer"):
        ops = DeviceOps(ctx, dev)
        if route == "peer" and not ops.attach_peers(rank, world):
            print(f"rank {rank}/{world} n={n}: peer attach FAILED", flush=True)
            ok = False
            continue
        prover = ShardedKZG(ops, rank, world, n)
        for rep in range(2):
            out = torch.zeros(64, dtype=torch.uint8, device=dev)
            y = torch.zeros(32, dtype=torch.uint8, device=dev)
            w = torch.zeros(64, dtype=torch.uint8, device=dev)
            prover.commit(d.data_ptr(), out)
            prover.open(d.data_ptr(), u, y, w)
            ctx.sync()
            torch.cuda.synchronize()
            c = mz.context.point_from_bytes(out.cpu().numpy().tobytes())
            yy = int.from_bytes(y.cpu().numpy().tobytes(), "little")
            ww = mz.context.point_from_bytes(w.cpu().numpy().tobytes())
            good = c == exp_c and (yy, ww) == exp_o
            if route == "peer" and hi > lo:
                good = good and ctx.commit_sharded(coefs[lo:hi]) == exp_c
            elif route == "peer":
                good = good and ctx.commit_sharded(coefs[:0]) == exp_c
            ok = ok and good
        print(f"rank {rank}/{world} n={n}: sharded commit+open via {route} {'OK' if good else 'MISMATCH'}", flush=True)
    ctx.peer_detach()
    dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

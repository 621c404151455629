"""Phase timings of device-resident commits (recode / sort / accumulate / merge / reduce) at given sizes.
usage: phase_probe.py LOG2N[:BAA] ...   (env knobs: MZ_NO_PARTITION, MZ_BAA_MINB)"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import myzkp_b200 as mz
from myzkp_b200 import synth

ctx = mz.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.enable_phase_timing(True)
alpha = synth.random_scalar(synth.SEED_ALPHA)
out = torch.zeros(64, dtype=torch.uint8, device="cuda")
last_n = None
for item in sys.argv[1:]:
    lg, _, baa = item.partition(":")
    lg = int(lg)
    n = 1 << lg
    if n != last_n:
        ctx.srs_generate(alpha, n)
        coefs = torch.from_numpy(synth.random_scalars(n, synth.SEED_SCALARS + lg).view(np.int64).reshape(-1)).cuda()
        last_n = n
    ctx.set_baa_rounds(int(baa) if baa else -1)
    for _ in range(3):
        ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 8
    e0.record()
    for _ in range(reps):
        ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    acc = {}
    for back in range(reps):
        ph, info = ctx.msm_phases(back)
        for k, v in ph.items():
            acc[k] = acc.get(k, 0.0) + v / reps
    print(json.dumps({"log2n": lg, "baa": baa or "auto", "c": info["window_bits"], "L": info["segment_len"],
                      "total_ms": round(e0.elapsed_time(e1) / reps, 3), **{k: round(v, 3) for k, v in acc.items()},
                      "point": out.cpu().numpy().tobytes().hex()[:16]}), flush=True)

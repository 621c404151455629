"""Accumulate-phase timing with 0..R batched-affine rounds (development aid)."""
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
for item in sys.argv[1:]:
    lg, rs = item.split(":")
    lg = int(lg)
    n = 1 << lg
    ctx.srs_generate(alpha, n)
    coefs = torch.from_numpy(synth.random_scalars(n, synth.SEED_SCALARS + lg).view(np.int64).reshape(-1)).cuda()
    ref = None
    for r in [int(x) for x in rs.split(",")]:
        ctx.set_baa_rounds(r)
        for _ in range(3):
            ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
        torch.cuda.synchronize()
        got = bytes(out.cpu().numpy().tobytes())
        ref = ref or got
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ph, info = ctx.msm_phases(0)
        print(json.dumps({"log2n": lg, "baa_rounds": r, "c": info["window_bits"], "L": info["segment_len"], "same_point": got == ref,
                          "total_ms": round(e0.elapsed_time(e1) / 5, 3), "accumulate_ms": round(ph["accumulate"], 3)}), flush=True)

"""Per-phase timing of the device-resident commit over sizes and window bits (development aid)."""
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
res = []
spec = sys.argv[1:] or ["16:0,12,16", "20:0,16,20", "22:0,16,20,24"]
for item in spec:
    parts = item.split(":")  # log2n : window bits list [: segment length list]
    lg, wbs = parts[0], parts[1]
    segs = [int(x) for x in parts[2].split(",")] if len(parts) > 2 else [0]
    lg = int(lg)
    n = 1 << lg
    ctx.srs_generate(alpha, n)
    coefs = torch.from_numpy(synth.random_scalars(n, synth.SEED_SCALARS + lg).view(np.int64).reshape(-1)).cuda()
    for wb in [int(x) for x in wbs.split(",")]:
        for seg in segs:
            ctx.set_msm_params(wb, seg)
            for _ in range(3):
                ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ph, info = ctx.msm_phases(0)
            row = {"log2n": lg, "c": info["window_bits"], "L": info["segment_len"], "segs": info["segments"],
                   "total_ms": round(e0.elapsed_time(e1) / reps, 3), **{k: round(v, 3) for k, v in ph.items()},
                   "out": bytes(out.cpu().numpy())[:8].hex()}
            res.append(row)
            print(json.dumps(row), flush=True)
json.dump(res, open("gpurun_out/phase_sweep.json", "w"), indent=1)

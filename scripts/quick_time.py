"""Rough per-size timing of the device-resident commit (development aid, not the bench)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import myzkp_b200 as mz
from myzkp_b200 import synth

torch.cuda.init()
ctx = mz.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
alpha = synth.random_scalar(synth.SEED_ALPHA)
res = {}
sizes = [int(a) for a in sys.argv[1:]] or [16, 20]
for logn in sizes:
    n = 1 << logn
    t0 = time.perf_counter()
    ctx.srs_generate(alpha, n)
    t_srs = time.perf_counter() - t0
    coefs = torch.from_numpy(synth.random_scalars(n, synth.SEED_SCALARS + logn).view(np.int64)).cuda()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for wb in (0, 8, 16, 24):
        if wb == 8 and logn > 18:
            continue
        ctx.set_msm_params(wb, 0)
        for _ in range(2):
            ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[f"2^{logn}_c{wb}"] = {"ms": round(ms, 3), "Mpts_s": round(n / ms / 1e3, 2)}
    res[f"2^{logn}_srs_s"] = round(t_srs, 3)
    print(json.dumps(res), flush=True)
json.dump(res, open("gpurun_out/quick_time.json", "w"), indent=1)

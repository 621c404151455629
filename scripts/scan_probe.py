"""Quotient scan alone on device-resident coefficients (timing with CUDA events; also the ncu target)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import myzkp_b200 as mz
from myzkp_b200 import synth

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 1 << lg
ctx = mz.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
coefs = torch.from_numpy(synth.random_scalars(n, synth.SEED_SCALARS + lg).view(np.int64).reshape(-1)).cuda()
q = torch.zeros(n * 4, dtype=torch.int64, device="cuda")
c0 = torch.zeros(16, dtype=torch.int64, device="cuda")
u = synth.random_scalar(synth.SEED_OPEN)
for mode in ("quotient", "eval"):
    def run():
        if mode == "quotient":
            ctx.fr_range_quotient_dev(coefs.data_ptr(), n, u, 0, q.data_ptr(), c0.data_ptr())
        else:
            ctx.fr_range_eval_dev(coefs.data_ptr(), n, u, c0.data_ptr(), c0.data_ptr() + 32)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'{{"scan": "{mode}", "log2n": {lg}, "ms": {ms:.4f}, "GBps_algorithmic": {n * (64 if mode == "quotient" else 32) / ms / 1e6:.1f}}}')
ctx.sync()

"""One device-resident commit at a given size / window / BAA rounds (for ncu captures)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import myzkp_b200 as mz
from myzkp_b200 import synth

lg, wb = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
baa = int(sys.argv[4]) if len(sys.argv) > 4 else -1
n = 1 << lg
ctx = mz.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.srs_generate(synth.random_scalar(synth.SEED_ALPHA), n)
coefs = torch.from_numpy(synth.random_scalars(n, synth.SEED_SCALARS + lg).view(np.int64).reshape(-1)).cuda()
out = torch.zeros(64, dtype=torch.uint8, device="cuda")
ctx.set_msm_params(wb, 0)
ctx.set_baa_rounds(baa)
for _ in range(reps):
    ctx.commit_dev(coefs.data_ptr(), n, out.data_ptr())
torch.cuda.synchronize()
print("done")

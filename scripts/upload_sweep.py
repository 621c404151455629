"""Host-API commit (pinned input, H2D inside) over sizes and upload-chunk counts (development aid)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import myzkp_b200 as mz
from myzkp_b200 import synth

ctx = mz.Context(0)
alpha = synth.random_scalar(synth.SEED_ALPHA)
res = []
for lg in [int(x) for x in (sys.argv[1:] or ["19", "20", "21", "22"])]:
    n = 1 << lg
    ctx.srs_generate(alpha, n)
    pinned = ctx.host_alloc(n * 32)
    pinned[:] = synth.random_scalars(n, synth.SEED_SCALARS + lg).view(np.uint8).reshape(-1)
    coefs = pinned.reshape(n, 32)
    import torch
    d = torch.from_numpy(np.ascontiguousarray(coefs).view(np.int64).reshape(-1).copy()).cuda()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.commit_dev(d.data_ptr(), n, out.data_ptr())
    ctx.sync()
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter()
        ctx.commit_dev(d.data_ptr(), n, out.data_ptr())
        ctx.sync()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    print(json.dumps({"log2n": lg, "resident_ms": round(best, 3), "out": bytes(out.cpu().numpy())[:8].hex()}), flush=True)
    del d
    for k in (1, 2, 3, 4):
        ctx.set_upload_chunks(k)
        for _ in range(3):
            ctx.commit(coefs)
        best = 1e9
        for _ in range(7):
            t0 = time.perf_counter()
            ctx.commit(coefs)
            best = min(best, (time.perf_counter() - t0) * 1e3)
        c = ctx.commit(coefs)
        res.append({"log2n": lg, "chunks": k, "ms": round(best, 3), "out": str(c)[:24]})
        print(json.dumps(res[-1]), flush=True)
    ctx.host_free(pinned)

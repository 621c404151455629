"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of the path at small sizes.
usage: compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_small.py [gemini_log2n]"""
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # same-device emulated ranks: no lazy loads while a peer spins

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import myzkp_b200 as mz
import myzkp_oracle as o
from myzkp_b200 import synth
from myzkp_b200.dist import shard_range

R = o.R_MOD
ctx = mz.Context(0)
alpha, u = 123456789, 5
# window / segment / accumulate variants: XYZZ, multi-pass BAA rounds, fused pair sums; partitioned sort (c >= 20)
for n, c, seg, baa in ((300, 0, 0, -1), (1500, 12, 3, -1), (5000, 16, 0, 2), (70000, 20, 0, -1), (70000, 22, 64, -2),
                       (9000, 20, 300, -2), (3000, 24, 0, 0)):
    ctx.srs_generate(alpha, n)
    coefs = [(i * 7919 + 13) % R for i in range(n)]
    ctx.set_msm_params(c, seg)
    ctx.set_baa_rounds(baa)
    assert ctx.commit(coefs) == o.expected_commit(coefs, alpha)
    assert ctx.open(coefs, u) == o.expected_open(coefs, u, alpha)
# heavy buckets: long head chains -> several merge levels
ctx.srs_generate(alpha, 70000)
ctx.set_msm_params(8, 4)
ctx.set_baa_rounds(0)
ones = [1] * 70000
assert ctx.commit(ones) == o.expected_commit(ones, alpha)
ctx.set_msm_params(0, 0)
ctx.set_baa_rounds(-1)
# MSD sort of the large windows: group-local pass, oversize groups (capacity lowered), sentinel partition (zeros);
# upload pipeline in 3 chunks with ONE deferred head merge
ctx.srs_generate(alpha, 20000)
skew = [((i % 7) + 1) if i % 3 else 0 for i in range(20000)]
for cap in (64, 0):
    ctx._lib.myzkp_test_set_sort_group_cap(cap)
    for c in (20, 22):
        ctx.set_msm_params(c, 0)
        assert ctx.commit(skew) == o.expected_commit(skew, alpha)
ctx.set_msm_params(20, 0)
ctx.set_upload_chunks(3)
coefs = [(i * 104729 + 1) % R for i in range(20000)]
assert ctx.commit(coefs) == o.expected_commit(coefs, alpha)
assert ctx.open(coefs, u) == o.expected_open(coefs, u, alpha)
ctx.set_upload_chunks(0)
ctx.set_msm_params(0, 0)
# batch of small polynomials in one pipeline, Gemini (single-stream form), batch open, degree bound
polys = [[(i * 31 + j) % R for i in range(200 + 37 * j)] for j in range(9)]
got = ctx.commit_batch(polys)
assert got == [o.expected_commit(p, alpha) for p in polys]
ctx.srs_generate(alpha, 16)
print(ctx.gemini_fold_commit(list(range(1, 17)), [2, 3, 4, 5])[:1])
print(ctx.batch_open(list(range(1, 17)), [7, 8, 9])[0])
print(ctx.prove_degree_bound(list(range(1, 9)), 8) is not None)
# caller-supplied points: windowed MSM without a table
pts = [o.fast_mul(k + 2) for k in range(300)] + [None]
sc = [(k * 977 + 5) % R for k in range(301)]
exp = None
for k, p_ in zip(sc, pts):
    exp = o._fast_add(exp, o.fast_mul(k, p_) if p_ else None)
assert ctx.g1_msm(sc, pts) == exp
# Gemini with the concurrent child-stream batch (levels below 2^19 next to the large ones)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if lg:
    n = 1 << lg
    ctx.srs_generate(alpha, n)
    coefs = synth.random_scalars(n, 5)
    rhos = [synth.random_scalar(100 + i) for i in range(lg)]
    pts = ctx.gemini_fold_commit(coefs, rhos)
    assert pts[0] == o.expected_commit(synth.limbs_to_ints(coefs), alpha)
    print("gemini", lg, "ok")
# two emulated ranks in this process: range-sharded commit + open with the exchange over peer memory.
# MZ_SANITIZE_NO_PEERS=1 skips it: under memcheck the two ranks' kernels are not always run concurrently (the tool may
# serialise launches), and a rank that spins for its peer then runs into the exchange time-out; racecheck and
# synccheck run this section.
if os.environ.get("MZ_SANITIZE_NO_PEERS"):
    print("sanitize run ok (peer section skipped)")
    sys.exit(0)
import torch

n, world = 3001, 2
coefs = synth.random_scalars(n, 71)
ints = synth.limbs_to_ints(coefs)
d_all = torch.from_numpy(coefs.view(np.int64).reshape(-1).copy()).cuda()
ranks = []
for r in range(world):
    c = mz.Context(0)
    lo, hi = shard_range(n, r, world)
    c.srs_generate(alpha, hi - lo, first=lo)
    c.peer_export()
    c.peer_set_timeout_ms(60000)
    ranks.append((c, lo, hi))
ctxs = [c for c, _, _ in ranks]
outs = torch.zeros(world, 64, dtype=torch.uint8, device="cuda")
ys = torch.zeros(world, 32, dtype=torch.uint8, device="cuda")
ws = torch.zeros(world, 64, dtype=torch.uint8, device="cuda")
for r, (c, lo, hi) in enumerate(ranks):
    c.reserve(hi - lo)
    c.peer_attach_local(r, ctxs)
for r, (c, lo, hi) in enumerate(ranks):
    c.commit_sharded_dev(d_all.data_ptr() + lo * 32, hi - lo, outs[r].data_ptr())
    c.open_sharded_dev(d_all.data_ptr() + lo * 32, hi - lo, u, ys[r].data_ptr(), ws[r].data_ptr())
for c in ctxs:
    c.sync()
for r in range(world):
    assert mz.context.point_from_bytes(outs[r].cpu().numpy().tobytes()) == o.expected_commit(ints, alpha)
    assert (int.from_bytes(ys[r].cpu().numpy().tobytes(), "little"),
            mz.context.point_from_bytes(ws[r].cpu().numpy().tobytes())) == o.expected_open(ints, u, alpha)
for c in ctxs:
    c.close()
print("sanitize run ok")

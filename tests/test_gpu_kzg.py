"""GPU parity: setup / commit / open / Gemini through the C ABI vs the oracle and the golden vectors."""
import json
import os
import random

import numpy as np
import pytest

import myzkp_oracle as o
import oracle_c as oc
from myzkp_oracle import Fr

import myzkp_b200 as mz
from myzkp_b200 import synth

pytestmark = pytest.mark.gpu
R = o.R_MOD
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kzg_golden.json")))


def _pt(p):
    return None if p is None else (int(p[0]), int(p[1]))


def test_golden_kzg_vectors(ctx):
    """setup -> commit -> open against vectors produced by the faithful oracle."""
    for case in G["kzg"]:
        alpha, coefs, u = int(case["alpha"]), [int(c) for c in case["coefs"]], int(case["u"])
        pk = mz.setup_kzg(mz.BN128.generator_g1(), None, len(coefs) - 1, alpha=alpha, ctx=ctx)
        assert [p.as_tuple() for p in pk.powers_1] == [_pt(p) for p in case["srs"]], case["name"]
        c = mz.commit_kzg(mz.Polynomial(coefs), pk)
        assert c.as_tuple() == _pt(case["commit"]), case["name"]
        pr = mz.open_kzg(mz.Polynomial(coefs), u, pk)
        assert pr.y == int(case["y"]) and pr.w.as_tuple() == _pt(case["w"]), case["name"]


def test_golden_gemini_vectors(ctx):
    for case in G["gemini"]:
        coefs, rhos, alpha = [int(c) for c in case["coefs"]], [int(r) for r in case["rhos"]], int(case["alpha"])
        pk = mz.setup_kzg(mz.BN128.generator_g1(), None, len(coefs) - 1, alpha=alpha, ctx=ctx)
        cms, polys = mz.split_and_fold_commit(coefs, rhos, pk, want_folds=True)
        assert [c.as_tuple() for c in cms] == [_pt(p) for p in case["commitments"]], case["name"]
        assert [list(p.coef) for p in polys] == [[int(v) for v in f] for f in case["folds"][1:]]
        # commit_gemini on explicit polynomials (gemini.rs:112-114)
        all_polys = [mz.Polynomial(coefs)] + polys
        assert [c.as_tuple() for c in mz.commit_gemini(all_polys, pk)] == [_pt(p) for p in case["commitments"]]
    with pytest.raises(mz.SplitFoldError):
        mz.split_and_fold_commit([1, 2, 3], [1], pk)
    with pytest.raises(mz.SplitFoldError):
        mz.split_and_fold_commit([1, 2, 3, 4], [1], pk)


def test_config1_2pow10_vs_faithful_oracle_sample(ctx):
    """BASELINE config 1 (degree 2^10): SRS spot-checked against the faithful scalar-mul,
    commit/open against the algebraic expected value (validated vs the faithful path in the CPU suite)."""
    n = 1 << 10
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    u = synth.random_scalar(synth.SEED_OPEN)
    coefs = synth.random_scalars(n, synth.SEED_SCALARS + 10)
    ints = synth.limbs_to_ints(coefs)
    pk = mz.setup_kzg(mz.BN128.generator_g1(), None, n - 1, alpha=alpha, ctx=ctx)
    assert len(pk) == n
    g = o.generator_g1()
    for i in (0, 1, 2, 513, n - 1):
        assert ctx.srs_read(i, 1)[0] == (g * pow(alpha, i, R)).affine_ints()
    srs = ctx.srs_read(0, n)
    assert srs[:64] == [o.fast_mul(pow(alpha, i, R)) for i in range(64)]
    assert mz.commit_kzg(mz.Polynomial(coefs), pk).as_tuple() == o.expected_commit(ints, alpha)
    pr = mz.open_kzg(mz.Polynomial(coefs), u, pk)
    assert (pr.y, pr.w.as_tuple()) == o.expected_open(ints, u, alpha)
    # faithful naive MSM on a 48-coefficient prefix (seconds on the CPU)
    pk_small = o.PublicKeyKZG([o.G1Point.new(o.Fq(x), o.Fq(y)) for x, y in srs[:48]])
    f_small = o.Polynomial([Fr(c) for c in ints[:48]])
    assert mz.commit_kzg(mz.Polynomial(ints[:48]), pk).as_tuple() == o.commit_kzg(f_small, pk_small).affine_ints()


@pytest.mark.parametrize("logn", [16, 20])
def test_commit_open_large_vs_expected_value(ctx, logn):
    """BASELINE configs 2 and 3 on one GPU."""
    n = 1 << logn
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    u = synth.random_scalar(synth.SEED_OPEN)
    coefs = synth.random_scalars(n, synth.SEED_SCALARS + logn)
    ints = synth.limbs_to_ints(coefs)
    ctx.srs_generate(alpha, n)
    for i in (0, 1, n // 3, n - 1):
        assert ctx.srs_read(i, 1)[0] == o.fast_mul(pow(alpha, i, R))
    assert ctx.commit(coefs) == o.expected_commit(ints, alpha)
    y, w = ctx.open(coefs, u)
    assert (y, w) == o.expected_open(ints, u, alpha)
    # quotient coefficients themselves
    yq, q = ctx.fr_quotient(coefs, u)
    ey, eq = o.synthetic_division(ints, u)
    assert yq == ey
    assert synth.limbs_to_ints(q.view(np.uint64).reshape(-1, 4)) == eq
    assert ctx.fr_eval(coefs, u) == ey


def test_window_sizes_and_segments_agree(ctx):
    n = 5000  # not a power of two
    alpha = 0xABCDEF123456789
    coefs = synth.random_scalars(n, 99)
    ints = synth.limbs_to_ints(coefs)
    ctx.srs_generate(alpha, n)
    exp = o.expected_commit(ints, alpha)
    try:
        for c in (4, 8, 12, 16, 20, 22, 24):
            for seg in (0, 1, 7, 1000):
                ctx.set_msm_params(c, seg)
                assert ctx.commit(coefs) == exp, (c, seg)
    finally:
        ctx.set_msm_params(0, 0)


def test_scalar_distributions(ctx):
    """byte-valued scalars (das/avail.rs:93, das/eigenda.rs:96), zeros, all-equal, r-1, P/-P cancellation."""
    n = 1 << 12
    alpha = 31337
    ctx.srs_generate(alpha, n)
    rnd = random.Random(5)
    dists = {
        "bytes": [rnd.randrange(256) for _ in range(n)],
        "zeros10": [0 if rnd.random() < 0.1 else rnd.randrange(R) for _ in range(n)],
        "all_zero": [0] * n,
        "all_one": [1] * n,
        "all_rm1": [R - 1] * n,
        "single": [0] * (n - 1) + [5],
        "two_pow": [1 << rnd.randrange(254) for _ in range(n)],
        "half_windows": [sum((1 << 15) << (16 * w) for w in range(15))] * n,
    }
    for name, sc in dists.items():
        exp = o.expected_commit(sc, alpha)
        for c in (4, 8, 12, 16, 20, 22, 24):
            ctx.set_msm_params(c, 0)
            assert ctx.commit(sc) == exp, (name, c)
    ctx.set_msm_params(0, 0)
    # SRS with repeated points, P / -P pairs and infinity: load explicit points
    base = [o.fast_mul(k) for k in (3, 3, 5, 7)]
    neg = lambda p: (p[0], o.P_MOD - p[1])
    pts = [base[0], base[1], neg(base[0]), None, base[2], neg(base[2]), base[3], base[3]]
    ctx.srs_load(pts)
    assert ctx.srs_read(0, 8) == pts
    sc = [2, 2, 4, 99, 6, 6, 1, R - 1]  # 2*3G + 2*3G - 4*3G + 0 + 6*5G - 6*5G + 7G - 7G = inf
    assert ctx.commit(sc) is None
    sc = [1, 1, 0, 5, 0, 0, 0, 0]
    assert ctx.commit(sc) == o.fast_mul(6)  # P + P must double


def test_msd_sort_group_paths(ctx):
    """The large windows sort by partition -> 256-way pass -> group-local shared-memory sort (sort.cu).  With the group
    capacity lowered, small inputs take the paths full-size inputs take: several groups per block, oversize groups
    (skewed scalars) through the generic pass restricted to them, the sentinel partition (zero digits), empty groups."""
    lib = ctx._lib
    n = 1 << 14
    alpha = 777
    ctx.srs_generate(alpha, n)
    rnd = random.Random(11)
    dists = {
        "uniform": [rnd.randrange(R) for _ in range(n)],
        "bytes": [rnd.randrange(256) for _ in range(n)],
        "zeros50": [0 if rnd.random() < 0.5 else rnd.randrange(R) for _ in range(n)],
        "all_one": [1] * n,
        "few_values": [rnd.choice([3, R - 2, 1 << 200, (1 << 253) + 12345]) for _ in range(n)],
        "two_pow": [1 << rnd.randrange(254) for _ in range(n)],
    }
    exp = {name: o.expected_commit(sc, alpha) for name, sc in dists.items()}
    try:
        for cap in (0, 4096, 600, 64):
            assert lib.myzkp_test_set_sort_group_cap(cap) == 0
            for name, sc in dists.items():
                for c in (20, 22, 24):
                    ctx.set_msm_params(c, 0)
                    assert ctx.commit(sc) == exp[name], (cap, name, c)
    finally:
        lib.myzkp_test_set_sort_group_cap(0)
        ctx.set_msm_params(0, 0)


def test_edge_cases_and_errors(ctx):
    alpha = 4242
    ctx.srs_generate(alpha, 8)
    assert ctx.commit([]) is None  # empty polynomial -> infinity
    assert ctx.commit([0, 0, 0]) is None
    y, w = ctx.open([9], 5)  # constant: W = infinity, y = f_0 (polynomial.rs:372-374)
    assert (y, w) == (9, None)
    assert ctx.open([], 5) == (0, None)
    y, w = ctx.open([3, 0, 0, 1], 0)  # u = 0
    assert (y, w) == o.expected_open([3, 0, 0, 1], 0, alpha)
    coefs = [1, 2, 3, 4, 5, 6, 7, 8]
    assert ctx.open(coefs, alpha) [0] == o.synthetic_division(coefs, alpha)[0]  # u == alpha still opens
    with pytest.raises(RuntimeError):  # longer than the SRS: reference panics (polynomial.rs:162)
        ctx.commit(list(range(9)))
    with pytest.raises(RuntimeError):  # non-canonical scalar
        ctx.commit(np.frombuffer(int(R).to_bytes(32, "little"), dtype=np.uint8).reshape(1, 32).copy())
    with pytest.raises(RuntimeError):
        ctx.srs_load([(o.P_MOD, 2)])
    ctx.srs_generate(alpha, 8)
    assert ctx.commit([R - 1]) == (1, o.P_MOD - 2)


def test_linearity_and_shift_properties_full_size(ctx):
    """Size-independent properties at 2^20: commit(a) + commit(b) == commit(a+b) via the device group law,
    and commit of x*f against SRS == shifted MSM."""
    n = 1 << 18
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    ctx.srs_generate(alpha, n)
    a = synth.random_scalars(n, 1)
    b = synth.random_scalars(n, 2)
    ai, bi = synth.limbs_to_ints(a), synth.limbs_to_ints(b)
    s = synth.ints_to_limbs([(x + y) % R for x, y in zip(ai, bi)])
    ca, cb, cs = ctx.commit(a), ctx.commit(b), ctx.commit(s)
    assert ctx.test_g1_op(0, [ca], [cb])[0] == cs
    # opening identity: C - [y]G == [alpha - u] W   <=>  checked through scalars: f(alpha) - y == (alpha-u) q(alpha)
    u = 77
    y, w = ctx.open(a, u)
    fa = synth.horner_mod_r(a, alpha)
    k = (fa - y) * pow((alpha - u) % R, -1, R) % R
    assert w == o.fast_mul(k)


def test_sharded_commit_and_open_emulated_ranks(ctx):
    """The multi-GPU path on one GPU: one Context per emulated rank, each with its SRS range
    (srs_generate first=lo), partial XYZZ per rank, sum_partials; open through range_eval /
    compose_carries / range_quotient - everything but the NCCL all-gather itself."""
    import torch

    from myzkp_b200.dist import compose_carries, shard_range

    n, world = 3001, 3
    alpha, u = 0x1234567890ABCDEF1234567, 0xFEDCBA987654321
    coefs = synth.random_scalars(n, 7)
    ints = synth.limbs_to_ints(coefs)
    dev = torch.device("cuda", 0)
    d_all = torch.from_numpy(coefs.view(np.int64).reshape(-1)).to(dev)
    ranks = []
    for r in range(world):
        c = mz.Context(0)
        lo, hi = shard_range(n, r, world)
        c.srs_generate(alpha, hi - lo, first=lo)
        assert c.srs_read(0, 1)[0] == o.fast_mul(pow(alpha, lo, R))
        ranks.append((c, lo, hi))
    partials = torch.zeros(world * 128, dtype=torch.uint8, device=dev)
    pairs = torch.zeros(world * 64, dtype=torch.uint8, device=dev)
    out = torch.zeros(64, dtype=torch.uint8, device=dev)
    for r, (c, lo, hi) in enumerate(ranks):
        c.msm_partial_dev(d_all.data_ptr() + lo * 32, hi - lo, 0, partials.data_ptr() + 128 * r)
        c.fr_range_eval_dev(d_all.data_ptr() + lo * 32, hi - lo, u, pairs.data_ptr() + 64 * r, pairs.data_ptr() + 64 * r + 32)
        c.sync()
    ranks[0][0].sum_partials_dev(partials.data_ptr(), world, out.data_ptr())
    ranks[0][0].sync()
    assert mz.context.point_from_bytes(out.cpu().numpy().tobytes()) == o.expected_commit(ints, alpha)
    # the same partials from host scalars through the chunked upload pipeline
    for r, (c, lo, hi) in enumerate(ranks):
        c.set_upload_chunks(3)
        c.msm_partial_host(coefs[lo:hi], 0, partials.data_ptr() + 128 * r)
        c.sync()
        c.set_upload_chunks(0)
    ranks[2][0].sum_partials_dev(partials.data_ptr(), world, out.data_ptr())
    ranks[2][0].sync()
    assert mz.context.point_from_bytes(out.cpu().numpy().tobytes()) == o.expected_commit(ints, alpha)
    raw = pairs.cpu().numpy().tobytes()
    hs = [int.from_bytes(raw[64 * r : 64 * r + 32], "little") for r in range(world)]
    ms = [int.from_bytes(raw[64 * r + 32 : 64 * r + 64], "little") for r in range(world)]
    for r, (c, lo, hi) in enumerate(ranks):
        assert ms[r] == pow(u, hi - lo, R)
    carries = compose_carries(hs, ms)
    c0s = torch.zeros(world * 32, dtype=torch.uint8, device=dev)
    qs = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
    for r, (c, lo, hi) in enumerate(ranks):
        c.fr_range_quotient_dev(d_all.data_ptr() + lo * 32, hi - lo, u, carries[r], qs.data_ptr() + lo * 32, c0s.data_ptr() + 32 * r)
        c.msm_partial_dev(qs.data_ptr() + lo * 32, hi - lo, 0, partials.data_ptr() + 128 * r)
        c.sync()
    ranks[1][0].sum_partials_dev(partials.data_ptr(), world, out.data_ptr())
    ranks[1][0].sync()
    ey, ew = o.expected_open(ints, u, alpha)
    assert int.from_bytes(c0s.cpu().numpy().tobytes()[:32], "little") == ey
    assert mz.context.point_from_bytes(out.cpu().numpy().tobytes()) == ew
    for c, _, _ in ranks:
        c.close()


@pytest.mark.parametrize("n,world", [(3001, 3), (4, 5), (777, 1), (20000, 8)])
def test_sharded_fused_peer_exchange_emulated_ranks(ctx, n, world):
    """commit / open with the exchange fused over peer memory (csrc/peer.cu): one Context (own
    stream) per emulated rank in this process, buffers attached by pointer; every rank must end
    with the oracle's commitment and (y, W).  Run twice so both epoch parities are exercised."""
    import torch

    from myzkp_b200.dist import shard_range

    alpha, u = 0xABCDEF0123456789ABCDEF, 0x1122334455667788
    coefs = synth.random_scalars(n, 70 + n)
    ints = synth.limbs_to_ints(coefs)
    dev = torch.device("cuda", 0)
    d_all = torch.from_numpy(coefs.view(np.int64).reshape(-1).copy()).to(dev)
    ranks = []
    for r in range(world):
        c = mz.Context(0)
        lo, hi = shard_range(n, r, world)
        c.srs_generate(alpha, hi - lo, first=lo)
        c.peer_export()
        c.peer_set_timeout_ms(4000)
        ranks.append((c, lo, hi))
    ctxs = [c for c, _, _ in ranks]
    outs = torch.zeros(world, 64, dtype=torch.uint8, device=dev)
    ys = torch.zeros(world, 32, dtype=torch.uint8, device=dev)
    ws = torch.zeros(world, 64, dtype=torch.uint8, device=dev)
    scratch = torch.zeros(128, dtype=torch.uint8, device=dev)
    try:
        for r, (c, lo, hi) in enumerate(ranks):
            c.reserve(hi - lo)  # ranks sharing a device must not allocate once a peer may be spinning
            c.peer_attach_local(r, ctxs)
        exp_c = o.expected_commit(ints, alpha)
        exp_y, exp_w = o.expected_open(ints, u, alpha)
        for _ in range(2):
            outs.zero_(); ys.zero_(); ws.zero_()
            torch.cuda.synchronize()
            for r, (c, lo, hi) in enumerate(ranks):
                c.commit_sharded_dev(d_all.data_ptr() + lo * 32, hi - lo, outs[r].data_ptr())
                c.open_sharded_dev(d_all.data_ptr() + lo * 32, hi - lo, u, ys[r].data_ptr(), ws[r].data_ptr())
            for c in ctxs:
                c.sync()
            for r in range(world):
                assert mz.context.point_from_bytes(outs[r].cpu().numpy().tobytes()) == exp_c
                assert int.from_bytes(ys[r].cpu().numpy().tobytes(), "little") == exp_y
                assert mz.context.point_from_bytes(ws[r].cpu().numpy().tobytes()) == exp_w
    finally:
        for c in ctxs:
            c.close()


def test_peer_exchange_timeout_is_reported(ctx):
    """A rank whose peer never shows up gives up after the timeout and the next sync reports it."""
    import torch

    a, b = mz.Context(0), mz.Context(0)
    try:
        for c in (a, b):
            c.srs_generate(5, 4)
            c.reserve(4)
            c.peer_export()
            c.peer_set_timeout_ms(200)
        a.peer_attach_local(0, [a, b])
        b.peer_attach_local(1, [a, b])
        d = torch.zeros(4 * 4, dtype=torch.int64, device="cuda:0")
        out = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
        a.commit_sharded_dev(d.data_ptr(), 4, out.data_ptr())  # b never calls
        with pytest.raises(mz.MyzkpError):
            a.sync()
        # ranks sharing a device refuse to grow their scratch (a device-wide synchronisation could deadlock a
        # spinning peer) and say so, instead of stalling
        d2 = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda:0")
        b.srs_generate(5, 4096)  # set-up calls may allocate; the commit below would have to grow the MSM scratch
        with pytest.raises(mz.MyzkpError, match="myzkp_ctx_reserve"):
            b.commit_sharded_dev(d2.data_ptr(), 4096, out.data_ptr())
    finally:
        a.close(); b.close()


def test_commit_batch_one_pipeline(ctx):
    """myzkp_kzg_commit_batch: many small polynomials (DAS rows, Gemini levels) share one MSM pipeline
    (bucket range per polynomial); empty, constant, zero and max-length polynomials included."""
    rnd = random.Random(77)
    alpha = 0x5DEECE66D1234567
    nmax = 700
    ctx.srs_generate(alpha, nmax)
    sizes = [0, 1, 2, nmax, 3, 64, 65, 256] + [rnd.randrange(0, nmax + 1) for _ in range(150)]
    polys = []
    for j, n in enumerate(sizes):
        kind = j % 5
        if kind == 0:
            ints = [rnd.randrange(R) for _ in range(n)]
        elif kind == 1:
            ints = [rnd.randrange(256) for _ in range(n)]  # byte-valued, as the DAS callers produce
        elif kind == 2:
            ints = [rnd.choice([0, 0, 1, R - 1, R - 2]) for _ in range(n)]
        elif kind == 3:
            ints = [0] * n
        else:
            ints = [rnd.randrange(R) if rnd.random() < 0.3 else 0 for _ in range(n)]
        polys.append(ints)
    got = ctx.commit_batch(polys)
    assert len(got) == len(polys)
    for ints, g in zip(polys, got):
        assert g == o.expected_commit(ints, alpha)
    # batches of one and of none
    assert ctx.commit_batch([polys[3]]) == [o.expected_commit(polys[3], alpha)]
    assert ctx.commit_batch([]) == []
    # a polynomial longer than the SRS is rejected for the whole batch
    with pytest.raises(mz.MyzkpError):
        ctx.commit_batch([polys[0], [1] * (nmax + 1)])
    # non-canonical scalar anywhere in the batch
    bad = np.zeros((3, 32), np.uint8)
    bad[1] = np.frombuffer(R.to_bytes(32, "little"), dtype=np.uint8)
    with pytest.raises(mz.MyzkpError):
        ctx.commit_batch([polys[5], bad])


def test_commit_batch_with_partitioned_sort(ctx):
    """A batch whose bucket keys are wider than 16 bits takes the MSD-partitioned recode with per-polynomial bucket
    ranges (msm_recode_count / msm_recode_scatter with descriptors, csrc/msm.cu): forced windows 20, 22 and 24 on a
    small batch, ragged lengths, an empty and an all-zero polynomial."""
    rnd = random.Random(33)
    alpha = 0x1234567
    ctx.srs_generate(alpha, 3000)
    polys = [[rnd.randrange(R) for _ in range(n)] for n in (3000, 1, 517, 2048)] + [[], [0] * 40, [rnd.randrange(256) for _ in range(999)]]
    exp = [o.expected_commit(p_, alpha) for p_ in polys]
    try:
        for c in (20, 22, 24, 16):
            ctx.set_msm_params(c, 0)
            assert ctx.commit_batch(polys) == exp, c
    finally:
        ctx.set_msm_params(0, 0)


def test_commit_batch_mixed_large_and_small(ctx):
    """A batch with one polynomial above the own-MSM threshold (2^19) next to small ones: results come
    back in caller order."""
    alpha = 0x1F2E3D4C5B6A7988
    n_big = (1 << 19) + 5
    ctx.srs_generate(alpha, n_big)
    big = synth.random_scalars(n_big, 991)
    small = [synth.random_scalars(n, 992 + n) for n in (1000, 17, 40000)]
    got = ctx.commit_batch([small[0], big, small[1], small[2]])
    exp = [o.expected_commit(synth.limbs_to_ints(a), alpha) for a in (small[0], big, small[1], small[2])]
    assert got == exp


def test_chunked_upload_pipeline(ctx):
    """Host-buffer commit/open split into upload chunks (auto from 2^23) - forced here at a small size."""
    n = 5003
    alpha, u = 0x77777777777777777, 0x123456789
    coefs = synth.random_scalars(n, 21)
    ints = synth.limbs_to_ints(coefs)
    ctx.srs_generate(alpha, n)
    exp_c = o.expected_commit(ints, alpha)
    exp_o = o.expected_open(ints, u, alpha)
    try:
        for k in (1, 2, 3, 4, 7):
            ctx.set_upload_chunks(k)
            assert ctx.commit(coefs) == exp_c, k
            assert ctx.open(coefs, u) == exp_o, k
        ctx.set_upload_chunks(4)
        assert ctx.open([5, 7], 3) == o.expected_open([5, 7], 3, alpha)  # fewer coefficients than chunks
        assert ctx.commit([9]) == o.expected_commit([9], alpha)
    finally:
        ctx.set_upload_chunks(0)


def test_heavy_buckets_hierarchical_merge(ctx):
    """Many segments per bucket (skewed scalars, tiny segments): the fan-in head merge levels."""
    n = 1 << 15
    alpha = 0xDEADBEEFCAFE
    ctx.srs_generate(alpha, n)
    rnd = random.Random(8)
    dists = {
        "all_one": [1] * n,
        "bytes": [rnd.randrange(256) for _ in range(n)],
        "two_values": [rnd.choice([3, R - 3]) for _ in range(n)],
        "uniform": synth.limbs_to_ints(synth.random_scalars(n, 5)),
    }
    try:
        for name, sc in dists.items():
            exp = o.expected_commit(sc, alpha)
            for c, seg in ((8, 1), (12, 3), (16, 1), (0, 0)):
                ctx.set_msm_params(c, seg)
                assert ctx.commit(sc) == exp, (name, c, seg)
    finally:
        ctx.set_msm_params(0, 0)


def test_batch_open_degree_bound_and_open_gemini(ctx):
    """batch_open_kzg (kzg.rs:74-88), prove_degree_bound (kzg.rs:121-134),
    open_gemini (gemini.rs:116-144) against the faithful oracle (small) and algebraic values (larger)."""
    rnd = random.Random(77)
    # the reference's test_gemini shape: 8 coefficients, SRS of 9 points, rho = (2,3,4), beta = 1234
    alpha = rnd.randrange(R)
    coefs = list(range(1, 9))
    pk = mz.setup_kzg(mz.BN128.generator_g1(), None, 8, alpha=alpha, ctx=ctx)
    opk = o.setup_kzg(o.generator_g1(), 8, alpha)
    cms, polys = mz.split_and_fold_commit(coefs, [2, 3, 4], pk, want_folds=True)
    all_polys = [mz.Polynomial(coefs)] + polys
    proof = mz.open_gemini(all_polys, 1234, pk)
    ofs = o.split_and_fold([Fr(c) for c in coefs], [Fr(2), Fr(3), Fr(4)])
    oproof = o.open_gemini(ofs, Fr(1234), opk)
    assert len(proof.es) == 3 and len(proof.degree_proofs) == 4
    for e, oe in zip(proof.es, oproof.es):
        assert e.ys == [y.sanitize().value for y in oe.ys]
        assert e.w.as_tuple() == oe.w.affine_ints()
    assert [p.as_tuple() for p in proof.degree_proofs] == [p.affine_ints() for p in oproof.degree_proofs]
    with pytest.raises(RuntimeError):  # deg f > d
        mz.prove_degree_bound(mz.Polynomial(coefs), pk, 3)
    # larger, algebraic expected values
    n = 3000
    alpha = rnd.randrange(R)
    ints = [rnd.randrange(R) for _ in range(n)]
    ctx.srs_generate(alpha, n + 5)
    us = [rnd.randrange(R) for _ in range(3)]
    ys, w = ctx.batch_open(ints, us)
    assert (ys, w) == o.expected_batch_open(ints, us, alpha)
    ys1, w1 = ctx.batch_open(ints, us[:1])
    assert (ys1[0], w1) == o.expected_open(ints, us[0], alpha)  # k = 1 is open_kzg
    assert ctx.batch_open(ints[:2], us) == ([o.synthetic_division(ints[:2], u)[0] for u in us], None)
    for d in (n - 1, n + 2, n + 4):
        assert ctx.prove_degree_bound(ints, d) == o.expected_degree_bound(ints, alpha, n + 4, d)


def test_batched_affine_rounds_on_device(ctx):
    """BAA rounds (csrc/baa.cu) against the XYZZ-only accumulate and the oracle, across windows,
    segment lengths, round counts and scalar distributions (incl. duplicate / opposite SRS points)."""
    n = 6000
    alpha = 0x5A5A5A5A5A5A5A5A5A5A
    ctx.srs_generate(alpha, n)
    rnd = random.Random(4)
    dists = {
        "uniform": synth.limbs_to_ints(synth.random_scalars(n, 31)),
        "bytes": [rnd.randrange(256) for _ in range(n)],
        "all_one": [1] * n,
        "zeros": [0 if rnd.random() < 0.3 else rnd.randrange(R) for _ in range(n)],
    }
    try:
        for name, sc in dists.items():
            exp = o.expected_commit(sc, alpha)
            for rounds in (-2, 1, 2, 3, 6):  # -2 = the fused pair-sum accumulate (msm_accumulate_baa)
                for c, seg in ((8, 0), (12, 5), (16, 0), (16, 64), (20, 0)) + (((8, 300), (12, 130), (4, 1000), (8, 7)) if rounds == -2 else ()):
                    ctx.set_baa_rounds(rounds)
                    ctx.set_msm_params(c, seg)
                    assert ctx.commit(sc) == exp, (name, rounds, c, seg)
        # explicit SRS with repeated points, P / -P and infinity: doubling and cancellation inside a batch
        base = [o.fast_mul(k) for k in (3, 5, 7)]
        neg = lambda p: (p[0], o.P_MOD - p[1])
        pts = [base[0]] * 6 + [neg(base[0])] * 3 + [None, base[1], base[1], neg(base[1]), base[2]] * 2
        ctx.srs_load(pts)
        sc = [1] * len(pts)
        exp = None
        for p_ in pts:
            exp = o._fast_add(exp, p_)
        for rounds in (-2, 1, 2, 4):
            ctx.set_baa_rounds(rounds)
            for c, seg in ((8, 0), (8, 4), (16, 3), (8, 64), (4, 10)):
                ctx.set_msm_params(c, seg)
                assert ctx.commit(sc) == exp, (rounds, c, seg)
    finally:
        ctx.set_baa_rounds(-1)
        ctx.set_msm_params(0, 0)


def test_generic_msm_with_caller_points(ctx):
    """myzkp_g1_msm with caller-supplied points (accumulate_curve_points call sites, zksnark/utils.rs:83-93);
    the resident SRS must survive the call."""
    rnd = random.Random(12)
    alpha = 987654321
    ctx.srs_generate(alpha, 64)
    ks = [rnd.randrange(1, R) for _ in range(300)]
    pts = [o.fast_mul(k) for k in ks[:40]] * 7 + [None] * 20  # repeats and infinities
    sc = [rnd.randrange(R) for _ in pts]
    exp_k = sum(s * ks[i % 40] for i, s in enumerate(sc[:280])) % R
    assert ctx.g1_msm(sc, pts) == o.fast_mul(exp_k)
    # the faithful naive MSM of the reference on a small case
    small_pts, small_sc = pts[:12], sc[:12]
    opk = o.PublicKeyKZG([o.G1Point.new(o.Fq(x), o.Fq(y)) for x, y in small_pts])
    assert ctx.g1_msm(small_sc, small_pts) == o.commit_kzg(o.Polynomial([Fr(c) for c in small_sc]), opk).affine_ints()
    assert ctx.g1_msm([], []) is None
    # resident SRS intact
    coefs = [rnd.randrange(R) for _ in range(64)]
    assert ctx.srs_len == 64
    assert ctx.g1_msm(coefs) == o.expected_commit(coefs, alpha) == ctx.commit(coefs)
    # a large call: the points are an SRS read back from the device, so the expected value is the commitment;
    # windowed Pippenger without a table (per-window buckets + Horner over the windows) at two window choices
    for n in (5000, 70000):
        ctx.srs_generate(alpha, n)
        big_pts = ctx.srs_read(0, n)
        big_sc = synth.limbs_to_ints(synth.random_scalars(n, 900 + n))
        big_sc[7] = 0
        big_sc[8] = R - 1
        ctx.srs_generate(alpha, 8)  # a different resident SRS: the call must not use (or disturb) it
        assert ctx.g1_msm(big_sc, big_pts) == o.expected_commit(big_sc, alpha)
        assert ctx.srs_len == 8
    # error behaviour: coordinate >= p, scalar >= r
    bad_pt = [(o.P_MOD, 2)] + big_pts[1:5]
    with pytest.raises(mz.MyzkpError):
        ctx.g1_msm([1, 2, 3, 4, 5], bad_pt)
    bad_sc = synth.ints_to_limbs([1, 2, R, 4, 5])
    with pytest.raises(mz.MyzkpError):
        ctx.g1_msm(bad_sc, big_pts[:5])


def test_g2_powers_of_the_public_key(ctx):
    """setup_kzg's powers_2 = [g2, [alpha]g2] (kzg.rs:37) and setup_kzg_with_full_g2's [alpha^i]g2
    (kzg.rs:47-52), computed on the device, bit-exact vs the oracle's restatement of the reference's
    affine law over Fq2 (small cases) and its fast cross-check (larger ones)."""
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    pk = mz.setup_kzg(mz.BN128.generator_g1(), mz.BN128.generator_g2(), 7, alpha=alpha, ctx=ctx)
    assert len(pk.powers_2) == 2
    assert pk.powers_2[0] == mz.BN128.generator_g2()
    assert pk.powers_2[1].as_tuple() == o.g2_fast_mul(alpha)
    assert pk.powers_2[1].as_tuple() == o.setup_kzg_g2(o.generator_g2(), alpha, 2)[1].affine_ints()  # faithful path
    pk = mz.setup_kzg_with_full_g2(mz.BN128.generator_g1(), None, 40, alpha=alpha, ctx=ctx)
    assert len(pk) == 41 and len(pk.powers_2) == 41
    for i, p in enumerate(pk.powers_2):
        assert p.as_tuple() == o.g2_fast_mul(pow(alpha, i, R)), i
    # committed golden vectors (faithful oracle path)
    for case in G["g2"]:
        base = tuple(tuple(int(v) for v in c) for c in case["base"])
        got = ctx.srs_generate_g2(int(case["alpha"]), len(case["powers_2"]), 0, base)
        exp = [None if p is None else tuple(tuple(int(v) for v in c) for c in p) for p in case["powers_2"]]
        assert got == exp, case["name"]
    # window into the powers, an arbitrary base point, a small alpha against the faithful oracle
    base = o.g2_fast_mul(0xABCDEF)
    got = ctx.srs_generate_g2(3, 5, first=2, base=base)
    assert got == [o.g2_fast_mul(3 ** (2 + i), base) for i in range(5)]
    assert got[0] == o.G2Point.new(o.Fq2(base[0]), o.Fq2(base[1])).mul_ref(9).affine_ints()
    # alpha = 0: [g2, infinity, ...]; infinity as base stays infinity; the G1 and G2 halves share alpha
    assert ctx.srs_generate_g2(0, 3) == [(o.G2_GEN_X, o.G2_GEN_Y), None, None]
    assert ctx.srs_generate_g2(5, 2, base=None)[1] == o.g2_fast_mul(5)
    pk = mz.setup_kzg(mz.BN128.generator_g1(), mz.G2Point.point_at_infinity(), 1, alpha=alpha, ctx=ctx)
    assert all(p.is_point_at_infinity() for p in pk.powers_2)
    # non-canonical coordinates are rejected
    bad = np.zeros(128, np.uint8)
    bad[:32] = 0xFF
    a = np.frombuffer(int(5).to_bytes(32, "little"), dtype=np.uint8).copy()
    out = np.zeros(128, np.uint8)
    import ctypes
    code = ctx._lib.myzkp_srs_generate_g2(ctx.h, a.ctypes.data_as(ctypes.c_void_p), bad.ctypes.data_as(ctypes.c_void_p), 0, 1,
                                          out.ctypes.data_as(ctypes.c_void_p))
    assert code != 0


def test_g2_msm_with_caller_points(ctx):
    """accumulate_curve_points over G2 (zksnark/utils.rs:83-93; e.g. g2_r in tutorial_snark/protocol_2.rs:68):
    fold of acc + g * a over caller-supplied G2 points, vs the oracle."""
    rnd = random.Random(21)
    for n in (0, 1, 2, 63, 64, 65, 300):
        ks = [rnd.randrange(1, R) for _ in range(n)]
        pts = [o.g2_fast_mul(k) for k in ks]
        sc = [rnd.choice([0, 1, R - 1, rnd.randrange(R)]) for _ in range(n)]
        if n >= 2:
            pts[1] = None  # infinity among the points
        exp = o.g2_fast_mul(sum(s * k for s, k, p in zip(sc, ks, pts) if p is not None) % R)
        assert ctx.g2_msm(sc, pts) == exp, n
    # duplicates and opposite points cancel; the faithful affine law on a tiny case
    g = (o.G2_GEN_X, o.G2_GEN_Y)
    neg = (g[0], ((-g[1][0]) % o.P_MOD, (-g[1][1]) % o.P_MOD))
    assert ctx.g2_msm([5, 5], [g, neg]) is None
    assert ctx.g2_msm([3, 3], [g, g]) == o.g2_fast_mul(6)
    G2 = o.generator_g2()
    acc = o.G2Point.point_at_infinity()
    for a, k in [(7, 2), (11, 3)]:
        acc = acc + (G2 * k) * a
    assert ctx.g2_msm([7, 11], [o.g2_fast_mul(2), o.g2_fast_mul(3)]) == acc.affine_ints()


def test_pairing_values_match_the_reference_algorithm(ctx):
    """myzkp_pairing vs the oracle's restatement of optimal_ate_pairing (bn128.rs:147-181): identical Fq12
    coefficients; plus the identities of bn128.rs:341-365 (test_pairing) on the device values."""
    g1, g2 = o.generator_g1(), o.generator_g2()
    P1, Q1 = o.fast_mul(1), o.g2_fast_mul(1)
    vals = ctx.pairing([P1, o.fast_mul(R - 1), o.fast_mul(2), P1, o.fast_mul(37), o.fast_mul(999), None, P1],
                       [Q1, Q1, Q1, o.g2_fast_mul(2), o.g2_fast_mul(27), Q1, Q1, None])
    p1, pn1, p2, po2, p3, po3, inf_a, inf_b = [o.Fq12(v) for v in vals]
    assert p1 == o.optimal_ate_pairing(g1, g2)
    assert p3 == o.optimal_ate_pairing(g1.mul_ref(37), g2.mul_ref(27))
    assert p1 * pn1 == o.Fq12.one() and p1 * p1 == p2 and p1 * p1 == po2 and p3 == po3
    assert p1 != p2 and p1 != pn1 and inf_a == o.Fq12.one() and inf_b == o.Fq12.one()
    assert ctx.pairing_product_is_one([P1, o.fast_mul(R - 1)], [Q1, Q1])
    assert not ctx.pairing_product_is_one([P1, P1], [Q1, Q1])
    assert ctx.pairing_product_is_one([], [])


def test_verify_kzg_batch_degree_bound_and_gemini(ctx):
    """The reference's own protocol tests, end to end on the device: test_kzg (kzg.rs:152-175), test_batch_kzg
    (:177-205), test_degree_bound (:207-233) and test_gemini (gemini.rs:288-328); verify_kzg's boolean also
    against the oracle's restatement (three reference pairings)."""
    g1, g2 = mz.BN128.generator_g1(), mz.BN128.generator_g2()
    alpha = 123456789
    f = mz.Polynomial([6, 11, 6, 1])  # from_monomials([-1, -2, -3])
    pk = mz.setup_kzg(g1, g2, 3, alpha=alpha, ctx=ctx)
    c = mz.commit_kzg(f, pk)
    proof = mz.open_kzg(f, 5, pk)
    assert mz.verify_kzg(5, c, proof, pk)
    assert not mz.verify_kzg(6, c, proof, pk)
    assert not mz.verify_kzg(5, c, mz.ProofKZG(proof.y + 1, proof.w), pk)
    assert not mz.verify_kzg(5, mz.commit_kzg(mz.Polynomial([6, 11, 6, 2]), pk), proof, pk)
    opk = o.setup_kzg(o.generator_g1(), 3, alpha)
    op2 = o.setup_kzg_g2(o.generator_g2(), alpha, 2)
    oproof = o.open_kzg(o.Polynomial([Fr(v) for v in [6, 11, 6, 1]]), Fr(5), opk)
    oc = o.commit_kzg(o.Polynomial([Fr(v) for v in [6, 11, 6, 1]]), opk)
    assert o.verify_kzg(Fr(5), oc, oproof, opk.powers_1, op2) is True
    assert o.verify_kzg(Fr(6), oc, oproof, opk.powers_1, op2) is False
    # test_batch_kzg
    pk = mz.setup_kzg_with_full_g2(g1, g2, 3, alpha=alpha, ctx=ctx)
    c = mz.commit_kzg(f, pk)
    zs = [5, 7]
    bp = mz.batch_open_kzg(f, zs, pk)
    assert mz.batch_verify_kzg(zs, c, bp, pk)
    bp.ys[0] = (bp.ys[0] + 1) % R
    assert not mz.batch_verify_kzg(zs, c, bp, pk)
    # test_degree_bound: deg f = 3 is bounded by 3, and a proof for d = 3 does not verify for d = 2
    dp = mz.prove_degree_bound(f, pk, 3)
    assert mz.verify_degree_bound(c, dp, pk, 3)
    assert not mz.verify_degree_bound(c, dp, pk, 2)
    pk4 = mz.setup_kzg_with_full_g2(g1, g2, 4, alpha=alpha, ctx=ctx)  # the reference's case: max_d = 4, d = 3
    assert mz.verify_degree_bound(mz.commit_kzg(f, pk4), mz.prove_degree_bound(f, pk4, 3), pk4, 3)
    # test_gemini
    pk = mz.setup_kzg_with_full_g2(g1, g2, 8, alpha=synth.random_scalar(synth.SEED_ALPHA), ctx=ctx)
    coef, rhos = list(range(1, 9)), [2, 3, 4]
    cms, polys = mz.split_and_fold_commit(coef, rhos, pk, want_folds=True)
    fs = [mz.Polynomial(coef)] + polys
    mu = 382  # gemini.rs:294-307: sum coef_i * tensor(rhos)_i
    assert fs[-1]._wire() == [mu]
    gp = mz.open_gemini(fs, 1234, pk)
    assert mz.verify_gemini(rhos, mu, 1234, cms, gp, pk)
    assert not mz.verify_gemini(rhos, mu + 1, 1234, cms, gp, pk)


def _build_cpp(name):
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", name)
    src = exe + ".cpp"
    if not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(root, "include"), "-o", exe, src,
                               "-L", os.path.join(root, "myzkp_b200"), "-lmyzkp_b200",
                               "-Wl,-rpath," + os.path.join(root, "myzkp_b200")])
    return exe


def test_cpp_host_through_header_mirror():
    """tests/cpp/abi_smoke.cpp: a C++ host over include/myzkp_b200.hpp reproduces the test_kzg anchor, including the
    range-sharded commit and open with two ranks (own host threads) sharing this GPU.  No retry: the ranks' scratch
    is reserved before they attach, so nothing synchronises the device while a peer spins in the exchange."""
    import subprocess

    exe = _build_cpp("abi_smoke")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    out = res.stdout
    vals = dict(line.split() for line in out.strip().splitlines() if " " in line)
    assert out.strip().endswith("OK"), f"rc={res.returncode} stdout={out!r} stderr={res.stderr[-2000:]!r}"
    assert int(vals["C.x"], 16) == 8096424998935924997123460782489249937183001369792870392058374165119638207724
    assert int(vals["C.y"], 16) == 14698683656276342473960081670169131092130433153277961881223581660609015832377
    assert int(vals["y"], 16) == 336
    assert int(vals["G2.2x0"], 16) == 18029695676650738226693292988307914797657423701064905010927197838374790804409
    assert int(vals["W.x"], 16) == 15737316170989375530370354340609809222984715696988518295913516551941326522818
    assert int(vals["W.y"], 16) == 13254863773102499080085687663253363332659578358445016579269372167128026496803


def _splitmix_coefs(n, seed=42):
    """the coefficients tests/cpp/abi_multi.cpp generates (splitmix64, top limb >> 3)"""
    M = (1 << 64) - 1
    x = seed
    out = np.empty((n, 4), dtype=np.uint64)
    for i in range(n):
        w = []
        for _ in range(4):
            x = (x + 0x9E3779B97F4A7C15) & M
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
            w.append(z ^ (z >> 31))
        w[3] >>= 3
        out[i] = w
    return out


def test_cpp_host_drives_several_gpus_in_one_process():
    """tests/cpp/abi_multi.cpp: ONE C++ process, no torch, no NCCL - the multi-device context (csrc/multi.cu) shards
    the SRS over every visible GPU (two ranks on GPU 0 when there is only one), commits and opens a 2^14 + 3
    polynomial; the printed commitment and proof must be the oracle's."""
    import subprocess

    exe = _build_cpp("abi_multi")
    res = subprocess.run([exe, "14"], capture_output=True, text=True, timeout=300)
    out = res.stdout
    assert out.strip().endswith("OK"), f"rc={res.returncode} stdout={out!r} stderr={res.stderr[-2000:]!r}"
    vals = dict(line.split() for line in out.strip().splitlines() if " " in line)
    n = (1 << 14) + 3
    assert int(vals["n"]) == n and int(vals["world"]) >= 2
    ints = synth.limbs_to_ints(_splitmix_coefs(n))
    alpha, u = 0x1234567890ABCDEF, 0xFEDCBA9876543
    assert (int(vals["C.x"], 16), int(vals["C.y"], 16)) == o.expected_commit(ints, alpha)
    ey, ew = o.expected_open(ints, u, alpha)
    assert int(vals["y"], 16) == ey and (int(vals["W.x"], 16), int(vals["W.y"], 16)) == ew


@pytest.mark.parametrize("logn", [16, 20])
def test_multi_device_context_commit_open(logn):
    """BASELINE config 3 (commit + open at 2^20) through the multi-device context on every visible GPU: real peers
    over NVLink when the box has several GPUs, two ranks sharing GPU 0 otherwise.  Checked against the oracle."""
    from myzkp_b200 import _lib

    ndev = _lib.load().myzkp_device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0]
    n = (1 << logn) - 5
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    u = synth.random_scalar(synth.SEED_OPEN)
    coefs = synth.random_scalars(n, synth.SEED_SCALARS + logn)
    mc = mz.MultiContext(devices)
    try:
        mc.srs_generate(alpha, n)
        assert mc.srs_len == n and mc.world == len(devices)
        cb = coefs.tobytes()
        fa = oc.fr_eval_bytes(cb, n, alpha)
        y = oc.fr_eval_bytes(cb, n, u)
        exp_c = o.fast_mul(fa)
        exp_w = o.fast_mul((fa - y) * pow((alpha - u) % R, -1, R) % R)
        for _ in range(2):  # both exchange parities
            assert mc.commit(coefs) == exp_c
            assert mc.open(coefs, u) == (y, exp_w)
        # a short polynomial leaves the upper ranks empty; constants and the empty polynomial
        ints = synth.limbs_to_ints(coefs[:7])
        assert mc.commit(coefs[:7]) == o.expected_commit(ints, alpha)
        assert mc.open(coefs[:7], u) == o.expected_open(ints, u, alpha)
        assert mc.commit(coefs[:0]) is None
        assert mc.open(coefs[:1], u) == (ints[0], None)
        with pytest.raises(mz.MyzkpError):
            mc.commit(np.concatenate([coefs, coefs[:1]]))
        bad = coefs[:64].copy()
        bad[3] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)  # >= r
        with pytest.raises(mz.MyzkpError):
            mc.commit(bad)
        assert mc.commit(coefs) == exp_c  # still usable after the error
    finally:
        mc.close()


@pytest.mark.parametrize("logn", [18, 19, 20])
def test_gemini_fold_commit_at_config_size(ctx, logn):
    """BASELINE config 5: split_and_fold + commit_gemini of a 2^20 multilinear (21 commitments), and the sizes either
    side of the batching threshold (2^19: one own-stream MSM + the batched child pipeline; 2^18: everything on the
    child).  Every level's commitment against [f_i(alpha)]G with f_i from the oracle's fold, every folded
    coefficient against the oracle; run twice back to back so the child stream's fork / join is exercised while
    the previous call's buffers are being reused."""
    n = 1 << logn
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    coefs = synth.random_scalars(n, synth.SEED_GEMINI_COEF)
    rhos = synth.limbs_to_ints(synth.random_scalars(logn, synth.SEED_GEMINI_RHO))
    ctx.srs_generate(alpha, n)
    level = coefs.tobytes()
    exp_pts, exp_folds, ln = [], [], n
    for i in range(logn + 1):
        exp_pts.append(o.fast_mul(oc.fr_eval_bytes(level, ln, alpha)))
        if i < logn:
            level = oc.fold_bytes(level, ln // 2, rhos[i])  # gemini.rs:71-98
            exp_folds.append(level)
            ln //= 2
    pts, folds = ctx.gemini_fold_commit(coefs, rhos, want_folds=True)
    pts2 = ctx.gemini_fold_commit(coefs, rhos)
    pts3 = ctx.gemini_fold_commit(coefs, rhos)
    assert len(pts) == logn + 1
    assert pts == exp_pts
    assert pts2 == exp_pts and pts3 == exp_pts
    assert folds.tobytes() == b"".join(exp_folds)
    # spot-check the C fold against the Python oracle on the small tail levels
    small = synth.limbs_to_ints(np.frombuffer(exp_folds[logn - 4], dtype=np.uint64).reshape(-1, 4))
    assert o.fold_ints(small, rhos[logn - 3:])[-1] == synth.limbs_to_ints(np.frombuffer(exp_folds[-1], dtype=np.uint64).reshape(-1, 4))
    # the mirror's argument checks (SplitFoldError, gemini.rs:55-66) and the ABI's own n_rhos check
    with pytest.raises(mz.MyzkpError):
        ctx.gemini_fold_commit(coefs[:16], rhos[:3])


def test_open_and_commit_at_2pow24(ctx):
    """North-star size on one GPU: commit AND open of a degree-(2^24 - 1) polynomial, bit-exact against
    C = [f(alpha)]G, y = f(u), W = [(f(alpha) - y)/(alpha - u)]G (f(alpha), f(u) by the oracle's C evaluation),
    through the host API (chunked upload pipeline) and through the device-pointer entry points."""
    import torch

    logn = 24
    n = 1 << logn
    free, _ = torch.cuda.mem_get_info(0)
    if free < 100 * (1 << 30):
        pytest.skip("needs ~85 GiB of free HBM for the 2^24 table")
    alpha = synth.random_scalar(synth.SEED_ALPHA)
    u = synth.random_scalar(synth.SEED_OPEN)
    coefs = synth.random_scalars(n, synth.SEED_SCALARS + logn)
    cb = coefs.tobytes()
    fa = oc.fr_eval_bytes(cb, n, alpha)
    y = oc.fr_eval_bytes(cb, n, u)
    exp_c = o.fast_mul(fa)
    exp_w = o.fast_mul((fa - y) * pow((alpha - u) % R, -1, R) % R)
    ctx.srs_generate(alpha, n)
    try:
        assert ctx.commit(coefs) == exp_c
        assert ctx.open(coefs, u) == (y, exp_w)
        dev = torch.device("cuda", 0)
        d = torch.from_numpy(coefs.view(np.int64).reshape(-1)).to(dev)
        out = torch.zeros(96, dtype=torch.uint8, device=dev)
        ctx.open_dev(d.data_ptr(), n, u, out.data_ptr(), out.data_ptr() + 32)
        ctx.sync()
        raw = out.cpu().numpy().tobytes()
        assert int.from_bytes(raw[:32], "little") == y
        assert mz.context.point_from_bytes(raw[32:]) == exp_w
        # an asynchronous device-pointer call with a scalar >= r is reported by the next sync (sticky flag)
        d[4 * 12345 : 4 * 12345 + 4] = -1
        ctx.open_dev(d.data_ptr(), n, u, out.data_ptr(), out.data_ptr() + 32)
        with pytest.raises(mz.MyzkpError):
            ctx.sync()
        ctx.sync()  # reported once, then cleared
    finally:
        ctx.srs_generate(alpha, 16)  # release the 70 GiB table for the tests that follow


def test_quotient_scan_ragged_sizes_carries_and_special_points(ctx):
    """The single-pass look-back scan (csrc/poly.cu) against the oracle's synthetic division
    (polynomial.rs:371-405 for a linear divisor): sizes around the tile (2048) and warp-chunk (16, 512)
    boundaries, u in {0, 1, r-1, random}, a non-zero carry entering the range (sharded open), and the
    evaluation-only form; also a non-canonical coefficient must raise the flag through the fused check."""
    import torch

    rng = random.Random(77)
    sizes = [1, 2, 15, 16, 17, 511, 512, 513, 2047, 2048, 2049, 4095, 4096, 4097, 6144, 70001, (1 << 17) + 5]
    for n in sizes:
        ints = [rng.randrange(R) for _ in range(n)]
        coefs = synth.ints_to_limbs(ints)
        us = [rng.randrange(R)] if n > 5000 else [0, 1, R - 1, rng.randrange(R)]
        for u in us:
            ey, eq = o.synthetic_division(ints, u)
            yq, q = ctx.fr_quotient(coefs, u)
            assert yq == ey, (n, u)
            assert synth.limbs_to_ints(q.view(np.uint64).reshape(-1, 4)) == eq, (n, u)
            assert ctx.fr_eval(coefs, u) == ey, (n, u)
        # carry entering the range from above + (h, u^n) of the range, device-pointer forms
        u = rng.randrange(R)
        carry = rng.randrange(R)
        d = torch.from_numpy(coefs.view(np.int64).reshape(-1).copy()).cuda()
        dq = torch.zeros(n * 4, dtype=torch.int64, device="cuda")
        dc0 = torch.zeros(12, dtype=torch.int64, device="cuda")
        ctx.fr_range_quotient_dev(d.data_ptr(), n, u, carry, dq.data_ptr(), dc0.data_ptr())
        ctx.fr_range_eval_dev(d.data_ptr(), n, u, dc0.data_ptr() + 32, dc0.data_ptr() + 64)
        ctx.sync()
        c = carry
        exp_q = [0] * n
        for i in range(n - 1, -1, -1):
            exp_q[i] = c
            c = (ints[i] + u * c) % R
        got_q = synth.limbs_to_ints(dq.cpu().numpy().view(np.uint64).reshape(-1, 4))
        assert got_q == exp_q, n
        got = synth.limbs_to_ints(dc0.cpu().numpy().view(np.uint64).reshape(-1, 4))
        assert got[0] == c, n
        h = 0
        for i in range(n - 1, -1, -1):
            h = (ints[i] + u * h) % R
        assert got[1] == h and got[2] == pow(u, n, R), n
    # fused canonicity check
    n = 5000
    ints = [rng.randrange(R) for _ in range(n)]
    coefs = synth.ints_to_limbs(ints)
    coefs[3333] = synth.ints_to_limbs([R])[0]  # r itself: the smallest non-canonical value
    with pytest.raises(mz.MyzkpError):
        ctx.fr_quotient(coefs, 5)
    with pytest.raises(mz.MyzkpError):
        ctx.fr_eval(coefs, 5)

"""Randomised parity soak on the GPU: sizes around tile/segment boundaries, forced windows and segment
lengths, skewed scalar distributions; commit and open against the oracle's expected values."""
import random
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import myzkp_b200 as mz
import myzkp_oracle as o

R = o.R_MOD
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 90.0
rnd = random.Random(seed)
ctx = mz.Context(0)
t0 = time.time()
cases = 0
specials = [1, 2, 3, 15, 16, 17, 255, 256, 257, 2047, 2048, 2049, 4095, 4096, 4097, 8191, 8192, 8193, 65535, 65536, 65537]
while time.time() - t0 < budget:
    n = rnd.choice(specials) if rnd.random() < 0.5 else rnd.randrange(1, 40000)
    alpha = rnd.randrange(1, R)
    u = rnd.choice([0, 1, rnd.randrange(R), alpha])
    kind = rnd.randrange(6)
    if kind == 0:
        sc = [rnd.randrange(R) for _ in range(n)]
    elif kind == 1:
        sc = [rnd.randrange(256) for _ in range(n)]
    elif kind == 2:
        sc = [0 if rnd.random() < 0.5 else rnd.randrange(R) for _ in range(n)]
    elif kind == 3:
        v = rnd.choice([1, R - 1, rnd.randrange(R)])
        sc = [v] * n
    elif kind == 4:
        sc = [(1 << rnd.randrange(254)) % R for _ in range(n)]
    else:
        sc = [rnd.choice([R - 1, R - 2, 1 << 253, (1 << 253) - 1, 0x8000800080008000]) % R for _ in range(n)]
    ctx.srs_generate(alpha, n)
    ctx.set_msm_params(rnd.choice([0, 0, 4, 8, 12, 16, 20, 22, 24]), rnd.choice([0, 0, 0, 1, 2, 5, 33, 300]))
    ctx.set_upload_chunks(rnd.choice([0, 0, 1, 2, 3, 5]))
    ctx.set_baa_rounds(rnd.choice([-1, -1, -1, 1, 3]))
    ctx._lib.myzkp_test_set_sort_group_cap(rnd.choice([0, 0, 64, 600, 4096]))  # MSD sort: oversize-group paths
    exp_c = o.expected_commit(sc, alpha)
    got_c = ctx.commit(sc)
    assert got_c == exp_c, ("commit", seed, cases, n, kind)
    if u != alpha:
        assert ctx.open(sc, u) == o.expected_open(sc, u, alpha), ("open", seed, cases, n, kind)
    else:
        assert ctx.open(sc, u)[0] == o.synthetic_division(sc, u)[0]
    cases += 1
print(f"fuzz ok: seed {seed}, {cases} cases in {time.time() - t0:.0f} s")

"""CPU: the C restatement (oracle/oracle.c) against the Python oracle and the golden vectors."""
import json
import os
import random
import time

import myzkp_oracle as o
import oracle_c as oc

P, R = o.P_MOD, o.R_MOD
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kzg_golden.json")))


def _pt(p):
    return None if p is None else (int(p[0]), int(p[1]))


def test_field_ops_vs_python_ints():
    rnd = random.Random(2)
    for field, m in ((0, P), (1, R)):
        edge = [0, 1, 2, m - 1, m - 2, m // 2, (1 << 253), 0xFFFFFFFFFFFFFFFF, (1 << 128) - 1]
        vals = edge + [rnd.randrange(m) for _ in range(200)]
        for x in vals:
            for y in rnd.sample(vals, 6) + edge[:4]:
                assert oc.fe_op(field, 0, x, y) == (x + y) % m
                assert oc.fe_op(field, 1, x, y) == (x - y) % m
                assert oc.fe_op(field, 2, x, y) == x * y % m
            assert oc.fe_op(field, 3, x) == (pow(x, -1, m) if x else 0)


def test_group_law_vs_python_oracle():
    rnd = random.Random(3)
    g = (1, 2)
    assert oc.g1_mul(g, 2) == (o.generator_g1() * 2).affine_ints()
    assert oc.g1_mul(g, R) is None and oc.g1_mul(g, 0) is None
    pts = [o.fast_mul(rnd.randrange(1, R)) for _ in range(6)] + [None, g]
    neg = lambda p: None if p is None else (p[0], P - p[1])
    for a in pts:
        for b in pts + [neg(a)]:
            assert oc.g1_add(a, b) == o._fast_add(a, b)
    for _ in range(5):
        k = rnd.randrange(R)
        assert oc.g1_mul(g, k) == o.fast_mul(k)
    assert oc.g1_mul(g, 12345) == (o.generator_g1() * 12345).affine_ints()


def test_golden_vectors():
    for case in G["kzg"]:
        alpha, coefs, u = int(case["alpha"]), [int(c) for c in case["coefs"]], int(case["u"])
        n = len(coefs)
        srs = oc.setup_kzg_bytes(alpha, n, threads=2)
        assert [oc._unpt(srs[64 * i : 64 * i + 64]) for i in range(n)] == [_pt(p) for p in case["srs"]]
        cb = b"".join(int(c).to_bytes(32, "little") for c in coefs)
        assert oc.commit_kzg_bytes(cb, srs, n, threads=3) == _pt(case["commit"])
        assert oc.commit_kzg_bytes(cb, srs, n, threads=1) == _pt(case["commit"])
        y, w = oc.open_kzg_bytes(cb, n, u, srs)
        assert (y, w) == (int(case["y"]), _pt(case["w"])), case["name"]
    for case in G["gemini"]:
        cur = [int(c) for c in case["coefs"]]
        for lvl, rho in enumerate(case["rhos"]):
            cb = b"".join(int(c).to_bytes(32, "little") for c in cur)
            out = oc.fold_bytes(cb, len(cur) // 2, int(rho))
            cur = [int.from_bytes(out[32 * i : 32 * i + 32], "little") for i in range(len(cur) // 2)]
            assert cur == [int(v) for v in case["folds"][lvl + 1]]


def test_quotient_and_eval_vs_python():
    rnd = random.Random(9)
    for n in (1, 2, 3, 17, 200):
        coefs = [rnd.randrange(R) for _ in range(n)]
        if n > 3:
            coefs[-1] = 0  # trailing zero is trimmed by the reference
        u = rnd.randrange(R)
        cb = b"".join(int(c).to_bytes(32, "little") for c in coefs)
        ey, eq = o.synthetic_division(coefs, u)
        assert oc.fr_eval_bytes(cb, n, u) == ey
        y, q = oc.quotient_bytes(cb, n, u)
        assert y == ey and q == eq

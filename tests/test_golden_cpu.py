"""CPU: the committed golden vectors agree with the oracle's O(N) expected-value
path (the path used for sizes the faithful oracle cannot reach)."""
import json
import os

import myzkp_oracle as o

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kzg_golden.json")))


def _pt(p):
    return None if p is None else (int(p[0]), int(p[1]))


def test_golden_kzg_matches_expected_value_path():
    for case in G["kzg"]:
        alpha, coefs, u = int(case["alpha"]), [int(c) for c in case["coefs"]], int(case["u"])
        assert o.expected_commit(coefs, alpha) == _pt(case["commit"]), case["name"]
        y, w = o.expected_open(coefs, u, alpha)
        assert y == int(case["y"]) and w == _pt(case["w"]), case["name"]
        for i, p in enumerate(case["srs"]):
            assert o.fast_mul(pow(alpha, i, o.R_MOD)) == _pt(p)


def test_golden_gemini_matches_fold_ints():
    for case in G["gemini"]:
        coefs, rhos, alpha = [int(c) for c in case["coefs"]], [int(r) for r in case["rhos"]], int(case["alpha"])
        folds = o.fold_ints(coefs, rhos)
        assert [[int(v) for v in f] for f in case["folds"]] == folds
        for f, c in zip(folds, case["commitments"]):
            assert o.expected_commit(f, alpha) == _pt(c)


def _pt2(p):
    return None if p is None else ((int(p[0][0]), int(p[0][1])), (int(p[1][0]), int(p[1][1])))


def test_golden_g2_matches_fast_path():
    for case in G["g2"]:
        alpha, base = int(case["alpha"]), _pt2(case["base"])
        for i, p in enumerate(case["powers_2"]):
            assert o.g2_fast_mul(pow(alpha, i, o.R_MOD), base) == _pt2(p), case["name"]


def test_golden_pairing_values():
    """The committed pairing values are what the oracle's restatement of optimal_ate_pairing gives, and they are bilinear:
    e(37 G1, 27 G2) = e(G1, G2)^999 (bn128.rs:362-364)."""
    vals = {}
    for case in G["pairing"]:
        k1, k2 = int(case["g1_multiple"]), int(case["g2_multiple"])
        e = o.optimal_ate_pairing(o.generator_g1().mul_ref(k1), o.generator_g2().mul_ref(k2))
        assert e.c == [int(v) for v in case["fq12"]], (k1, k2)
        vals[(k1, k2)] = e
    assert vals[(1, 1)].pow(999) == vals[(37, 27)]

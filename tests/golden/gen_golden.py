"""Generates tests/golden/kzg_golden.json with the FAITHFUL oracle path
(oracle/myzkp_oracle.py: affine double-and-add with ext-Euclid inversions, naive
MSM, schoolbook division - the restatement of kzg.rs / polynomial.rs / curve.rs).
Run from the repo root:  python tests/golden/gen_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import myzkp_oracle as o  # noqa: E402
from myzkp_oracle import Fr, Polynomial  # noqa: E402


def pt(p):
    t = p.affine_ints()
    return None if t is None else [str(t[0]), str(t[1])]


def main():
    rnd = random.Random(0xB200)
    cases = []
    g = o.generator_g1()
    specs = [
        ("test_kzg_poly", 123456789, [6, 11, 6, 1], 5),
        ("n1", 987654321, [rnd.randrange(o.R_MOD)], rnd.randrange(o.R_MOD)),
        ("n2", 55555, [rnd.randrange(o.R_MOD) for _ in range(2)], rnd.randrange(o.R_MOD)),
        ("n16", rnd.randrange(o.R_MOD), [rnd.randrange(o.R_MOD) for _ in range(16)], rnd.randrange(o.R_MOD)),
        ("n33_bytes", rnd.randrange(o.R_MOD), [rnd.randrange(256) for _ in range(33)], rnd.randrange(o.R_MOD)),
        ("n48_zeros_and_edges", rnd.randrange(o.R_MOD),
         [0, 1, o.R_MOD - 1, 0, 2, o.R_MOD - 2] + [rnd.choice([0, rnd.randrange(o.R_MOD)]) for _ in range(42)], 0),
        ("n64", rnd.randrange(o.R_MOD), [rnd.randrange(o.R_MOD) for _ in range(64)], rnd.randrange(o.R_MOD)),
    ]
    for name, alpha, coefs, u in specs:
        pk = o.setup_kzg(g, len(coefs) - 1, alpha)
        f = Polynomial([Fr(c) for c in coefs])
        c = o.commit_kzg(f, pk)
        pr = o.open_kzg(f, Fr(u), pk)
        cases.append({
            "name": name, "alpha": str(alpha), "coefs": [str(x) for x in coefs], "u": str(u),
            "srs": [pt(p) for p in pk.powers_1],
            "commit": pt(c), "y": str(pr.y.sanitize().value), "w": pt(pr.w),
        })
        print(name, "done", file=sys.stderr)
    # Gemini: test_gemini coefficients (gemini.rs:294-297) and a random 16
    gem = []
    for name, alpha, coefs, rhos in [
        ("test_gemini", 424242, list(range(1, 9)), [2, 3, 4]),
        ("book_example", 424242, list(range(1, 9)), [1, 2, 3]),
        ("n16", rnd.randrange(o.R_MOD), [rnd.randrange(o.R_MOD) for _ in range(16)], [rnd.randrange(o.R_MOD) for _ in range(4)]),
    ]:
        pk = o.setup_kzg(g, len(coefs) - 1, alpha)
        fs = o.split_and_fold([Fr(c) for c in coefs], [Fr(r) for r in rhos])
        cm = o.commit_gemini(fs, pk)
        gem.append({"name": name, "alpha": str(alpha), "coefs": [str(x) for x in coefs], "rhos": [str(x) for x in rhos],
                    "folds": [[str(v) for v in p.canonical()] for p in fs], "commitments": [pt(p) for p in cm]})
        print(name, "done", file=sys.stderr)
    # G2 half of the public key (kzg.rs:37, 47-52) with the reference's affine law over Fq2
    def pt2(p):
        t = p.affine_ints()
        return None if t is None else [[str(t[0][0]), str(t[0][1])], [str(t[1][0]), str(t[1][1])]]

    g2 = o.generator_g2()
    g2s = []
    other = g2.mul_ref(0xABCDEF)
    for name, alpha, n, base in [("setup_kzg_alpha_123456789", 123456789, 2, g2),
                                 ("full_g2_random_alpha", rnd.randrange(o.R_MOD), 5, g2),
                                 ("other_base", rnd.randrange(o.R_MOD), 3, other),
                                 ("alpha_zero", 0, 3, g2)]:
        g2s.append({"name": name, "alpha": str(alpha), "base": pt2(base), "powers_2": [pt2(p) for p in o.setup_kzg_g2(base, alpha, n)]})
        print(name, "done", file=sys.stderr)
    # pairing values (optimal_ate_pairing, bn128.rs:147-181) as the 12 coefficients of w^k
    pair = []
    for k1, k2 in [(1, 1), (37, 27)]:
        e = o.optimal_ate_pairing(o.generator_g1().mul_ref(k1), g2.mul_ref(k2))
        pair.append({"g1_multiple": k1, "g2_multiple": k2, "fq12": [str(v) for v in e.c]})
        print("pairing", k1, k2, "done", file=sys.stderr)
    out = {"generator": "tests/golden/gen_golden.py (faithful oracle path)", "kzg": cases, "gemini": gem, "g2": g2s, "pairing": pair}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kzg_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

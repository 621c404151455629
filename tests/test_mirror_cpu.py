"""CPU: the host logic of the Python mirror (myzkp_b200/kzg.py, gemini.py) - sanitisation, argument checks, the
verifier equations, the Gemini open / verify bookkeeping - driven through a stand-in for the device context that
answers every call from the oracle.  TEST INFRASTRUCTURE: the stand-in is the checker, not a fallback; the product's
Context refuses to exist without a GPU (tests/test_abi_symbols.py::test_no_cpu_fallback_without_gpu).  The same flows
run on the device in tests/test_gpu_kzg.py."""
import numpy as np
import pytest

import myzkp_oracle as o

import myzkp_b200 as mz
from myzkp_b200 import _lib

R, P = o.R_MOD, o.P_MOD


class OracleContext:
    """Answers the Context calls the mirror makes, from the oracle's expected-value paths (SRS = [alpha^i]G)."""

    def __init__(self):
        self.alpha, self.n = None, 0

    # SRS
    def srs_generate(self, alpha, n, first=0):
        assert first == 0
        self.alpha, self.n = int(alpha) % R, n

    def srs_generate_g2(self, alpha, n, first=0, base=None):
        base = (o.G2_GEN_X, o.G2_GEN_Y) if base is None else base
        return [o.g2_fast_mul(pow(int(alpha) % R, first + i, R), base) for i in range(n)]

    def srs_read(self, off, n):
        return [o.fast_mul(pow(self.alpha, off + i, R)) for i in range(n)]

    @property
    def srs_len(self):
        return self.n

    def _canon(self, coefs):
        coefs = [int(c) for c in coefs]
        if any(not 0 <= c < R for c in coefs):
            raise _lib.MyzkpError(-3, "non-canonical scalar")  # the raw ABI rejects what the shim must sanitise
        if len(coefs) > self.n:
            raise _lib.MyzkpError(-2, "polynomial longer than the SRS (reference panics at polynomial.rs:162)")
        return coefs

    # prover
    def commit(self, coefs):
        return o.expected_commit(self._canon(coefs), self.alpha)

    def commit_batch(self, polys):
        return [self.commit(p) for p in polys]

    def open(self, coefs, u):
        return o.expected_open(self._canon(coefs), u, self.alpha)

    def batch_open(self, coefs, us):
        return o.expected_batch_open(self._canon(coefs), us, self.alpha)

    def prove_degree_bound(self, coefs, d):
        coefs = self._canon(coefs)
        if len(coefs) - 1 > d:
            raise _lib.MyzkpError(-2, "deg f > d")
        return o.expected_degree_bound(coefs, self.alpha, self.n - 1, d)

    def gemini_fold_commit(self, coefs, rhos, want_folds=False):
        folds = o.fold_ints(self._canon(coefs), [int(r) for r in rhos])
        pts = [o.expected_commit(f, self.alpha) for f in folds]
        if not want_folds:
            return pts
        rows = [int(v).to_bytes(32, "little") for f in folds[1:] for v in f]
        return pts, np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(-1, 32)

    # group operations and pairings of the verifier
    def g1_msm(self, scalars, points=None):
        acc = None
        pts = self.srs_read(0, len(scalars)) if points is None else points
        for s, p in zip(scalars, pts):
            acc = o._fast_add(acc, o.fast_mul(int(s) % R, p) if p is not None else None)
        return acc

    def g2_msm(self, scalars, points):
        acc = None
        for s, p in zip(scalars, points):
            acc = o._g2_fast_add(acc, o.g2_fast_mul(int(s) % R, p) if p is not None else None)
        return acc

    def pairing_product_is_one(self, g1s, g2s):
        prod = o.Fq12.one()
        for p, q in zip(g1s, g2s):
            gp = o.G1Point.point_at_infinity() if p is None else o.G1Point.new(o.Fq(p[0]), o.Fq(p[1]))
            gq = o.G2Point.point_at_infinity() if q is None else o.G2Point.new(o.Fq2(list(q[0])), o.Fq2(list(q[1])))
            prod = prod * o.optimal_ate_pairing(gp, gq)
        return prod == o.Fq12.one()


ALPHA = 123456789


def _pk(max_d, full_g2=False):
    setup = mz.setup_kzg_with_full_g2 if full_g2 else mz.setup_kzg
    return setup(mz.BN128.generator_g1(), mz.BN128.generator_g2(), max_d, alpha=ALPHA, ctx=OracleContext())


def test_setup_commit_open_sanitise_and_errors():
    pk = _pk(3)
    assert len(pk) == 4 and len(pk.powers_1) == 4 and len(pk.powers_2) == 2  # kzg.rs:32, :37
    assert pk.powers_1[0] == mz.BN128.generator_g1() and pk.powers_2[0] == mz.BN128.generator_g2()
    # from_monomials([-1, -2, -3]) with negative internal values: the shim sanitises (polynomial.rs:162)
    f_neg = mz.Polynomial([6 - R, 11, 6 - 2 * R, 1])
    c = mz.commit_kzg(f_neg, pk)
    assert c.as_tuple() == o.expected_commit([6, 11, 6, 1], ALPHA)
    pr = mz.open_kzg(f_neg, 5 - R, pk)
    assert pr.y == 336 and pr.w.as_tuple() == o.expected_open([6, 11, 6, 1], 5, ALPHA)[1]  # kzg.rs:157-169
    assert mz.commit_kzg(mz.Polynomial([]), pk).is_point_at_infinity()
    with pytest.raises(_lib.MyzkpError):
        mz.commit_kzg(mz.Polynomial([1, 2, 3, 4, 5]), pk)  # longer than the SRS: the reference index-panics
    with pytest.raises(ValueError):
        mz.setup_kzg(mz.G1Point(1, P - 2), None, 3, alpha=ALPHA, ctx=OracleContext())  # only the standard G1 generator
    pk_inf = mz.setup_kzg(mz.BN128.generator_g1(), mz.G2Point.point_at_infinity(), 1, alpha=ALPHA, ctx=OracleContext())
    assert all(p.is_point_at_infinity() for p in pk_inf.powers_2)


def test_verify_kzg_and_degree_bound_equations():
    pk = _pk(4, full_g2=True)
    f = mz.Polynomial([6, 11, 6, 1])
    c, pr = mz.commit_kzg(f, pk), mz.open_kzg(f, 5, pk)
    assert mz.verify_kzg(5, c, pr, pk)                                  # kzg.rs:152-175
    assert not mz.verify_kzg(5, c, mz.ProofKZG(pr.y + 1, pr.w), pk)
    dp = mz.prove_degree_bound(f, pk, 3)
    assert mz.verify_degree_bound(c, dp, pk, 3)                         # kzg.rs:207-233
    assert not mz.verify_degree_bound(c, dp, pk, 2)
    with pytest.raises(ValueError):
        mz.verify_degree_bound(c, dp, _pk(4), 1)  # needs powers_2[max_d - d]: only two G2 powers here


def test_batch_open_and_verify():
    pk = _pk(3, full_g2=True)
    f = mz.Polynomial([6, 11, 6, 1])
    c = mz.commit_kzg(f, pk)
    zs = [5, 7]
    bp = mz.batch_open_kzg(f, zs, pk)
    assert bp.ys == [336, 720]
    assert mz.batch_verify_kzg(zs, c, bp, pk)                            # kzg.rs:177-205
    bp.ys[0] += 1
    assert not mz.batch_verify_kzg(zs, c, bp, pk)
    with pytest.raises(ValueError):
        mz.batch_verify_kzg([1, 2], c, bp, _pk(3))  # Z has three coefficients, two G2 powers


def test_gemini_fold_open_verify_bookkeeping():
    with pytest.raises(mz.SplitFoldError):
        mz.split_and_fold_commit([1, 2, 3], [1, 2], _pk(8))               # gemini.rs:55-59
    with pytest.raises(mz.SplitFoldError):
        mz.split_and_fold_commit([1, 2, 3, 4], [1], _pk(8))               # gemini.rs:60-66
    pk = _pk(4, full_g2=True)
    coef, rhos = [1, 2, 3, 4], [2, 3]
    cms, polys = mz.split_and_fold_commit(coef, rhos, pk, want_folds=True)
    fs = [mz.Polynomial(coef)] + polys
    assert [p._wire() for p in fs] == o.fold_ints(coef, rhos)
    assert [c.as_tuple() for c in cms] == [c.as_tuple() for c in mz.commit_gemini(fs, pk)]
    mu = fs[-1]._wire()[0]
    assert mu == (1 + 2 * 2 + 3 * 3 + 4 * 6) % R  # sum coef_i * tensor(rhos)_i, gemini.rs:298-307
    proof = mz.open_gemini(fs, 1234, pk)
    assert len(proof.es) == 2 and len(proof.degree_proofs) == 3
    assert mz.verify_gemini(rhos, mu, 1234, cms, proof, pk)              # gemini.rs:288-328
    assert not mz.verify_gemini(rhos, mu + 1, 1234, cms, proof, pk)
    assert not mz.verify_gemini(rhos + [5], mu, 1234, cms, proof, pk)    # challenge count != commitments - 1

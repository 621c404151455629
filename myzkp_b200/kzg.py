"""Host-side mirror of the reference's KZG surface for the hot path.

Same names, argument meaning and error behaviour as
myzkp/src/modules/algebra/kzg.rs (setup_kzg :27-40, commit_kzg :57-59,
open_kzg :61-72) and polynomial.rs (Polynomial{coef} low->high, :69-74), with
the work done by libmyzkp_b200.so on the GPU.  Scalars are Python ints (the
reference's FqOrder values); they are sanitized to [0, r) before they cross
the ABI, exactly where the reference sanitizes (polynomial.rs:162).
"""
from __future__ import annotations

import secrets
from typing import List, Optional, Sequence

import numpy as np

from .context import Context, R_MOD, P_MOD


class G1Point:
    """EllipticCurvePoint<Fq, BN128Curve> (curve/curve.rs:17-46): affine; infinity = (None, None)."""

    __slots__ = ("x", "y")

    def __init__(self, x: Optional[int] = None, y: Optional[int] = None):
        self.x = None if x is None else int(x) % P_MOD
        self.y = None if y is None else int(y) % P_MOD

    @staticmethod
    def point_at_infinity() -> "G1Point":
        return G1Point(None, None)

    def is_point_at_infinity(self) -> bool:
        return self.x is None or self.y is None

    @staticmethod
    def _from_tuple(t) -> "G1Point":
        return G1Point(None, None) if t is None else G1Point(t[0], t[1])

    def as_tuple(self):
        return None if self.is_point_at_infinity() else (self.x, self.y)

    def __eq__(self, o):
        return isinstance(o, G1Point) and self.as_tuple() == o.as_tuple()

    def __repr__(self):
        return f"G1Point({self.as_tuple()})"


class G2Point:
    """EllipticCurvePoint<Fq2, BN128Curve> (curve/bn128.rs:49): affine over Fq2 = Fq[u]/(u^2 + 1);
    a coordinate is the pair (c0, c1) of the reference's polynomial coefficients; infinity = (None, None)."""

    __slots__ = ("x", "y")

    def __init__(self, x=None, y=None):
        self.x = None if x is None else (int(x[0]) % P_MOD, int(x[1]) % P_MOD)
        self.y = None if y is None else (int(y[0]) % P_MOD, int(y[1]) % P_MOD)

    @staticmethod
    def point_at_infinity() -> "G2Point":
        return G2Point(None, None)

    def is_point_at_infinity(self) -> bool:
        return self.x is None or self.y is None

    @staticmethod
    def _from_tuple(t) -> "G2Point":
        return G2Point(None, None) if t is None else G2Point(t[0], t[1])

    def as_tuple(self):
        return None if self.is_point_at_infinity() else (self.x, self.y)

    def __eq__(self, o):
        return isinstance(o, G2Point) and self.as_tuple() == o.as_tuple()

    def __repr__(self):
        return f"G2Point({self.as_tuple()})"


class BN128:
    """curve/bn128.rs:183-212."""

    @staticmethod
    def generator_g1() -> G1Point:
        return G1Point(1, 2)

    @staticmethod
    def generator_g2() -> G2Point:
        """bn128.rs:190-205."""
        return G2Point(
            (10857046999023057135944570762232829481370756359578518086990519993285655852781,
             11559732032986387107991004021392285783925812861821192530917403151452391805634),
            (8495653923123431417604973247489272438418190587263600148770280649306958101930,
             4082367875863433681332203403145435568316851327593401208105741076214120093531))

    @staticmethod
    def order() -> int:
        return R_MOD


class Polynomial:
    """Polynomial<FqOrder> (polynomial.rs:69-74): `coef` in increasing degree.

    `coef` may be a list of ints (any sign; sanitized like field.rs:260-270) or
    an (n,32) uint8 / (n,4) uint64 array of canonical little-endian scalars.
    """

    def __init__(self, coef):
        self.coef = coef

    def _wire(self):
        if isinstance(self.coef, np.ndarray):
            return self.coef
        return [int(c) % R_MOD for c in self.coef]  # sanitize (polynomial.rs:162)

    def __len__(self):
        return len(self.coef)


class PublicKeyKZG:
    """kzg.rs:8-11.  powers_1 lives on the GPU as the resident SRS table; powers_2 (G2, used by the
    verifier only) is computed on the GPU at setup and kept on the host."""

    def __init__(self, ctx: Context, powers_2: Optional[List[G2Point]] = None):
        self.ctx = ctx
        self.powers_2 = powers_2 if powers_2 is not None else []

    @property
    def powers_1(self) -> List[G1Point]:
        return [G1Point._from_tuple(t) for t in self.ctx.srs_read(0, self.ctx.srs_len)]

    def __len__(self):
        return self.ctx.srs_len


class ProofKZG:
    """kzg.rs:15-18."""

    def __init__(self, y: int, w: G1Point):
        self.y = y
        self.w = w


CommitmentKZG = G1Point


def _setup(g1: G1Point, g2: Optional[G2Point], max_d: int, n_g2: int, alpha, ctx, device) -> PublicKeyKZG:
    if g1 != BN128.generator_g1():
        raise ValueError("setup_kzg on the GPU path generates powers of BN128::generator_g1(); load other SRS with Context.srs_load")
    if alpha is None:
        alpha = secrets.randbelow(R_MOD)
    ctx = ctx or Context(device)
    ctx.srs_generate(alpha, max_d + 1)
    base = None if g2 is None or g2 == BN128.generator_g2() else g2.as_tuple()
    if g2 is not None and g2.is_point_at_infinity():
        p2 = [None] * n_g2
    else:
        p2 = ctx.srs_generate_g2(alpha, n_g2, 0, base)
    return PublicKeyKZG(ctx, [G2Point._from_tuple(t) for t in p2])


def setup_kzg(g1: G1Point, g2: Optional[G2Point] = None, max_d: int = 0, *, alpha: Optional[int] = None,
              ctx: Optional[Context] = None, device: int = 0) -> PublicKeyKZG:
    """kzg.rs:27-40: max_d + 1 powers [alpha^i]g1 and powers_2 = [g2, [alpha]g2].  The reference draws
    alpha from an unseeded thread_rng (field.rs:198-206); pass `alpha` for reproducibility.  g2 = None
    means BN128::generator_g2(); any G2 base point is accepted.  On the G1 side only the standard
    generator is supported on the device path; any other base is handled by loading explicit powers
    with Context.srs_load."""
    return _setup(g1, g2, max_d, 2, alpha, ctx, device)


def setup_kzg_with_full_g2(g1: G1Point, g2: Optional[G2Point] = None, max_d: int = 0, *, alpha: Optional[int] = None,
                           ctx: Optional[Context] = None, device: int = 0) -> PublicKeyKZG:
    """kzg.rs:42-55: as setup_kzg, with powers_2 = [alpha^i]g2 for i = 0..=max_d."""
    return _setup(g1, g2, max_d, max_d + 1, alpha, ctx, device)


def commit_kzg(f: Polynomial, pk: PublicKeyKZG) -> CommitmentKZG:
    """kzg.rs:57-59.  Raises (reference: index panic, polynomial.rs:162) if f is longer than the SRS."""
    return G1Point._from_tuple(pk.ctx.commit(f._wire()))


def open_kzg(f: Polynomial, u: int, pk: PublicKeyKZG) -> ProofKZG:
    """kzg.rs:61-72."""
    y, w = pk.ctx.open(f._wire(), int(u) % R_MOD)
    return ProofKZG(y, G1Point._from_tuple(w))


class BatchProofKZG:
    """kzg.rs:20-23."""

    def __init__(self, ys: List[int], w: G1Point):
        self.ys = ys
        self.w = w


ProofDegreeBound = G1Point


def batch_open_kzg(f: Polynomial, us: Sequence[int], pk: PublicKeyKZG) -> BatchProofKZG:
    """kzg.rs:74-88."""
    ys, w = pk.ctx.batch_open(f._wire(), [int(u) % R_MOD for u in us])
    return BatchProofKZG(ys, G1Point._from_tuple(w))


def prove_degree_bound(f: Polynomial, pk: PublicKeyKZG, d: int) -> ProofDegreeBound:
    """kzg.rs:121-134."""
    return G1Point._from_tuple(pk.ctx.prove_degree_bound(f._wire(), d))


def _neg_g1(t):
    return None if t is None else (t[0], (-t[1]) % P_MOD)


def verify_kzg(u: int, c: CommitmentKZG, proof: ProofKZG, pk: PublicKeyKZG) -> bool:
    """kzg.rs:90-102: e(C, g2) == e(W, [alpha]g2 - [u]g2) * e(g1, g2)^y, checked on the GPU as
    e(C, g2) * e(-W, [alpha - u]g2) * e([-y]g1, g2) == 1 (three Miller loops, one final exponentiation)."""
    ctx = pk.ctx
    u, y = int(u) % R_MOD, int(proof.y) % R_MOD
    g1 = ctx.srs_read(0, 1)[0]
    g2, g2_alpha = pk.powers_2[0].as_tuple(), pk.powers_2[1].as_tuple()
    g2_alpha_minus_u = ctx.g2_msm([1, (-u) % R_MOD], [g2_alpha, g2])
    g1_minus_y = ctx.g1_msm([(-y) % R_MOD], [g1])
    return ctx.pairing_product_is_one([c.as_tuple(), _neg_g1(proof.w.as_tuple()), g1_minus_y], [g2, g2_alpha_minus_u, g2])


def _interpolate(xs, ys):
    """Polynomial::interpolate (polynomial.rs:177-200) on k (= a handful of) points: Lagrange form, ints mod r."""
    k = len(xs)
    out = [0] * k
    for i in range(k):
        num, den = [1], 1
        for j in range(k):
            if j != i:
                num = [(a - xs[j] * b) % R_MOD for a, b in zip([0] + num, num + [0])]
                den = den * (xs[i] - xs[j]) % R_MOD
        scale = ys[i] * pow(den, -1, R_MOD) % R_MOD
        for t in range(len(num)):
            out[t] = (out[t] + scale * num[t]) % R_MOD
    return out


def _from_monomials(us):
    """Polynomial::from_monomials (polynomial.rs:202-212): prod (x - u_i)."""
    z = [1]
    for u in us:
        z = [(a - u * b) % R_MOD for a, b in zip([0] + z, z + [0])]
    return z


def batch_verify_kzg(us: Sequence[int], c: CommitmentKZG, proof: BatchProofKZG, pk: PublicKeyKZG) -> bool:
    """kzg.rs:104-119: e(W, [Z(alpha)]g2) == e(C - [I(alpha)]g1, g2) with I the interpolant of (us, ys) and
    Z = prod (x - u_i); needs len(us) + 1 powers of g2 (setup_kzg_with_full_g2 for more than one point)."""
    ctx = pk.ctx
    us = [int(u) % R_MOD for u in us]
    ip = _interpolate(us, [int(y) % R_MOD for y in proof.ys])
    z = _from_monomials(us)
    if len(z) > len(pk.powers_2):
        raise ValueError("batch_verify_kzg needs len(us) + 1 powers of g2 (reference: index panic, polynomial.rs:162)")
    g1_ip = ctx.commit(ip)
    g2_z = ctx.g2_msm(z, [p.as_tuple() for p in pk.powers_2[: len(z)]])
    lhs = ctx.g1_msm([1, R_MOD - 1], [c.as_tuple(), g1_ip])  # C - [I(alpha)]g1
    return ctx.pairing_product_is_one([proof.w.as_tuple(), _neg_g1(lhs)], [g2_z, pk.powers_2[0].as_tuple()])


def verify_degree_bound(c: CommitmentKZG, proof: ProofDegreeBound, pk: PublicKeyKZG, d: int) -> bool:
    """kzg.rs:136-144: e(proof, g2) == e(C, [alpha^(max_d - d)]g2); needs setup_kzg_with_full_g2."""
    max_d = len(pk) - 1
    if not 0 <= max_d - d < len(pk.powers_2):
        raise ValueError("verify_degree_bound needs powers_2[max_d - d] (reference: index panic)")
    return pk.ctx.pairing_product_is_one([proof.as_tuple(), _neg_g1(c.as_tuple())],
                                         [pk.powers_2[0].as_tuple(), pk.powers_2[max_d - d].as_tuple()])


def optimal_ate_pairing(p: G1Point, q: G2Point, pk: PublicKeyKZG) -> List[int]:
    """curve/bn128.rs:147-181 on the GPU: the Fq12 value as its 12 coefficients of w^k."""
    return pk.ctx.pairing([p.as_tuple()], [q.as_tuple()])[0]

"""CPU oracle for the MyZKP KZG prover hot path (BN128).

TEST INFRASTRUCTURE ONLY.  This module is the *checker*, never the product:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it.  The product path (myzkp_b200/) never does and
fails loudly when its CUDA library is missing.

It restates, line by line and on Python integers (== num-bigint BigInt
semantics; num-bigint = "0.4", myzkp/Cargo.toml:7, un-vendored, exact integer
arithmetic), the reference functions the hot path consists of.  Citations are
relative to /root/reference/myzkp/src/modules/algebra/.

Parity pinning: the reference cannot be compiled here (Rust toolchain absent)
and its own KZG tests are pairing-verified property tests with an unseeded
random alpha (kzg.rs:152-175), so commitment coordinates are *not* pinned by
the reference; what those tests check - verify_kzg accepts - is restated here too (optimal_ate_pairing,
verify_kzg) and holds for this oracle's own commitments and proofs.  What is pinned (tests/test_oracle_kats.py): every known-answer
test the reference holds under this path - bn128.rs:285-301 (test_g1),
bn128.rs:240-251 (test_fq), field.rs:443-550, polynomial.rs:727-768,805-821,
cuda/test_fr.cu:5-42, gemini.rs:288-307 / book gemini.md:311-328 - plus the
public EIP-196 value of 2G and the algebraic identity C == [f(alpha)]G.  The G2 / Fq2 / Fq12 / pairing
restatements (bn128.rs:33-181, efield.rs, curve.rs:285-339) are pinned by bn128.rs:254-365 (test_fq2, test_g2,
test_g12, test_pairing) and the public value of 2*G2.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

# --- constants: curve/bn128.rs:19-23, field.rs:428-431 ----------------------
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
CURVE_A = 0  # bn128.rs:23 define_myzkp_curve_type!(BN128Curve, "0", "3")
CURVE_B = 3


def _trunc_rem(a: int, m: int) -> int:
    """Rust/num-bigint `%`: truncated remainder, sign follows the dividend."""
    r = abs(a) % m
    return -r if a < 0 else r


def _trunc_div(a: int, b: int) -> int:
    """Rust/num-bigint `/`: quotient truncated toward zero."""
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def extended_euclidean(a: int, b: int) -> Tuple[int, int, int]:
    """utils.rs:52-81."""
    r0, r1 = a, b
    s0, s1 = 1, 0
    t0, t1 = 0, 1
    while r1 != 0:
        q = _trunc_div(r0, r1)
        r = _trunc_rem(r0, r1)
        r0, r1 = r1, r
        s0, s1 = s1, s0 - q * s1
        t0, t1 = t1, t0 - q * t1
    return r0, s0, t0


def mod_pow(x: int, y: int, modulus: int) -> int:
    """utils.rs:108-137 (non-negative exponent branch)."""
    assert y >= 0
    base = _trunc_rem(x, modulus)
    result = 1
    e = y
    while e > 0:
        if e & 1:
            result = _trunc_rem(result * base, modulus)
        e >>= 1
        base = _trunc_rem(base * base, modulus)
    return result


class FE:
    """FiniteFieldElement<M> (field.rs:87-91): `value` may be negative, exactly
    as in the reference (new() uses truncated `%`, field.rs:102-110)."""

    __slots__ = ("value",)
    MOD = 0

    def __init__(self, value: int):
        self.value = _trunc_rem(int(value), self.MOD)  # field.rs:104-105

    # Ring (field.rs:157-195)
    def add_ref(self, o: "FE") -> "FE":
        return type(self)(self.value + o.value)

    def sub_ref(self, o: "FE") -> "FE":
        return type(self)(self.value - o.value)

    def mul_ref(self, o: "FE") -> "FE":
        return type(self)(self.value * o.value)

    def pow(self, n: int) -> "FE":
        return type(self)(mod_pow(self.value, n, self.MOD))

    @classmethod
    def from_value(cls, v: int) -> "FE":
        return cls(v)

    def get_value(self) -> int:
        return self.value

    @classmethod
    def zero(cls) -> "FE":
        return cls(0)

    @classmethod
    def one(cls) -> "FE":
        return cls(1)

    # Field (field.rs:210-270)
    def inverse(self) -> "FE":
        modulus = self.MOD
        r = modulus
        new_r = self.value
        if new_r < 0:
            new_r += modulus
        _, _, t = extended_euclidean(r, new_r)
        t = _trunc_rem(t, modulus)
        if t < 0:
            t += modulus
        return type(self)(t)

    def div_ref(self, o: "FE") -> "FE":
        return self.mul_ref(o.inverse())

    def sanitize(self) -> "FE":
        v = _trunc_rem(self.value, self.MOD)
        if v < 0:
            v += self.MOD
        out = type(self).__new__(type(self))
        out.value = v
        return out

    # operators (field.rs:290-387)
    def __eq__(self, o) -> bool:  # field.rs:290-294: equality on sanitized values
        return isinstance(o, FE) and self.MOD == o.MOD and self.sanitize().value == o.sanitize().value

    def __hash__(self):
        return hash((self.MOD, self.sanitize().value))

    def __add__(self, o):
        return self.add_ref(o)

    def __sub__(self, o):
        return self.sub_ref(o)

    def __mul__(self, o):
        return self.mul_ref(o)

    def __neg__(self):  # field.rs:373-379
        return type(self)(0) - self

    def __truediv__(self, o):
        return self.div_ref(o)

    def __repr__(self):
        return f"{type(self).__name__}({self.value})"


class Fq(FE):  # bn128.rs:29 (base field, BN128Modulus)
    __slots__ = ()
    MOD = P_MOD


class Fr(FE):  # bn128.rs:30 FqOrder = FiniteFieldElement<ModEIP197>
    __slots__ = ()
    MOD = R_MOD


FqOrder = Fr


def make_field(modulus: int, name: str = "Fm"):
    """define_myzkp_modulus_type! (field.rs:408-425) for the small-modulus KATs."""
    return type(name, (FE,), {"MOD": modulus, "__slots__": ()})


class G1Point:
    """EllipticCurvePoint<Fq, BN128Curve> (curve/curve.rs:17-46): affine,
    infinity = (None, None), no on-curve check (curve.rs:26-28)."""

    __slots__ = ("x", "y")
    F = Fq  # coordinate field (EllipticCurvePoint is generic over it, curve.rs:17-24)

    def __init__(self, x: Optional[Fq], y: Optional[Fq]):
        self.x = x
        self.y = y

    @classmethod
    def new(cls, x, y):
        return cls(x, y)

    @classmethod
    def point_at_infinity(cls):
        return cls(None, None)

    def is_point_at_infinity(self) -> bool:
        return self.x is None or self.y is None

    def clone(self):
        return type(self)(self.x, self.y)

    def line_slope(self, other: "G1Point") -> Fq:
        """curve.rs:56-70."""
        F = self.F
        a = F.from_value(CURVE_A)
        x1, y1, x2, y2 = self.x, self.y, other.x, other.y
        if self.x == other.x:
            return (x1.mul_ref(x1) * F.from_value(3) + a) / (y1.mul_ref(F.from_value(2)))
        return y2.sub_ref(y1) / x2.sub_ref(x1)

    def double(self) -> "G1Point":
        """curve.rs:72-85."""
        if self.is_point_at_infinity():
            return self.clone()
        slope = self.line_slope(self)
        x, y = self.x, self.y
        new_x = slope.mul_ref(slope).sub_ref(x).sub_ref(x)
        new_y = -slope.mul_ref(new_x) + slope * x - y
        return type(self)(new_x, new_y)

    def inplace_double(self) -> None:
        """curve.rs:87-101."""
        d = self.double()
        self.x, self.y = d.x, d.y

    def add_ref(self, other: "G1Point") -> "G1Point":
        """curve.rs:103-128."""
        if self.is_point_at_infinity():
            return other.clone()
        if other.is_point_at_infinity():
            return self.clone()
        if self.x == other.x and self.y == other.y:
            return self.double()
        elif self.x == other.x:
            return type(self).point_at_infinity()
        slope = self.line_slope(other)
        x1, y1, x2 = self.x, self.y, other.x
        new_x = slope.mul_ref(slope).sub_ref(x1).sub_ref(x2)
        new_y = (-slope).mul_ref(new_x) + slope.mul_ref(x1).sub_ref(y1)
        return type(self)(new_x, new_y)

    def add_assign_ref(self, other: "G1Point") -> None:
        """curve.rs:130-161."""
        s = self.add_ref(other)
        self.x, self.y = s.x, s.y

    def mul_ref(self, scalar: int) -> "G1Point":
        """curve.rs:163-191: LSB-first double-and-add; 0 -> infinity; negative panics."""
        if scalar == 0:
            return type(self).point_at_infinity()
        if scalar < 0:
            raise ValueError("multiplier should be non-negative")  # curve.rs:174-176
        result = type(self).point_at_infinity()
        current = self.clone()
        bits = scalar
        while bits != 0:
            if bits & 1:
                result.add_assign_ref(current)
            current.inplace_double()
            bits >>= 1
        return result

    def __neg__(self):  # curve.rs:216-225
        if self.is_point_at_infinity():
            return self
        return type(self)(self.x, -self.y)

    def __add__(self, o):
        return self.add_ref(o)

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, k: int):
        return self.mul_ref(k)

    def __eq__(self, o) -> bool:
        if type(o) is not type(self):
            return False
        if self.is_point_at_infinity() or o.is_point_at_infinity():
            return self.is_point_at_infinity() and o.is_point_at_infinity()
        return self.x == o.x and self.y == o.y

    def affine_ints(self) -> Optional[Tuple[int, int]]:
        """Canonical (x mod p, y mod p) or None for infinity."""
        if self.is_point_at_infinity():
            return None
        return self.x.sanitize().value, self.y.sanitize().value

    def __repr__(self):
        return f"G1Point({self.affine_ints()})"


def generator_g1() -> G1Point:
    """bn128.rs:185-188."""
    return G1Point.new(Fq.from_value(1), Fq.from_value(2))


def order() -> int:
    """bn128.rs:207-212."""
    return R_MOD


class Polynomial:
    """Polynomial<F> (polynomial.rs:69-74): coefficients low -> high degree."""

    def __init__(self, coef: Sequence[FE], field=Fr):
        self.coef: List[FE] = list(coef)
        self.F = field if not self.coef else type(self.coef[0])

    @staticmethod
    def trim_trailing_zeros(coef: Sequence[FE]) -> List[FE]:
        """polynomial.rs:85-91 (the slice copy is the reference's O(d^2) source)."""
        end = 0
        for pos in range(len(coef) - 1, -1, -1):
            if coef[pos].sanitize().value != 0:
                end = pos + 1
                break
        return list(coef[:end])

    def degree(self) -> int:
        t = self.trim_trailing_zeros(self.coef)
        return len(t) - 1 if t else -1

    def is_zero(self) -> bool:
        return self.degree() == -1

    def eval(self, point: FE) -> FE:
        """polynomial.rs:120-128: running-power sum."""
        F = self.F
        result = F.zero()
        tp = F.one()
        for c in self.coef:
            result = result + tp.mul_ref(c)
            tp = tp.mul_ref(point)
        return result

    def eval_with_powers_on_curve(self, powers: Sequence[G1Point]) -> G1Point:
        """polynomial.rs:156-165: the naive MSM (index panics if too few powers)."""
        result = G1Point.point_at_infinity()
        for i, c in enumerate(self.coef):
            result.add_assign_ref(powers[i].mul_ref(c.sanitize().get_value()))
        return result

    @staticmethod
    def from_monomials(x_values: Sequence[FE], field=Fr) -> "Polynomial":
        """polynomial.rs:202-212."""
        F = type(x_values[0]) if x_values else field
        poly = Polynomial([F.one()], F)
        for x in x_values:
            poly = poly.mul_ref(Polynomial([F.zero() - x, F.one()], F))
        return poly

    def add_ref(self, other: "Polynomial") -> "Polynomial":
        """polynomial.rs:214-228."""
        F = self.F
        n = max(len(self.coef), len(other.coef))
        zero = F.zero()
        out = []
        for i in range(n):
            a = self.coef[i] if i < len(self.coef) else zero
            b = other.coef[i] if i < len(other.coef) else zero
            out.append(a.add_ref(b))
        return Polynomial(self.trim_trailing_zeros(out), F)

    def __neg__(self):  # polynomial.rs:465-473
        return Polynomial([-c for c in self.coef], self.F)

    def __sub__(self, other):  # polynomial.rs:517-523
        return self.add_ref(-other)

    def __add__(self, other):
        return self.add_ref(other)

    def mul_ref(self, other: "Polynomial") -> "Polynomial":
        """polynomial.rs:302-316."""
        F = self.F
        if self.is_zero() or other.is_zero():
            return Polynomial([], F)
        result = [F.zero() for _ in range(self.degree() + other.degree() + 1)]
        for i, a in enumerate(self.coef):
            for j, b in enumerate(other.coef):
                result[i + j] = result[i + j].add_ref(a.mul_ref(b))
        return Polynomial(self.trim_trailing_zeros(result), F)

    def scalar_mul(self, s: FE) -> "Polynomial":  # polynomial.rs:553-561
        return Polynomial([c.mul_ref(s) for c in self.coef], self.F)

    def div_rem_ref(self, other: "Polynomial") -> Tuple["Polynomial", "Polynomial"]:
        """polynomial.rs:371-405: schoolbook long division (trim per iteration)."""
        F = self.F
        if self.degree() < other.degree():
            return Polynomial([], F), Polynomial(self.coef, F)
        remainder = self.trim_trailing_zeros(self.coef)
        divisor = self.trim_trailing_zeros(other.coef)
        if len(divisor) == 0:
            return Polynomial([], F), Polynomial(self.coef, F)
        lead_inv = divisor[-1].inverse()
        quotient = [F.zero() for _ in range(self.degree() - other.degree() + 1)]
        while len(remainder) >= len(divisor):
            lead_term = remainder[-1].mul_ref(lead_inv)
            deg_diff = len(remainder) - len(divisor)
            quotient[deg_diff] = lead_term
            for i in range(len(divisor)):
                remainder[deg_diff + i] = remainder[deg_diff + i].sub_ref(lead_term.mul_ref(divisor[i]))
            remainder = self.trim_trailing_zeros(remainder)
        return Polynomial(self.trim_trailing_zeros(quotient), F), Polynomial(remainder, F)

    def __truediv__(self, other):  # polynomial.rs:583-597
        return self.div_rem_ref(other)[0]

    @staticmethod
    def interpolate(x_values: Sequence[FE], y_values: Sequence[FE]) -> "Polynomial":
        """polynomial.rs:177-200: Lagrange interpolation."""
        F = type(x_values[0])
        lagrange_polys = []
        numerators = Polynomial.from_monomials(x_values)
        for j in range(len(x_values)):
            denominator = F.one()
            for i in range(len(x_values)):
                if i != j:
                    denominator = denominator * x_values[j].sub_ref(x_values[i])
            cur_poly = numerators / Polynomial.from_monomials([x_values[j]]).scalar_mul(denominator)
            lagrange_polys.append(cur_poly)
        result = Polynomial([], F)
        for j, lp in enumerate(lagrange_polys):
            result = result + lp.scalar_mul(y_values[j])
        return result

    def canonical(self) -> List[int]:
        return [c.sanitize().value for c in self.coef]


# --- kzg.rs -----------------------------------------------------------------
class PublicKeyKZG:
    """kzg.rs:8-11 (powers_1; the G2 half is restated separately by setup_kzg_g2)."""

    def __init__(self, powers_1: List[G1Point]):
        self.powers_1 = powers_1


class ProofKZG:
    """kzg.rs:15-18."""

    def __init__(self, y: Fr, w: G1Point):
        self.y = y
        self.w = w


def setup_kzg(g1: G1Point, max_d: int, alpha: int) -> PublicKeyKZG:
    """kzg.rs:27-40 with the unseeded alpha (field.rs:198-206) injected:
    max_d + 1 points [alpha^i]G (kzg.rs:32)."""
    a = Fr.from_value(alpha)
    powers_1 = []
    alpha_power = Fr.one()
    for _ in range(1 + max_d):
        powers_1.append(g1.mul_ref(alpha_power.get_value()))
        alpha_power = alpha_power * a
    return PublicKeyKZG(powers_1)


def commit_kzg(f: Polynomial, pk: PublicKeyKZG) -> G1Point:
    """kzg.rs:57-59."""
    return f.eval_with_powers_on_curve(pk.powers_1)


def open_kzg(f: Polynomial, u: Fr, pk: PublicKeyKZG) -> ProofKZG:
    """kzg.rs:61-72."""
    y = f.eval(u)
    y_poly = Polynomial([y], Fr)
    f_u = (f - y_poly) / Polynomial.from_monomials([u])
    return ProofKZG(y, f_u.eval_with_powers_on_curve(pk.powers_1))


class BatchProofKZG:
    """kzg.rs:20-23."""

    def __init__(self, ys: List[Fr], w: G1Point):
        self.ys = ys
        self.w = w


def batch_open_kzg(f: Polynomial, us: Sequence[Fr], pk: PublicKeyKZG) -> BatchProofKZG:
    """kzg.rs:74-88."""
    ys = [f.eval(z) for z in us]
    ip = Polynomial.interpolate(us, ys)
    z = Polynomial.from_monomials(us)
    f_u = (f - ip) / z
    return BatchProofKZG(ys, f_u.eval_with_powers_on_curve(pk.powers_1))


def prove_degree_bound(f: Polynomial, pk: PublicKeyKZG, d: int) -> G1Point:
    """kzg.rs:121-134: commit to f * x^(max_d - d)."""
    max_d = len(pk.powers_1) - 1
    q_coef = [Fr.zero() for _ in range(max_d + 1 - d)]
    q_coef[max_d - d] = Fr.one()
    r = f.mul_ref(Polynomial(q_coef, Fr))
    return r.eval_with_powers_on_curve(pk.powers_1)


# --- gemini.rs --------------------------------------------------------------
class SplitFoldError(ValueError):
    """gemini.rs:15-32."""


def split_and_fold(coef: Sequence[FE], rhos: Sequence[FE]) -> List[Polynomial]:
    """gemini.rs:51-103: log2(n)+1 polynomials incl. the original."""
    n = len(coef)
    if bin(n).count("1") != 1:
        raise SplitFoldError(f"coefs.len() must be a power of two, but got {n}")
    log2_n = n.bit_length() - 1
    if len(rhos) != log2_n:
        raise SplitFoldError(f"points.len() must be {log2_n}, but got {len(rhos)}")
    F = type(coef[0])
    f = Polynomial(list(coef), F)
    fs = [Polynomial(list(coef), F)]
    for i in range(1, log2_n + 1):
        f_e = Polynomial([x if k % 2 == 0 else F.zero() for k, x in enumerate(f.coef)], F)
        f_o = Polynomial([x if k % 2 == 0 else F.zero() for k, x in enumerate(f.coef[1:])], F)
        f_i = f_e + f_o.scalar_mul(rhos[i - 1])
        # gemini.rs:88-96 keeps the even positions of f_i.  f_i has been trimmed
        # by add_ref, so a fold with trailing zero coefficients comes out short;
        # for the commitment that is immaterial.
        f = Polynomial([x for k, x in enumerate(f_i.coef) if k % 2 == 0], F)
        fs.append(Polynomial(list(f.coef), F))
    return fs


def commit_gemini(polys: Sequence[Polynomial], pk: PublicKeyKZG) -> List[G1Point]:
    """gemini.rs:112-114."""
    return [commit_kzg(p, pk) for p in polys]


class ProofGemini:
    """gemini.rs:107-110."""

    def __init__(self, es: List[BatchProofKZG], degree_proofs: List[G1Point]):
        self.es = es
        self.degree_proofs = degree_proofs


def open_gemini(polys: Sequence[Polynomial], beta: Fr, pk: PublicKeyKZG) -> ProofGemini:
    """gemini.rs:116-144."""
    num_polys = len(polys)
    es = [batch_open_kzg(p, [beta, (Fr.zero() - beta).sanitize(), beta.pow(2)], pk) for p in polys[: num_polys - 1]]
    degree_proofs = [prove_degree_bound(p, pk, 2 ** (num_polys - i - 1)) for i, p in enumerate(polys)]
    return ProofGemini(es, degree_proofs)

# --- G2: the same curve law over Fq2 = Fq[x] / (x^2 + 1) ----------------------
class Fq2:
    """ExtendedFieldElement<BN128Modulus, Fq2Poly> (efield.rs:93-118, bn128.rs:33-49): a polynomial
    over Fq kept reduced modulo x^2 + 1; `c` = [c0, c1] low -> high (canonical ints, always length 2
    here; the reference trims trailing zeros, which does not change the value)."""

    __slots__ = ("c",)

    def __init__(self, c: Sequence[int]):
        # efield.rs:103-108: poly % (x^2 + 1), i.e. x^2 = -1, x^3 = -x, x^4 = 1, ...
        lo = [0, 0]
        for i, v in enumerate(c):
            lo[i % 2] += -int(v) if (i // 2) % 2 else int(v)
        lo = [lo[0] % P_MOD, lo[1] % P_MOD]
        self.c = lo

    @classmethod
    def from_value(cls, v: int) -> "Fq2":  # from_base_field, efield.rs:115-117
        return cls([v])

    @classmethod
    def zero(cls):
        return cls([0])

    @classmethod
    def one(cls):
        return cls([1])

    def is_zero(self) -> bool:
        return self.c == [0, 0]

    def add_ref(self, o: "Fq2") -> "Fq2":  # efield.rs:342-344
        return Fq2([self.c[0] + o.c[0], self.c[1] + o.c[1]])

    def sub_ref(self, o: "Fq2") -> "Fq2":  # efield.rs:360-362
        return Fq2([self.c[0] - o.c[0], self.c[1] - o.c[1]])

    def mul_ref(self, o: "Fq2") -> "Fq2":  # efield.rs:351-353: polynomial product, then new() reduces
        a, b = self.c, o.c
        return Fq2([a[0] * b[0], a[0] * b[1] + a[1] * b[0], a[1] * b[1]])

    def inverse(self) -> "Fq2":
        """efield.rs:126-151: extended Euclid on polynomials (low, high) = (self, x^2 + 1)."""
        lm, hm = [1], [0]
        low, high = _ptrim(list(self.c)), [1, 0, 1]
        while low:
            q, r = _pdivmod(high, low)
            nm = _psub(hm, _pmul(lm, q))
            high, hm, low, lm = low, lm, r, nm
        inv0 = Fq.from_value(high[0]).inverse().sanitize().value
        return Fq2([v * inv0 for v in hm])

    def div_ref(self, o: "Fq2") -> "Fq2":  # efield.rs:153-155
        return self.mul_ref(o.inverse())

    def __eq__(self, o) -> bool:  # efield.rs:199-205
        return isinstance(o, Fq2) and self.c == o.c

    def __hash__(self):
        return hash(tuple(self.c))

    def __add__(self, o):
        return self.add_ref(o)

    def __sub__(self, o):
        return self.sub_ref(o)

    def __mul__(self, o):
        return self.mul_ref(o)

    def __neg__(self):  # efield.rs:329-337
        return Fq2([-self.c[0], -self.c[1]])

    def __truediv__(self, o):
        return self.div_ref(o)

    def sanitize(self) -> "Fq2":
        return self

    def __repr__(self):
        return f"Fq2({self.c})"


def _ptrim(a: List[int]) -> List[int]:
    a = [v % P_MOD for v in a]
    while a and a[-1] == 0:
        a.pop()
    return a


def _pmul(a: List[int], b: List[int]) -> List[int]:
    out = [0] * max(0, len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % P_MOD
    return _ptrim(out)


def _psub(a: List[int], b: List[int]) -> List[int]:
    n = max(len(a), len(b))
    return _ptrim([(a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0) for i in range(n)])


def _pdivmod(a: List[int], b: List[int]) -> Tuple[List[int], List[int]]:
    """Polynomial long division over Fq (polynomial.rs:371-405)."""
    a, b = _ptrim(list(a)), _ptrim(list(b))
    if len(a) < len(b):
        return [], a
    q = [0] * (len(a) - len(b) + 1)
    lead_inv = Fq.from_value(b[-1]).inverse().sanitize().value
    r = list(a)
    while len(r) >= len(b) and r:
        shift = len(r) - len(b)
        coef = r[-1] * lead_inv % P_MOD
        q[shift] = coef
        for i, v in enumerate(b):
            r[shift + i] = (r[shift + i] - coef * v) % P_MOD
        r = _ptrim(r)
    return _ptrim(q), r


class G2Point(G1Point):
    """EllipticCurvePoint<Fq2, BN128Curve> (bn128.rs:49): curve.rs's affine law with a = 0 over Fq2."""

    __slots__ = ()
    F = Fq2

    def affine_ints(self):
        """((x.c0, x.c1), (y.c0, y.c1)) canonical, or None for infinity."""
        if self.is_point_at_infinity():
            return None
        return (tuple(self.x.c), tuple(self.y.c))

    def __repr__(self):
        return f"G2Point({self.affine_ints()})"


G2_GEN_X = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634)
G2_GEN_Y = (8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531)


def generator_g2() -> G2Point:
    """bn128.rs:190-205."""
    return G2Point.new(Fq2(G2_GEN_X), Fq2(G2_GEN_Y))


def get_b2() -> Fq2:
    """bn128.rs:218-224: 3 / (9 + x)."""
    return Fq2([3]) / Fq2([9, 1])


def setup_kzg_g2(g2: G2Point, alpha: int, n: int) -> List[G2Point]:
    """powers_2 of kzg.rs:27-55 with alpha injected: n = 2 for setup_kzg ([g2, [alpha]g2], kzg.rs:37),
    n = max_d + 1 for setup_kzg_with_full_g2 (kzg.rs:47-52)."""
    a = Fr.from_value(alpha)
    out, alpha_power = [], Fr.one()
    for _ in range(n):
        out.append(g2.mul_ref(alpha_power.sanitize().get_value()))
        alpha_power = alpha_power * a
    return out


# --- pairing: Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), G12 points, Miller loop -------------------
FQ12_MODULUS = [82, 0, 0, 0, 0, 0, -18 % P_MOD, 0, 0, 0, 0, 0, 1]  # bn128.rs:55-72
ATE_LOOP_COUNT = 29793968203157093288  # bn128.rs:26


class Fq12:
    """ExtendedFieldElement<BN128Modulus, Fq12Poly> (bn128.rs:51-80, efield.rs): polynomial over Fq modulo
    w^12 - 18 w^6 + 82; `c` = 12 canonical ints, low -> high."""

    __slots__ = ("c",)

    def __init__(self, c: Sequence[int]):
        _, r = _pdivmod([int(v) % P_MOD for v in c], FQ12_MODULUS)  # efield.rs:103-108
        self.c = (r + [0] * 12)[:12]

    @classmethod
    def from_value(cls, v: int) -> "Fq12":
        return cls([v])

    @classmethod
    def zero(cls):
        return cls([0])

    @classmethod
    def one(cls):
        return cls([1])

    def is_zero(self) -> bool:
        return not any(self.c)

    def add_ref(self, o):
        return Fq12([a + b for a, b in zip(self.c, o.c)])

    def sub_ref(self, o):
        return Fq12([a - b for a, b in zip(self.c, o.c)])

    def mul_ref(self, o):  # efield.rs:351-353
        return Fq12(_pmul(self.c, o.c))

    def inverse(self):
        """efield.rs:126-151."""
        lm, hm = [1], [0]
        low, high = _ptrim(list(self.c)), list(FQ12_MODULUS)
        while low:
            q, r = _pdivmod(high, low)
            nm = _psub(hm, _pmul(lm, q))
            high, hm, low, lm = low, lm, r, nm
        inv0 = Fq.from_value(high[0]).inverse().sanitize().value
        return Fq12([v * inv0 for v in hm])

    def div_ref(self, o):
        return self.mul_ref(o.inverse())

    def pow(self, n: int):
        """Square-and-multiply; the result does not depend on the order of the multiplications."""
        result, base = Fq12.one(), self
        while n > 0:
            if n & 1:
                result = result.mul_ref(base)
            base = base.mul_ref(base)
            n >>= 1
        return result

    def __eq__(self, o):
        return isinstance(o, Fq12) and self.c == o.c

    def __hash__(self):
        return hash(tuple(self.c))

    def __add__(self, o):
        return self.add_ref(o)

    def __sub__(self, o):
        return self.sub_ref(o)

    def __mul__(self, o):
        return self.mul_ref(o)

    def __neg__(self):
        return Fq12([-v for v in self.c])

    def __truediv__(self, o):
        return self.div_ref(o)

    def sanitize(self):
        return self

    def __repr__(self):
        return f"Fq12({self.c})"


class G12Point(G1Point):
    """EllipticCurvePoint<Fq12, BN128Curve> (bn128.rs:81)."""

    __slots__ = ()
    F = Fq12


def cast_g1_to_g12(g: G1Point) -> G12Point:
    """bn128.rs:83-96."""
    if g.is_point_at_infinity():
        return G12Point.point_at_infinity()
    return G12Point.new(Fq12([g.x.sanitize().value]), Fq12([g.y.sanitize().value]))


def twist_g2_to_g12(g: G2Point) -> G12Point:
    """bn128.rs:98-145: u -> w^6 - 9, then (x w^2, y w^3)."""
    if g.is_point_at_infinity():
        return G12Point.point_at_infinity()
    w = Fq12([0, 1])
    x, y = g.x.c, g.y.c
    nx = Fq12([x[0] - 9 * x[1], 0, 0, 0, 0, 0, x[1]])
    ny = Fq12([y[0] - 9 * y[1], 0, 0, 0, 0, 0, y[1]])
    return G12Point.new(nx * w.pow(2), ny * w.pow(3))


def get_lambda(p, q, r):
    """curve.rs:285-311: the line through p and q (tangent if equal) over the vertical through p + q, at r."""
    F = type(p).F
    if p.is_point_at_infinity() or q.is_point_at_infinity() or r.is_point_at_infinity():
        return F.one()
    if (p == q and p.y == F.zero()) or (p != q and p.x == q.x):
        return r.x.sub_ref(p.x)
    slope = p.line_slope(q)
    numerator = r.y.sub_ref(p.y).sub_ref(slope.mul_ref(r.x.sub_ref(p.x)))
    denominator = r.x.add_ref(p.x).add_ref(q.x).sub_ref(slope.mul_ref(slope))
    return numerator / denominator


def miller(p, q, m: int):
    """curve.rs:313-339."""
    F = type(p).F
    if p.is_point_at_infinity() or q.is_point_at_infinity():
        return F.one(), type(p).point_at_infinity()
    if p == q:
        return F.one(), p.clone()
    f, t = F.one(), p.clone()
    for i in reversed(range(m.bit_length() - 1)):
        f = f.mul_ref(f) * get_lambda(t, t, q)
        t = t.add_ref(t)
        if (m >> i) & 1:
            f = f * get_lambda(t, p, q)
            t = t.add_ref(p)
    return f, t


FINAL_EXPONENT = (P_MOD ** 12 - 1) // R_MOD


def optimal_ate_pairing(p_g1: G1Point, q_g2: G2Point) -> Fq12:
    """bn128.rs:147-181."""
    p = cast_g1_to_g12(p_g1)
    q = twist_g2_to_g12(q_g2)
    if p.is_point_at_infinity() or q.is_point_at_infinity():
        return Fq12.one()
    f = Fq12.one()
    if p != q:
        f, r = miller(q, p, ATE_LOOP_COUNT)
        q1 = G12Point.new(q.x.pow(P_MOD), q.y.pow(P_MOD))
        nq2 = G12Point.new(q1.x.pow(P_MOD), -q1.y.pow(P_MOD))
        f = f * get_lambda(r, q1, p)
        r = r.add_ref(q1)
        f = f * get_lambda(r, nq2, p)
    return f.pow(FINAL_EXPONENT)


def verify_kzg(u: Fr, c: G1Point, proof: "ProofKZG", powers_1: Sequence[G1Point], powers_2: Sequence[G2Point]) -> bool:
    """kzg.rs:90-102."""
    g1, g2, g2_alpha = powers_1[0], powers_2[0], powers_2[1]
    g2_u = g2.mul_ref(u.sanitize().get_value())
    g2_alpha_minus_u = g2_alpha - g2_u
    e1 = optimal_ate_pairing(proof.w, g2_alpha_minus_u)
    e2 = optimal_ate_pairing(g1, g2)
    e3 = optimal_ate_pairing(c, g2)
    return e3 == e1 * e2.pow(proof.y.sanitize().get_value())


def verify_degree_bound(c: G1Point, proof: G1Point, powers_1: Sequence[G1Point], powers_2: Sequence[G2Point], d: int) -> bool:
    """kzg.rs:136-144 (needs setup_kzg_with_full_g2)."""
    max_d = len(powers_1) - 1
    return optimal_ate_pairing(proof, powers_2[0]) == optimal_ate_pairing(c, powers_2[max_d - d])


# fast independent path for G2 (Jacobian over Fq2 as int pairs) - used for larger checks
def _f2mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P_MOD, (a[0] * b[1] + a[1] * b[0]) % P_MOD)


def _f2inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P_MOD)
    return (a[0] * n % P_MOD, -a[1] * n % P_MOD)


def _g2_fast_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        if (y1[0] + y2[0]) % P_MOD == 0 and (y1[1] + y2[1]) % P_MOD == 0:
            return None
        xx = _f2mul(x1, x1)
        lam = _f2mul((3 * xx[0] % P_MOD, 3 * xx[1] % P_MOD), _f2inv((2 * y1[0] % P_MOD, 2 * y1[1] % P_MOD)))
    else:
        lam = _f2mul(((y2[0] - y1[0]) % P_MOD, (y2[1] - y1[1]) % P_MOD), _f2inv(((x2[0] - x1[0]) % P_MOD, (x2[1] - x1[1]) % P_MOD)))
    l2 = _f2mul(lam, lam)
    x3 = ((l2[0] - x1[0] - x2[0]) % P_MOD, (l2[1] - x1[1] - x2[1]) % P_MOD)
    t = _f2mul(lam, ((x1[0] - x3[0]) % P_MOD, (x1[1] - x3[1]) % P_MOD))
    return (x3, ((t[0] - y1[0]) % P_MOD, (t[1] - y1[1]) % P_MOD))


def g2_fast_mul(k: int, pt=(G2_GEN_X, G2_GEN_Y)):
    """[k mod r] pt on G2 as ((x0, x1), (y0, y1)) or None."""
    k %= R_MOD
    acc, cur = None, pt
    while k:
        if k & 1:
            acc = _g2_fast_add(acc, cur)
        cur = _g2_fast_add(cur, cur)
        k >>= 1
    return acc


def g2_to_bytes(pt) -> bytes:
    """Wire form: x.c0 | x.c1 | y.c0 | y.c1, 32 B little-endian each; infinity = 128 zero bytes."""
    if pt is None:
        return bytes(128)
    (x0, x1), (y0, y1) = pt
    return b"".join(int(v).to_bytes(32, "little") for v in (x0, x1, y0, y1))


def g2_from_bytes(b: bytes):
    v = [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(4)]
    return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))



# --- O(N) algebraic expected values for sizes the naive path cannot reach ---
def _pinv(v: int, m: int) -> int:
    return pow(v, -1, m)


def _fast_add(p1, p2):
    """Affine add on plain ints (None = infinity); same law as curve.rs:103-128,
    using pow(.,-1,p) instead of ext-Euclid.  Used only for expected values."""
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 != y2 or y1 == 0:
            return None
        s = 3 * x1 * x1 * _pinv(2 * y1, P_MOD) % P_MOD
    else:
        s = (y2 - y1) * _pinv(x2 - x1, P_MOD) % P_MOD
    x3 = (s * s - x1 - x2) % P_MOD
    return x3, (s * (x1 - x3) - y1) % P_MOD


def fast_mul(k: int, pt=(1, 2)):
    """[k]P on plain ints (MSB-first); validated against G1Point.mul_ref in tests."""
    k %= R_MOD
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = _fast_add(acc, acc)
        if bit == "1":
            acc = _fast_add(acc, pt)
    return acc


def expected_commit(coefs: Sequence[int], alpha: int):
    """C = [f(alpha) mod r]G for SRS [alpha^i]G (SURVEY 8c large-N oracle)."""
    acc = 0
    for c in reversed(coefs):
        acc = (acc * alpha + c) % R_MOD
    return fast_mul(acc)


def synthetic_division(coefs: Sequence[int], u: int) -> Tuple[int, List[int]]:
    """y = f(u) and q = (f - y)/(x - u), canonical ints; equals polynomial.rs:371-405
    for the monic linear divisor (checked against div_rem_ref in tests)."""
    n = len(coefs)
    if n == 0:
        return 0, []
    q = [0] * (n - 1)
    c = 0
    for i in range(n - 1, 0, -1):
        c = (coefs[i] + u * c) % R_MOD
        q[i - 1] = c
    y = (coefs[0] + u * c) % R_MOD
    return y, q


def expected_open(coefs: Sequence[int], u: int, alpha: int):
    """(y, W) with W = [(f(alpha) - y)/(alpha - u)]G, valid for alpha != u."""
    y, _ = synthetic_division(coefs, u)
    fa = 0
    for c in reversed(coefs):
        fa = (fa * alpha + c) % R_MOD
    k = (fa - y) * _pinv((alpha - u) % R_MOD, R_MOD) % R_MOD
    return y, fast_mul(k)


def fold_ints(coefs: Sequence[int], rhos: Sequence[int]) -> List[List[int]]:
    """Canonical-int form of split_and_fold (no trimming)."""
    out = [list(c % R_MOD for c in coefs)]
    cur = out[0]
    for rho in rhos:
        cur = [(cur[2 * k] + rho * cur[2 * k + 1]) % R_MOD for k in range(len(cur) // 2)]
        out.append(cur)
    return out


# --- wire encodings (examples/sumcheck/src/utils.rs:51-72) --------
def fe_to_bytes(v: int) -> bytes:
    return int(v).to_bytes(32, "little")


def fe_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little")


def point_to_bytes(pt) -> bytes:
    """affine (x, y) ints or None -> 64 B (infinity = 64 zero bytes)."""
    if pt is None:
        return bytes(64)
    return fe_to_bytes(pt[0]) + fe_to_bytes(pt[1])


def point_from_bytes(b: bytes):
    if b == bytes(64):
        return None
    return fe_from_bytes(b[:32]), fe_from_bytes(b[32:64])


def expected_batch_open(coefs: Sequence[int], us: Sequence[int], alpha: int):
    """(ys, W): the quotient of f by prod(x - u_i) is k successive synthetic divisions
    (floor division composes); W = [q_k(alpha)]G."""
    ys = [synthetic_division(coefs, u)[0] for u in us]
    cur = list(coefs)
    for u in us:
        cur = synthetic_division(cur, u)[1] if len(cur) else []
    qa = 0
    for c in reversed(cur):
        qa = (qa * alpha + c) % R_MOD
    return ys, fast_mul(qa)


def expected_degree_bound(coefs: Sequence[int], alpha: int, max_d: int, d: int):
    """[f(alpha) * alpha^(max_d - d)]G."""
    fa = 0
    for c in reversed(coefs):
        fa = (fa * alpha + c) % R_MOD
    return fast_mul(fa * pow(alpha, max_d - d, R_MOD) % R_MOD)

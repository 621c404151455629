"""CPU: the device field / group-law headers, compiled for the host with an
emulated carry flag (tests/emul/), against the oracle.  Validates the exact
algorithms the kernels run before any GPU time is spent."""
import ctypes
import os
import random
import subprocess

import pytest

import myzkp_oracle as o
from myzkp_oracle import _fast_add

HERE = os.path.dirname(os.path.abspath(__file__))
P, R = o.P_MOD, o.R_MOD
RR = 1 << 256


def _build(name):
    src = os.path.join(HERE, "emul", f"{name}.cpp")
    out = os.path.join(HERE, "emul", f"lib{name}.so")
    hdrs = [os.path.join(HERE, "..", "myzkp_b200", "csrc", h) for h in ("field.cuh", "g1.cuh", "g2.cuh", "pairing.cuh", "inv.cuh", "baa.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(x) > os.path.getmtime(out) for x in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


A8 = ctypes.c_uint32 * 8


def tl(x):
    return A8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def fl(a, off=0):
    return sum(int(a[off + i]) << (32 * i) for i in range(8))


@pytest.mark.parametrize("name,m", [("fq", P), ("fr", R)])
def test_field_ops(name, m):
    L = _build("emul_field")
    rnd = random.Random(1)
    rinv = pow(RR, -1, m)
    edge = [0, 1, 2, m - 1, m - 2, RR % m, (1 << 255) % m, m // 2, 0xFFFFFFFF, (1 << 224) - 1]
    vals = edge + [rnd.randrange(m) for _ in range(120)]

    def b2(fn, x, y):
        out = A8()
        getattr(L, f"emul_{name}_{fn}")(tl(x), tl(y), out)
        return fl(out)

    def u1(fn, x):
        out = A8()
        getattr(L, f"emul_{name}_{fn}")(tl(x), out)
        return fl(out)

    for x in vals:
        for y in rnd.sample(vals, 8) + edge:
            assert b2("mul", x, y) == x * y * rinv % m
            assert b2("mul_k", x, y) == x * y * rinv % m
            assert b2("add", x, y) == (x + y) % m
            assert b2("sub", x, y) == (x - y) % m
        assert u1("neg", x) == (-x) % m
        assert u1("sqr", x) == x * x * rinv % m
        assert u1("to_mont", x) == x * RR % m
        assert u1("from_mont", x) == x * rinv % m
    # the dedicated squaring: limb patterns that maximise carries, and a long random run
    limbs = [0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF]
    for _ in range(3000):
        x = sum(rnd.choice(limbs) << (32 * i) for i in range(8)) % m if rnd.random() < 0.5 else rnd.randrange(m)
        assert u1("sqr", x) == x * x * rinv % m
    # sum of two products with one reduction (bounds: all-maximal operands, carry-heavy limbs)
    def q4(fn, x, y, z, w):
        out = A8()
        getattr(L, f"emul_{name}_{fn}")(tl(x), tl(y), tl(z), tl(w), out)
        return fl(out)

    assert q4("mul2", m - 1, m - 1, m - 1, m - 1) == 2 * (m - 1) * (m - 1) * rinv % m
    for _ in range(4000):
        x, y, z, w = [(sum(rnd.choice(limbs) << (32 * i) for i in range(8)) % m if rnd.random() < 0.4
                       else rnd.choice(edge) if rnd.random() < 0.2 else rnd.randrange(m)) for _ in range(4)]
        assert q4("mul2", x, y, z, w) == (x * y + z * w) * rinv % m
        assert q4("mul_sub_mul", x, y, z, w) == (x * y - z * w) * rinv % m
        assert q4("mul2_k", x, y, z, w) == (x * y + z * w) * rinv % m
        assert b2("mul_k", x, y) == x * y * rinv % m
        assert b2("mul_k", z, w) == z * w * rinv % m
    for x in vals[:24]:
        exp = (pow(x, -1, m) if x else 0) * RR % m
        assert u1("inv", x * RR % m) == exp
        assert u1("inv_bingcd", x * RR % m) == exp
        assert u1("inv_safegcd", x * RR % m) == exp
    for x in [1, 2, m - 1, m - 2, (m + 1) // 2, 3, 1 << 200] + [rnd.randrange(1, m) for _ in range(300)]:
        assert u1("inv_bingcd", x * RR % m) == pow(x, -1, m) * RR % m
    # the branch-free safegcd inverse (csrc/inv.cuh): the Montgomery representative can be ANY residue, so
    # feed raw residues of every shape (small, near the modulus, single bits, sparse) and many random ones
    shapes = [1, 2, 3, m - 1, m - 2, (m + 1) // 2, (m - 1) // 2] + [1 << k for k in range(0, 254, 7)] + [m - (1 << k) for k in range(1, 250, 11)]
    for x in shapes + [rnd.randrange(1, m) for _ in range(3000)]:
        assert u1("inv_safegcd", x) == pow(x * pow(RR, -1, m) % m, -1, m) * RR % m



def test_group_law_special_cases():
    L = _build("emul_g1")
    rnd = random.Random(7)
    rinv = pow(RR, -1, P)

    def limbs(x):
        return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]

    def aff(pt):
        if pt is None:
            return (ctypes.c_uint32 * 16)()
        return (ctypes.c_uint32 * 16)(*(limbs(pt[0] * RR % P) + limbs(pt[1] * RR % P)))

    def xyzz(pt):
        if pt is None:
            return (ctypes.c_uint32 * 32)()
        z = rnd.randrange(1, P)
        zz, zzz = z * z % P, z * z * z % P
        out = []
        for v in (pt[0] * zz % P, pt[1] * zzz % P, zz, zzz):
            out += limbs(v * RR % P)
        return (ctypes.c_uint32 * 32)(*out)

    def to_aff(x):
        out = (ctypes.c_uint32 * 16)()
        L.emul_g1_to_affine(x, out)
        xs, ys = fl(out) * rinv % P, fl(out, 8) * rinv % P
        return None if (xs, ys) == (0, 0) else (xs, ys)

    pts = [o.fast_mul(rnd.randrange(1, R)) for _ in range(8)] + [o.fast_mul(1), o.fast_mul(2), None]
    neg = lambda p: None if p is None else (p[0], (-p[1]) % P)
    cases = [(a, b) for a in pts for b in pts] + [(a, neg(a)) for a in pts]
    for a, b in cases:
        exp = _fast_add(a, b)
        acc = xyzz(a)
        L.emul_g1_madd(acc, aff(b))
        assert to_aff(acc) == exp
        acc = xyzz(a)
        L.emul_g1_add(acc, xyzz(b))
        assert to_aff(acc) == exp
    for a in pts:
        acc = xyzz(a)
        L.emul_g1_dbl(acc)
        assert to_aff(acc) == _fast_add(a, a)


def test_batched_affine_rounds_emulated():
    """The BAA per-thread bodies (csrc/baa.cuh) on the CPU: bucket sums for random sorted entry
    lists incl. infinity points, duplicate points (doubling), P/-P pairs, odd runs, sentinels."""
    L = _build("emul_baa")
    rnd = random.Random(11)
    base = [o.fast_mul(rnd.randrange(1, R)) for _ in range(10)]
    table = base + [base[0], base[1], None, base[2], (base[3][0], P - base[3][1]), None]  # dup, inf, negation
    def limbs(x):
        return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    tb = []
    for pt in table:
        tb += ([0] * 16) if pt is None else (limbs(pt[0] * RR % P) + limbs(pt[1] * RR % P))
    tbl = (ctypes.c_uint32 * len(tb))(*tb)
    rinv = pow(RR, -1, P)
    for trial in range(12):
        nb = rnd.choice([1, 2, 5, 9])
        m_valid = rnd.randrange(1, 90) if trial != 5 else 330  # trial 5 spans several in-thread batches
        n_sent = rnd.randrange(0, 5)
        ents = sorted((rnd.randrange(nb), rnd.randrange(len(table)) | (rnd.randrange(2) << 31)) for _ in range(m_valid))
        if trial % 3 == 0:  # a long run of the same point, forces repeated doubling / cancellation
            ents = sorted(ents + [(0, 10 | (rnd.randrange(2) << 31)) for _ in range(9)])
        keys = [k for k, _ in ents] + [nb] * n_sent
        vals = [v for _, v in ents] + [0] * n_sent
        M = len(keys)
        exp = [None] * nb
        for k, v in ents:
            pt = table[v & 0x7FFFFFFF]
            if pt is not None and (v >> 31):
                pt = (pt[0], (P - pt[1]) % P)
            exp[k] = _fast_add(exp[k], pt)
        for seg in (1, 2, 3, 7, 16, 200, 400):
            for rounds, fused in ((1, 0), (2, 0), (3, 0), (8, 0)):
                out = (ctypes.c_uint32 * (16 * nb))()
                L.emul_baa_buckets((ctypes.c_uint32 * M)(*keys), (ctypes.c_uint32 * M)(*vals), M, seg, nb, tbl, rounds, nb, out, fused)
                got = []
                for b in range(nb):
                    x, y = fl(out, 16 * b) * rinv % P, fl(out, 16 * b + 8) * rinv % P
                    got.append(None if (x, y) == (0, 0) else (x, y))
                assert got == exp, (trial, seg, rounds, fused)


def test_g2_group_law_emulated():
    """csrc/g2.cuh (Fq2, Jacobian G2, scalar multiplication) against the oracle's restatement of the
    reference's affine law over Fq2 (curve.rs:56-191 with efield.rs) and its fast cross-check."""
    L = _build("emul_g2")
    rnd = random.Random(11)
    A32 = ctypes.c_uint32 * 32

    def enc(pt):
        if pt is None:
            return A32()
        vals = [pt[0][0], pt[0][1], pt[1][0], pt[1][1]]
        return A32(*[(v >> (32 * i)) & 0xFFFFFFFF for v in vals for i in range(8)])

    def dec(a):
        v = [fl(a, 8 * i) for i in range(4)]
        return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))

    gen = (o.G2_GEN_X, o.G2_GEN_Y)
    g = o.generator_g2()
    for k in [0, 1, 2, 3, 5, 12345, R - 1, R - 2, R, R + 7, (1 << 254) - 1, (1 << 256) - 1] + [rnd.randrange(R) for _ in range(6)]:
        out = A32()
        L.emul_g2_scalar_mul(enc(gen), tl(k), out)
        assert dec(out) == o.g2_fast_mul(k), k
    # the faithful (affine, ext-Euclid) oracle on small multipliers, incl. kzg.rs:37's [g2, [alpha]g2]
    for k in [1, 2, 9, 14, 123456789]:
        out = A32()
        L.emul_g2_scalar_mul(enc(gen), tl(k), out)
        assert dec(out) == g.mul_ref(k).affine_ints()
    # special cases of the addition (curve.rs:103-161)
    pts = [o.g2_fast_mul(rnd.randrange(1, R)) for _ in range(4)] + [gen, o.g2_fast_mul(2), None]
    neg = lambda p: None if p is None else (p[0], ((-p[1][0]) % P, (-p[1][1]) % P))
    for a in pts:
        for b in pts + [neg(a)]:
            out = A32()
            L.emul_g2_add(enc(a), enc(b), out)
            assert dec(out) == o._g2_fast_add(a, b)
    # Jacobian + Jacobian: generic, P + P, P + (-P), infinity on either side
    h = o.g2_fast_mul(777)
    for k1, pa, k2, pb in [(5, gen, 9, h), (3, gen, 3, gen), (4, gen, R - 4, gen), (0, gen, 6, h), (6, h, 0, gen), (0, gen, 0, h),
                           (rnd.randrange(R), gen, rnd.randrange(R), h)]:
        out = A32()
        L.emul_g2_lincomb(enc(pa), tl(k1), enc(pb), tl(k2), out)
        assert dec(out) == o._g2_fast_add(o.g2_fast_mul(k1, pa), o.g2_fast_mul(k2, pb)), (k1, k2)
    for _ in range(20):
        x = (rnd.randrange(P), rnd.randrange(P))
        out = (ctypes.c_uint32 * 16)()
        L.emul_fq2_inv((ctypes.c_uint32 * 16)(*[(v >> (32 * i)) & 0xFFFFFFFF for v in x for i in range(8)]), out)
        assert (fl(out), fl(out, 8)) == tuple((o.Fq2(list(x)).inverse()).c)


def test_pairing_emulated():
    """csrc/pairing.cuh (Fq12 product spread over lanes, twist-side Miller loop, final exponentiation) against the
    oracle's restatement of the reference's optimal_ate_pairing (bn128.rs:147-181): identical Fq12 VALUES."""
    L = _build("emul_pairing")
    rnd = random.Random(5)
    A96 = ctypes.c_uint32 * 96

    def enc12(c):
        return A96(*[(v >> (32 * i)) & 0xFFFFFFFF for v in c for i in range(8)])

    def dec12(a):
        return [fl(a, 8 * k) for k in range(12)]

    for _ in range(6):
        x = [rnd.randrange(P) for _ in range(12)]
        y = [rnd.randrange(P) for _ in range(12)] if rnd.random() < 0.7 else [rnd.choice([0, 1, P - 1]) for _ in range(12)]
        out = A96()
        L.emul_f12_mul(enc12(x), enc12(y), out)
        assert dec12(out) == (o.Fq12(x) * o.Fq12(y)).c

    # final exponentiation by parts == the plain power, on arbitrary (non-unitary) elements and on a Miller value
    for _ in range(3):
        x = [rnd.randrange(P) for _ in range(12)]
        slow, fast = A96(), A96()
        L.emul_final_exp(enc12(x), 0, slow)
        L.emul_final_exp(enc12(x), 1, fast)
        assert dec12(fast) == dec12(slow) == o.Fq12(x).pow(o.FINAL_EXPONENT).c

    def enc1(pt):
        vals = [0, 0] if pt is None else list(pt)
        return (ctypes.c_uint32 * 16)(*[(v >> (32 * i)) & 0xFFFFFFFF for v in vals for i in range(8)])

    def enc2(pt):
        vals = [0, 0, 0, 0] if pt is None else [pt[0][0], pt[0][1], pt[1][0], pt[1][1]]
        return (ctypes.c_uint32 * 32)(*[(v >> (32 * i)) & 0xFFFFFFFF for v in vals for i in range(8)])

    def pairing(p1, p2):
        out = A96()
        L.emul_pairing(enc1(p1), enc2(p2), 1, out)
        return dec12(out)

    def oracle_pairing(k1, k2):
        return o.optimal_ate_pairing(o.generator_g1().mul_ref(k1), o.generator_g2().mul_ref(k2)).c

    one = [1] + [0] * 11
    out = A96()
    L.emul_pairing(enc1(o.fast_mul(5)), enc2(o.g2_fast_mul(11)), 2, out)  # Miller loop + final exponentiation by parts
    assert dec12(out) == oracle_pairing(55, 1)
    assert pairing(o.fast_mul(1), o.g2_fast_mul(1)) == oracle_pairing(1, 1)
    assert pairing(o.fast_mul(37), o.g2_fast_mul(27)) == oracle_pairing(999, 1)  # bn128.rs:362-364
    assert pairing(None, o.g2_fast_mul(3)) == one and pairing(o.fast_mul(3), None) == one
    e1 = o.Fq12(pairing(o.fast_mul(1), o.g2_fast_mul(1)))
    neg = o.fast_mul(R - 1)
    assert (e1 * o.Fq12(pairing(neg, o.g2_fast_mul(1)))).c == one  # bn128.rs:345-347

"""Pins the oracle against every known-answer test the reference holds under the
KZG hot path plus the survey's anchors (appendix B)."""
import myzkp_oracle as o
from myzkp_oracle import Fq, Fr, G1Point, Polynomial, make_field


def test_g1_kats():  # bn128.rs:285-301 (test_g1)
    g1 = o.generator_g1()
    assert g1.y.pow(2) - g1.x.pow(3) == Fq.from_value(3)
    assert g1 * 2 + g1 + g1 == (g1 * 2) * 2
    assert g1 * 9 + g1 * 5 == g1 * 12 + g1 * 2
    assert (g1 * o.order()).is_point_at_infinity()


def test_2g_is_eip196_value():  # public EIP-196 vector
    x, y = (o.generator_g1() * 2).affine_ints()
    assert x == 0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3
    assert y == 0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4


def test_fq_kats():  # bn128.rs:240-251 (test_fq)
    F = Fq.from_value
    assert F(2) * F(2) == F(4)
    assert F(2) / F(7) + F(9) / F(7) == F(11) / F(7)
    assert F(2) * F(7) + F(9) * F(7) == F(11) * F(7)


def test_field_kats():  # field.rs:443-550
    F17 = make_field(17, "F17")
    F31 = make_field(31, "F31")
    assert F17(7).inverse() == F17(5)  # :491-497
    assert F31(-23) == F31(8)  # :544-550
    assert F17(-10).value == -10  # :499-504 negative stored as is (truncated %)
    assert F17(-10).sanitize().value == 7
    assert F17(0).inverse() == F17(0)  # ext-Euclid on 0 gives 0, no panic


def test_fr_limbs_of_small_negatives():  # cuda/test_fr.cu:16-17
    m2 = Fr(-2).sanitize().value
    assert [(m2 >> (64 * i)) & (2**64 - 1) for i in range(4)] == [
        0x43E1F593EFFFFFFF, 0x2833E84879B97091, 0xB85045B68181585D, 0x30644E72E131A029]
    assert (Fr(-2) * Fr(-12)) == Fr(24)
    assert Fr(5) * Fr(7) == Fr(35)


def test_polynomial_kats():
    F = Fr.from_value
    # polynomial.rs:727-755: (3 + 3x + x^2) / (1 + x) = 2 + x rem 1
    q, r = Polynomial([F(3), F(3), F(1)]).div_rem_ref(Polynomial([F(1), F(1)]))
    assert q.canonical() == [2, 1] and r.canonical() == [1]
    # polynomial.rs:757-768: (2 + 3x)(2) = 8
    assert Polynomial([F(2), F(3)]).eval(F(2)) == F(8)
    # polynomial.rs:805-821: (x-2)(x-3) = 6 - 5x + x^2
    assert Polynomial.from_monomials([F(2), F(3)]).canonical() == [6, o.R_MOD - 5, 1]


def test_kzg_anchor_fixed_alpha():  # kzg.rs:152-175 polynomial, SURVEY appendix B values
    F = Fr.from_value
    f = Polynomial.from_monomials([F(-1), F(-2), F(-3)])
    assert f.canonical() == [6, 11, 6, 1]
    pk = o.setup_kzg(o.generator_g1(), 3, 123456789)
    assert len(pk.powers_1) == 4  # max_d + 1 (kzg.rs:32)
    c = o.commit_kzg(f, pk)
    assert c.affine_ints() == (
        8096424998935924997123460782489249937183001369792870392058374165119638207724,
        14698683656276342473960081670169131092130433153277961881223581660609015832377)
    pr = o.open_kzg(f, F(5), pk)
    assert pr.y == F(336)
    assert pr.w.affine_ints() == (
        15737316170989375530370354340609809222984715696988518295913516551941326522818,
        13254863773102499080085687663253363332659578358445016579269372167128026496803)
    # algebraic identities used as the large-N oracle
    assert o.expected_commit(f.canonical(), 123456789) == c.affine_ints()
    assert o.expected_open(f.canonical(), 5, 123456789) == (336, pr.w.affine_ints())
    assert o.synthetic_division(f.canonical(), 5) == (336, [66, 11, 1])


def test_open_edge_cases():  # polynomial.rs:372-374
    pk = o.setup_kzg(o.generator_g1(), 2, 77)
    pr = o.open_kzg(Polynomial([Fr(9)]), Fr(5), pk)
    assert pr.y == Fr(9) and pr.w.is_point_at_infinity()
    assert o.commit_kzg(Polynomial([]), pk).is_point_at_infinity()
    assert o.commit_kzg(Polynomial([Fr(0), Fr(0)]), pk).is_point_at_infinity()


def test_gemini_fold_kats():  # gemini.rs:288-307, book gemini.md:311-320
    F = Fr.from_value
    coef = [F(i + 1) for i in range(8)]
    fs = o.split_and_fold(coef, [F(2), F(3), F(4)])
    assert [p.canonical() for p in fs] == [[1, 2, 3, 4, 5, 6, 7, 8], [5, 11, 17, 23], [38, 86], [382]]
    fs = o.split_and_fold(coef, [F(1), F(2), F(3)])
    assert [p.canonical() for p in fs][1:] == [[3, 7, 11, 15], [17, 41], [140]]
    assert o.fold_ints(list(range(1, 9)), [2, 3, 4])[1:] == [[5, 11, 17, 23], [38, 86], [382]]
    import pytest
    with pytest.raises(o.SplitFoldError):
        o.split_and_fold(coef[:7], [F(1), F(2)])
    with pytest.raises(o.SplitFoldError):
        o.split_and_fold(coef, [F(1), F(2)])


def test_fast_paths_agree_with_faithful():
    import random
    rnd = random.Random(5)
    g = o.generator_g1()
    for k in [1, 2, 3, 5, 12345, o.R_MOD - 1, rnd.randrange(o.R_MOD)]:
        assert (g * k).affine_ints() == o.fast_mul(k)
    coefs = [rnd.randrange(o.R_MOD) for _ in range(9)]
    u = rnd.randrange(o.R_MOD)
    f = Polynomial([Fr(c) for c in coefs])
    y, q = o.synthetic_division(coefs, u)
    assert f.eval(Fr(u)).sanitize().value == y
    qq = (f - Polynomial([Fr(y)])) / Polynomial.from_monomials([Fr(u)])
    assert qq.canonical() == q
    pk = o.setup_kzg(g, 8, 31337)
    assert o.commit_kzg(f, pk).affine_ints() == o.expected_commit(coefs, 31337)
    pr = o.open_kzg(f, Fr(u), pk)
    assert (pr.y.sanitize().value, pr.w.affine_ints()) == o.expected_open(coefs, u, 31337)


def test_g2_kats_bn128_rs_304_320():
    """bn128.rs:304-320 (test_g2) and bn128.rs:254-283 (test_fq2) on the oracle's G2 restatement, the
    public value of 2*G2 (EIP-197 / py_ecc bn128 test vector), and the fast cross-check path."""
    g = o.generator_g2()
    assert g.y.mul_ref(g.y) - g.x.mul_ref(g.x).mul_ref(g.x) == o.get_b2()
    assert g * 2 + g + g == (g * 2) * 2
    assert g * 9 + g * 5 == g * 12 + g * 2
    assert (g * o.order()).is_point_at_infinity()
    assert (g * 2).affine_ints() == (
        (18029695676650738226693292988307914797657423701064905010927197838374790804409,
         14583779054894525174450323658765874724019480979794335525732096752006891875705),
        (2140229616977736810657479771656733941598412651537078903776637920509952744750,
         11474861747383700316476719153975578001603231366361248090558603872215261634898))
    # test_fq2
    x, f = o.Fq2([1, 0]), o.Fq2([1, 2])
    assert x.add_ref(f) == o.Fq2([2, 2])
    assert o.Fq2([2, 1]).div_ref(o.Fq2([2, 1])) == o.Fq2.one()
    assert o.Fq2.one().div_ref(f) + x.div_ref(f) == (o.Fq2.one() + x) / f
    assert o.Fq2.one() * f + x * f == (o.Fq2.one() + x) * f
    # u^2 = -1; the fast path agrees with the faithful one; the wire format round-trips
    assert o.Fq2([0, 1]) * o.Fq2([0, 1]) == o.Fq2([-1])
    for k in (1, 2, 3, 77, 123456789):
        assert (g * k).affine_ints() == o.g2_fast_mul(k)
        assert o.g2_from_bytes(o.g2_to_bytes(o.g2_fast_mul(k))) == o.g2_fast_mul(k)
    assert o.g2_fast_mul(o.R_MOD) is None and o.g2_from_bytes(bytes(128)) is None
    pw = o.setup_kzg_g2(g, 5, 3)
    assert [p.affine_ints() for p in pw] == [o.g2_fast_mul(1), o.g2_fast_mul(5), o.g2_fast_mul(25)]


def test_pairing_kats_bn128_rs_341_365():
    """bn128.rs:341-365 (test_pairing) and bn128.rs:322-339 (test_g12) on the oracle's restatement of
    optimal_ate_pairing / miller / get_lambda, and kzg.rs:152-175 (test_kzg) through its verify_kzg."""
    g1, g2 = o.generator_g1(), o.generator_g2()
    one = o.Fq12.one()
    p1 = o.optimal_ate_pairing(g1, g2)
    pn1 = o.optimal_ate_pairing(-g1, g2)
    assert p1 * pn1 == one
    np1 = o.optimal_ate_pairing(g1, -g2)
    assert p1 * np1 == one and pn1 == np1
    p2 = o.optimal_ate_pairing(g1.mul_ref(2), g2)
    assert p1 * p1 == p2 and p1 != p2 and p1 != np1 and p2 != np1
    assert p1 * p1 == o.optimal_ate_pairing(g1, g2.mul_ref(2))
    assert o.optimal_ate_pairing(g1.mul_ref(37), g2.mul_ref(27)) == o.optimal_ate_pairing(g1.mul_ref(999), g2)
    # test_g12: the twisted generator lies on y^2 = x^3 + 3 over Fq12
    g12 = o.twist_g2_to_g12(g2)
    g12_9 = g12.mul_ref(9)
    assert g12_9.y.pow(2) - g12_9.x.pow(3) == o.Fq12([3])
    assert g12.mul_ref(2) + g12 + g12 == g12.mul_ref(2).mul_ref(2)
    # test_kzg with a fixed alpha
    alpha = 123456789
    pk = o.setup_kzg(g1, 3, alpha)
    p2s = o.setup_kzg_g2(g2, alpha, 2)
    f = o.Polynomial([o.Fr(v) for v in [6, 11, 6, 1]])
    c, proof = o.commit_kzg(f, pk), o.open_kzg(f, o.Fr(5), pk)
    assert o.verify_kzg(o.Fr(5), c, proof, pk.powers_1, p2s)

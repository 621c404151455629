// BN128 G2 for the verifier half of the KZG public key (`powers_2`): the reference's
// EllipticCurvePoint<Fq2, BN128Curve> (curve/bn128.rs:33-49) with Fq2 = Fq[u] / (u^2 + 1)
// (bn128.rs:36-41), multiplied by powers of alpha in setup_kzg / setup_kzg_with_full_g2
// (kzg.rs:37, 47-52).  The reference uses its affine law with one polynomial ext-Euclid inversion
// per operation (curve.rs:56-191, efield.rs:126-151); here: Jacobian coordinates over Fq2, a = 0,
// every special case of curve.rs:103-161 branched on, one inversion at the end.  Not a hot path.
#pragma once
#include "field.cuh"

namespace mz {

struct Fq2 {
  Fq c0, c1;  // c0 + c1 u
};

MZ_HD Fq2 f2_zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
MZ_HD Fq2 f2_one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
MZ_HD bool f2_is_zero(const Fq2& a) { return a.c0.is_zero() && a.c1.is_zero(); }
MZ_HD bool f2_eq(const Fq2& a, const Fq2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
MZ_HD Fq2 f2_add(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fe_add(a.c0, b.c0); r.c1 = fe_add(a.c1, b.c1); return r; }
MZ_HD Fq2 f2_sub(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fe_sub(a.c0, b.c0); r.c1 = fe_sub(a.c1, b.c1); return r; }
MZ_HD Fq2 f2_dbl(const Fq2& a) { return f2_add(a, a); }
MZ_HD Fq2 f2_neg(const Fq2& a) { Fq2 r; r.c0 = fe_neg(a.c0); r.c1 = fe_neg(a.c1); return r; }
// (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u: two sums of two products, one reduction each
MZ_HD Fq2 f2_mul(const Fq2& a, const Fq2& b) {
  Fq2 r;
  r.c0 = fe_mul_sub_mul(a.c0, b.c0, a.c1, b.c1);
  r.c1 = fe_mul2(a.c0, b.c1, a.c1, b.c0);
  return r;
}
// (a0 + a1)(a0 - a1) + 2 a0 a1 u
MZ_HD Fq2 f2_sqr(const Fq2& a) {
  Fq2 r;
  r.c0 = fe_mul(fe_add(a.c0, a.c1), fe_sub(a.c0, a.c1));
  r.c1 = fe_dbl(fe_mul(a.c0, a.c1));
  return r;
}
// conj(a) / (a0^2 + a1^2); inverse(0) = 0 like the base field
MZ_HD Fq2 f2_inv(const Fq2& a) {
  Fq n = fe_inv_bingcd(fe_mul2(a.c0, a.c0, a.c1, a.c1));
  Fq2 r;
  r.c0 = fe_mul(a.c0, n);
  r.c1 = fe_neg(fe_mul(a.c1, n));
  return r;
}

struct AffineG2 {  // infinity = all-zero ((0, 0) is not on y^2 = x^3 + 3/(9+u))
  Fq2 x, y;
};
struct JacG2 {  // x = X/Z^2, y = Y/Z^3; infinity <=> Z == 0
  Fq2 x, y, z;
};

MZ_HD bool g2_affine_is_inf(const AffineG2& p) { return f2_is_zero(p.x) && f2_is_zero(p.y); }
MZ_HD JacG2 g2_jac_inf() { JacG2 r; r.x = f2_one(); r.y = f2_one(); r.z = f2_zero(); return r; }

// a = 0 doubling (dbl-2009-l): A=X^2 B=Y^2 C=B^2 D=2((X+B)^2-A-C) E=3A F=E^2 X3=F-2D Y3=E(D-X3)-8C Z3=2YZ.
// A point with y = 0 (order 2; none in the r-torsion subgroup, possible for an arbitrary base) gives Z3 = 0.
MZ_HD void g2_jac_dbl(JacG2& p) {
  if (f2_is_zero(p.z)) return;
  Fq2 a = f2_sqr(p.x);
  Fq2 b = f2_sqr(p.y);
  Fq2 c = f2_sqr(b);
  Fq2 d = f2_dbl(f2_sub(f2_sub(f2_sqr(f2_add(p.x, b)), a), c));
  Fq2 e = f2_add(f2_dbl(a), a);
  Fq2 f = f2_sqr(e);
  Fq2 z3 = f2_dbl(f2_mul(p.y, p.z));
  Fq2 x3 = f2_sub(f, f2_dbl(d));
  Fq2 c8 = f2_dbl(f2_dbl(f2_dbl(c)));
  p.y = f2_sub(f2_mul(e, f2_sub(d, x3)), c8);
  p.x = x3;
  p.z = z3;
}

// acc += q (mixed, madd-2007-bl) with the reference's cases: inf + Q, P + inf, P + P, P + (-P)
MZ_HD void g2_jac_madd(JacG2& acc, const AffineG2& q) {
  if (g2_affine_is_inf(q)) return;
  if (f2_is_zero(acc.z)) { acc.x = q.x; acc.y = q.y; acc.z = f2_one(); return; }
  Fq2 z1z1 = f2_sqr(acc.z);
  Fq2 u2 = f2_mul(q.x, z1z1);
  Fq2 s2 = f2_mul(f2_mul(q.y, acc.z), z1z1);
  Fq2 h = f2_sub(u2, acc.x);
  Fq2 rr = f2_sub(s2, acc.y);
  if (f2_is_zero(h)) {
    if (f2_is_zero(rr)) g2_jac_dbl(acc);
    else acc = g2_jac_inf();
    return;
  }
  Fq2 hh = f2_sqr(h);
  Fq2 i = f2_dbl(f2_dbl(hh));
  Fq2 j = f2_mul(h, i);
  Fq2 r2 = f2_dbl(rr);
  Fq2 v = f2_mul(acc.x, i);
  Fq2 x3 = f2_sub(f2_sub(f2_sqr(r2), j), f2_dbl(v));
  Fq2 y3 = f2_sub(f2_mul(r2, f2_sub(v, x3)), f2_dbl(f2_mul(acc.y, j)));
  Fq2 z3 = f2_sub(f2_sub(f2_sqr(f2_add(acc.z, h)), z1z1), hh);
  acc.x = x3; acc.y = y3; acc.z = z3;
}

// acc += q (Jacobian + Jacobian, add-2007-bl) with the same case analysis
MZ_HD void g2_jac_add(JacG2& acc, const JacG2& q) {
  if (f2_is_zero(q.z)) return;
  if (f2_is_zero(acc.z)) { acc = q; return; }
  Fq2 z1z1 = f2_sqr(acc.z);
  Fq2 z2z2 = f2_sqr(q.z);
  Fq2 u1 = f2_mul(acc.x, z2z2);
  Fq2 u2 = f2_mul(q.x, z1z1);
  Fq2 s1 = f2_mul(f2_mul(acc.y, q.z), z2z2);
  Fq2 s2 = f2_mul(f2_mul(q.y, acc.z), z1z1);
  Fq2 h = f2_sub(u2, u1);
  Fq2 rr = f2_sub(s2, s1);
  if (f2_is_zero(h)) {
    if (f2_is_zero(rr)) g2_jac_dbl(acc);
    else acc = g2_jac_inf();
    return;
  }
  Fq2 i = f2_sqr(f2_dbl(h));
  Fq2 j = f2_mul(h, i);
  Fq2 r2 = f2_dbl(rr);
  Fq2 v = f2_mul(u1, i);
  Fq2 x3 = f2_sub(f2_sub(f2_sqr(r2), j), f2_dbl(v));
  Fq2 y3 = f2_sub(f2_mul(r2, f2_sub(v, x3)), f2_dbl(f2_mul(s1, j)));
  Fq2 z3 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(acc.z, q.z)), z1z1), z2z2), h);
  acc.x = x3; acc.y = y3; acc.z = z3;
}

MZ_HD AffineG2 g2_jac_to_affine(const JacG2& p) {
  AffineG2 r;
  if (f2_is_zero(p.z)) { r.x = f2_zero(); r.y = f2_zero(); return r; }
  Fq2 zi = f2_inv(p.z);
  Fq2 zi2 = f2_sqr(zi);
  r.x = f2_mul(p.x, zi2);
  r.y = f2_mul(p.y, f2_mul(zi2, zi));
  return r;
}

// [k] base, k = 8 raw (non-Montgomery) little-endian limbs, MSB first; k = 0 -> infinity (curve.rs:169-171)
MZ_HD JacG2 g2_scalar_mul_jac(const AffineG2& base, const uint32_t* k) {
  JacG2 acc = g2_jac_inf();
  bool started = false;
  for (int limb = 7; limb >= 0; limb--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) g2_jac_dbl(acc);
      if ((k[limb] >> bit) & 1u) {
        g2_jac_madd(acc, base);
        started = true;
      }
    }
  }
  return acc;
}
MZ_HD AffineG2 g2_scalar_mul(const AffineG2& base, const uint32_t* k) { return g2_jac_to_affine(g2_scalar_mul_jac(base, k)); }

}  // namespace mz

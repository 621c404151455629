// BN128 G1 group law for the MSM: affine inputs (Montgomery, 64 B) and XYZZ
// accumulators (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0).
//
// Replaces the reference's affine law with one ext-Euclid inversion per
// operation (myzkp/src/modules/algebra/curve/curve.rs:56-161).  Every special
// case the reference branches on (curve.rs:131-145: inf + Q, P + inf,
// P + P -> double, P + (-P) -> inf) is branched on here too, so results are
// the same group elements and - after the final to-affine - the same
// canonical coordinates.
//
// Curve: y^2 = x^3 + 3 (a = 0, bn128.rs:23).  G1 has prime order, so no finite
// point has y == 0 and doubling never hits 2-torsion.
#pragma once
#include "field.cuh"

// The mixed addition is the hot loop of the MSM.  Its multiplies can be switched to the Karatsuba
// forms (13 % fewer wide multiplies, ~80 more integer adds each); measured on B200 that is SLOWER
// (accumulate 29.2 -> 36.5 ms at 2^24: ptxas moves part of the extra carry work onto the multiply
// pipe as IMAD.X / IMAD.MOV and the rest competes for issue slots), so it stays off.
#ifndef MZ_MADD_KARATSUBA
#define MZ_MADD_KARATSUBA 0
#endif
#if MZ_MADD_KARATSUBA
#define MZ_MADD_MUL fe_mul_k
#define MZ_MADD_MUL2 fe_mul2_k
#else
#define MZ_MADD_MUL fe_mul
#define MZ_MADD_MUL2 fe_mul2
#endif

namespace mz {

struct Affine {  // infinity = all-zero ((0,0) is not on the curve)
  Fq x, y;
};
struct XYZZ {
  Fq x, y, zz, zzz;
};
struct Jac {  // x = X/Z^2, y = Y/Z^3, infinity <=> Z == 0 (SRS table build only)
  Fq x, y, z;
};

MZ_HD bool affine_is_inf(const Affine& p) { return p.x.is_zero() && p.y.is_zero(); }
MZ_HD bool xyzz_is_inf(const XYZZ& p) { return p.zz.is_zero(); }

MZ_HD XYZZ xyzz_inf() {
  XYZZ r;
  r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
  return r;
}
MZ_HD XYZZ xyzz_from_affine(const Affine& p) {
  XYZZ r;
  if (affine_is_inf(p)) return xyzz_inf();
  r.x = p.x; r.y = p.y; r.zz = Fq::one(); r.zzz = Fq::one();
  return r;
}
MZ_HD Affine affine_neg(const Affine& p) {
  Affine r;
  r.x = p.x;
  r.y = fe_neg(p.y);  // -0 = 0 keeps infinity
  return r;
}

// 2 * (affine), a = 0: 4M + 3S  (U=2y, V=U^2, W=U*V, S=x*V, M=3x^2)
MZ_HD XYZZ xyzz_mdbl(const Affine& p) {
  XYZZ r;
  Fq u = fe_dbl(p.y);
  Fq v = fe_sqr(u);
  Fq w = fe_mul(u, v);
  Fq s = fe_mul(p.x, v);
  Fq xx = fe_sqr(p.x);
  Fq m = fe_add(fe_dbl(xx), xx);
  r.x = fe_sub(fe_sqr(m), fe_dbl(s));
  r.y = fe_mul_sub_mul(m, fe_sub(s, r.x), w, p.y);  // one reduction for both products
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * (XYZZ), a = 0: 6M + 3S
MZ_HD void xyzz_dbl(XYZZ& p) {
  if (xyzz_is_inf(p)) return;
  Fq u = fe_dbl(p.y);
  Fq v = fe_sqr(u);
  Fq w = fe_mul(u, v);
  Fq s = fe_mul(p.x, v);
  Fq xx = fe_sqr(p.x);
  Fq m = fe_add(fe_dbl(xx), xx);
  Fq x3 = fe_sub(fe_sqr(m), fe_dbl(s));
  Fq y3 = fe_mul_sub_mul(m, fe_sub(s, x3), w, p.y);
  p.x = x3;
  p.y = y3;
  p.zz = fe_mul(v, p.zz);
  p.zzz = fe_mul(w, p.zzz);
}

// Field-multiply policy of the hot formulas: InlineOps expands every multiply in place (the XYZZ-only
// accumulate: its ~27 KB loop fits the 32 KB instruction cache); a kernel whose loop is larger passes a
// policy whose multiplies are out-of-line calls, so the loop holds ONE copy of each multiply.
struct InlineOps {
  static MZ_HD Fq mul(const Fq& a, const Fq& b) { return MZ_MADD_MUL(a, b); }
  static MZ_HD Fq sqr(const Fq& a) { return fe_sqr(a); }
  static MZ_HD Fq mul2(const Fq& a, const Fq& b, const Fq& c, const Fq& d) { return MZ_MADD_MUL2(a, b, c, d); }
};

// 2 * (affine) with a multiply policy (see xyzz_mdbl)
template <class OPS>
MZ_HD XYZZ xyzz_mdbl_t(const Affine& p) {
  XYZZ r;
  Fq u = fe_dbl(p.y);
  Fq v = OPS::sqr(u);
  Fq w = OPS::mul(u, v);
  Fq s = OPS::mul(p.x, v);
  Fq xx = OPS::sqr(p.x);
  Fq m = fe_add(fe_dbl(xx), xx);
  r.x = fe_sub(OPS::sqr(m), fe_dbl(s));
  r.y = OPS::mul2(m, fe_sub(s, r.x), fe_neg(w), p.y);
  r.zz = v;
  r.zzz = w;
  return r;
}

// acc += q (mixed add, 8M + 2S) with the reference's case analysis
template <class OPS>
MZ_HD void xyzz_madd_t(XYZZ& acc, const Affine& q) {
  if (affine_is_inf(q)) return;                      // P + inf       (curve.rs:135-137)
  if (xyzz_is_inf(acc)) {                            // inf + Q       (curve.rs:131-134)
    acc.x = q.x; acc.y = q.y; acc.zz = Fq::one(); acc.zzz = Fq::one();
    return;
  }
  Fq p = fe_sub(OPS::mul(q.x, acc.zz), acc.x);       // U2 - X1
  Fq r = fe_sub(OPS::mul(q.y, acc.zzz), acc.y);      // S2 - Y1
  if (p.is_zero()) {
    if (r.is_zero()) acc = xyzz_mdbl_t<OPS>(q);      // P + P         (curve.rs:139-141)
    else acc = xyzz_inf();                           // P + (-P)      (curve.rs:142-145)
    return;
  }
  Fq pp = OPS::sqr(p);
  Fq ppp = OPS::mul(p, pp);
  Fq qq = OPS::mul(acc.x, pp);
  Fq x3 = fe_sub(fe_sub(OPS::sqr(r), ppp), fe_dbl(qq));
  Fq y3 = OPS::mul2(r, fe_sub(qq, x3), fe_neg(acc.y), ppp);  // r (qq - x3) - y1 ppp, one reduction for both products
  acc.x = x3;
  acc.y = y3;
  acc.zz = OPS::mul(acc.zz, pp);
  acc.zzz = OPS::mul(acc.zzz, ppp);
}
MZ_HD void xyzz_madd(XYZZ& acc, const Affine& q) { xyzz_madd_t<InlineOps>(acc, q); }

// acc += q (XYZZ + XYZZ, 12M + 2S) with the same case analysis
MZ_HD void xyzz_add(XYZZ& acc, const XYZZ& q) {
  if (xyzz_is_inf(q)) return;
  if (xyzz_is_inf(acc)) { acc = q; return; }
  Fq u1 = fe_mul(acc.x, q.zz);
  Fq u2 = fe_mul(q.x, acc.zz);
  Fq s1 = fe_mul(acc.y, q.zzz);
  Fq s2 = fe_mul(q.y, acc.zzz);
  Fq p = fe_sub(u2, u1);
  Fq r = fe_sub(s2, s1);
  if (p.is_zero()) {
    if (r.is_zero()) xyzz_dbl(acc);
    else acc = xyzz_inf();
    return;
  }
  Fq pp = fe_sqr(p);
  Fq ppp = fe_mul(p, pp);
  Fq qq = fe_mul(u1, pp);
  Fq x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(qq));
  Fq y3 = fe_mul_sub_mul(r, fe_sub(qq, x3), s1, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = fe_mul(fe_mul(acc.zz, q.zz), pp);
  acc.zzz = fe_mul(fe_mul(acc.zzz, q.zzz), ppp);
}

// XYZZ -> affine given izzz = 1/ZZZ: 1/ZZ = (ZZ * izzz)^2 because ZZ^3 = ZZZ^2
MZ_HD Affine xyzz_to_affine_with_inv(const XYZZ& p, const Fq& izzz) {
  Affine r;
  Fq t = fe_mul(p.zz, izzz);
  r.x = fe_mul(p.x, fe_sqr(t));
  r.y = fe_mul(p.y, izzz);
  return r;
}
MZ_HD Affine xyzz_to_affine(const XYZZ& p) {
  if (xyzz_is_inf(p)) { Affine z; z.x = Fq::zero(); z.y = Fq::zero(); return z; }
  return xyzz_to_affine_with_inv(p, fe_inv_bingcd(p.zzz));
}

// Jacobian doubling, a = 0 (2M + 5S): A=X^2 B=Y^2 C=B^2 D=2((X+B)^2-A-C)
// E=3A F=E^2 X3=F-2D Y3=E(D-X3)-8C Z3=2YZ
MZ_HD void jac_dbl(Jac& p) {
  Fq a = fe_sqr(p.x);
  Fq b = fe_sqr(p.y);
  Fq c = fe_sqr(b);
  Fq t = fe_add(p.x, b);
  Fq d = fe_dbl(fe_sub(fe_sub(fe_sqr(t), a), c));
  Fq e = fe_add(fe_dbl(a), a);
  Fq f = fe_sqr(e);
  Fq z3 = fe_dbl(fe_mul(p.y, p.z));
  Fq x3 = fe_sub(f, fe_dbl(d));
  Fq c8 = fe_dbl(fe_dbl(fe_dbl(c)));
  p.y = fe_sub(fe_mul(e, fe_sub(d, x3)), c8);
  p.x = x3;
  p.z = z3;
}

}  // namespace mz

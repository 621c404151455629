// Modular inversion by the Bernstein-Yang "safegcd" divsteps on signed 30-bit limbs - a branch-free
// inverse whose cost does not depend on the value, so 32 lanes of a warp can each invert their own
// element in lockstep.  Used by the batched-affine bucket accumulation (one inversion per lane per
// batch of affine additions); the reference inverts once per affine addition with the extended
// Euclid algorithm (field.rs:210-237, utils.rs:52-81) - same value, inverse(0) = 0.
//
// Layout of the algorithm (public; the same structure as the 32-bit safegcd in common EC libraries):
//   * f = modulus, g = x, d = 0, e = 1, zeta = -1
//   * 20 rounds: 30 divsteps on the low limbs only (32-bit integer-add pipe work) give a 2x2
//     transition matrix t with entries in [-2^30, 2^30]; then (f, g) <- t (f, g) / 2^30 exactly and
//     (d, e) <- t (d, e) / 2^30 mod modulus (signed 32x32+64 multiply-adds)
//   * after 600 divsteps (590 suffice for 256-bit inputs) g = 0, f = +-1 and d = +-x^-1
// About 12 k integer-add-pipe instructions and 1.8 k wide multiplies per inverse: ~1/20 of the multiply
// pipe time of the Fermat ladder (fe_inv), and no divergence, unlike the binary GCD (fe_inv_bingcd).
#pragma once
#include "field.cuh"

namespace mz {
namespace safegcd {

constexpr int32_t kM30 = (int32_t)(0xffffffffu >> 2);

struct S30 {
  int32_t v[9];
};
struct Trans {
  int32_t u, v, q, r;
};

// limb i (30 bits) of a 256-bit value given as 8 x 32-bit limbs through an accessor
template <class A>
MZ_HD constexpr uint32_t limb30(const A& a, int i) {
  // bits [30 i, 30 i + 30)
  const int bit = 30 * i;
  const int w = bit >> 5, sh = bit & 31;
  uint32_t lo = a(w) >> sh;
  if (sh > 2 && w + 1 < 8) lo |= a(w + 1) << (32 - sh);
  return lo & (uint32_t)kM30;
}
template <class PR>
struct ModLimbs {
  MZ_HD constexpr uint32_t operator()(int i) const { return PR::mod(i); }
};
struct ArrLimbs {
  const uint32_t* p;
  MZ_HD uint32_t operator()(int i) const { return p[i]; }
};

MZ_HD void to_s30(const uint32_t* a, S30& r) {
  ArrLimbs acc{a};
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] = (int32_t)limb30(acc, i);
}
// normalized value in [0, 2^256) -> 8 x 32-bit limbs
MZ_HD void from_s30(const S30& s, uint32_t* a) {
  const uint32_t* v = reinterpret_cast<const uint32_t*>(s.v);
  a[0] = v[0] | (v[1] << 30);
  a[1] = (v[1] >> 2) | (v[2] << 28);
  a[2] = (v[2] >> 4) | (v[3] << 26);
  a[3] = (v[3] >> 6) | (v[4] << 24);
  a[4] = (v[4] >> 8) | (v[5] << 22);
  a[5] = (v[5] >> 10) | (v[6] << 20);
  a[6] = (v[6] >> 12) | (v[7] << 18);
  a[7] = (v[7] >> 14) | (v[8] << 16);
}

// 30 divsteps on the low 30 bits of f and g; returns the new zeta and the transition matrix
MZ_HD int32_t divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, Trans& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
#pragma unroll
  for (int i = 0; i < 30; i++) {
    uint32_t mask1 = (uint32_t)(zeta >> 31);  // zeta < 0
    const uint32_t mask2 = 0u - (g & 1u);     // g odd
    const uint32_t x = (f ^ mask1) - mask1;   // conditionally negated f, u, v
    const uint32_t y = (u ^ mask1) - mask1;
    const uint32_t z = (v ^ mask1) - mask1;
    g += x & mask2;
    q += y & mask2;
    r += z & mask2;
    mask1 &= mask2;                            // zeta < 0 and g odd: swap roles
    zeta = (int32_t)(((uint32_t)zeta ^ mask1) - 1u);
    f += g & mask1;
    u += q & mask1;
    v += r & mask1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
  return zeta;
}

// (d, e) <- t (d, e) / 2^30 mod modulus; d, e stay in (-2 modulus, modulus)
template <class PR>
MZ_HD void update_de_30(S30& d, S30& e, const Trans& t) {
  constexpr ModLimbs<PR> ml{};
  constexpr uint32_t inv30 = (0u - PR::INV) & (uint32_t)kM30;  // modulus^-1 mod 2^30 (PR::INV = -modulus^-1 mod 2^32)
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
  int32_t md = (u & sd) + (v & se);
  int32_t me = (q & sd) + (r & se);
  int32_t di = d.v[0], ei = e.v[0];
  int64_t cd = (int64_t)u * di + (int64_t)v * ei;
  int64_t ce = (int64_t)q * di + (int64_t)r * ei;
  md -= (int32_t)((inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)kM30);
  me -= (int32_t)((inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)kM30);
  cd += (int64_t)(int32_t)limb30(ml, 0) * md;
  ce += (int64_t)(int32_t)limb30(ml, 0) * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    di = d.v[i];
    ei = e.v[i];
    cd += (int64_t)u * di + (int64_t)v * ei;
    ce += (int64_t)q * di + (int64_t)r * ei;
    cd += (int64_t)(int32_t)limb30(ml, i) * md;
    ce += (int64_t)(int32_t)limb30(ml, i) * me;
    d.v[i - 1] = (int32_t)cd & kM30; cd >>= 30;
    e.v[i - 1] = (int32_t)ce & kM30; ce >>= 30;
  }
  d.v[8] = (int32_t)cd;
  e.v[8] = (int32_t)ce;
}

// (f, g) <- t (f, g) / 2^30 (exact)
MZ_HD void update_fg_30(S30& f, S30& g, const Trans& t) {
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  int32_t fi = f.v[0], gi = g.v[0];
  int64_t cf = (int64_t)u * fi + (int64_t)v * gi;
  int64_t cg = (int64_t)q * fi + (int64_t)r * gi;
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    fi = f.v[i];
    gi = g.v[i];
    cf += (int64_t)u * fi + (int64_t)v * gi;
    cg += (int64_t)q * fi + (int64_t)r * gi;
    f.v[i - 1] = (int32_t)cf & kM30; cf >>= 30;
    g.v[i - 1] = (int32_t)cg & kM30; cg >>= 30;
  }
  f.v[8] = (int32_t)cf;
  g.v[8] = (int32_t)cg;
}

// r in (-2 modulus, modulus), negated when sign < 0, brought to [0, modulus)
template <class PR>
MZ_HD void normalize_30(S30& r, int32_t sign) {
  constexpr ModLimbs<PR> ml{};
  int32_t cond_add = r.v[8] >> 31;
  const int32_t cond_negate = sign >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    r.v[i] += (int32_t)limb30(ml, i) & cond_add;
    r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= kM30;
  }
  cond_add = r.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] += (int32_t)limb30(ml, i) & cond_add;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= kM30;
  }
}

}  // namespace safegcd

// x^-1 mod modulus on raw residues (no Montgomery factor); 0 -> 0
template <class PR>
MZ_HD void limbs_inv_safegcd(const uint32_t* x, uint32_t* out) {
  using namespace safegcd;
  constexpr ModLimbs<PR> ml{};
  S30 d, e, f, g;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    d.v[i] = 0;
    e.v[i] = 0;
    f.v[i] = (int32_t)limb30(ml, i);
  }
  e.v[0] = 1;
  to_s30(x, g);
  int32_t zeta = -1;
#pragma unroll 1
  for (int it = 0; it < 20; it++) {
    Trans t;
    zeta = divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    update_de_30<PR>(d, e, t);
    update_fg_30(f, g, t);
  }
  normalize_30<PR>(d, f.v[8]);
  from_s30(d, out);
}

// Montgomery in, Montgomery out: (a R)^-1 = a^-1 R^-1, times R^2 twice -> a^-1 R
template <class PR>
MZ_HD Fe<PR> fe_inv_safegcd(const Fe<PR>& a) {
  Fe<PR> r;
  limbs_inv_safegcd<PR>(a.v, r.v);
  return fe_mul(fe_mul(r, Fe<PR>::r2()), Fe<PR>::r2());
}

}  // namespace mz

// Optimal ate pairing on BN128 for the verifier side of KZG (verify_kzg / batch_verify_kzg /
// verify_degree_bound, kzg.rs:90-144) - the reference's optimal_ate_pairing (curve/bn128.rs:147-181:
// miller() curve.rs:313-339 with get_lambda() curve.rs:285-311 over G12 points in
// Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), then f^((p^12-1)/r)).
//
// Same function, computed the cheap way.  The reference twists Q into Fq12 (bn128.rs:98-145) and runs
// the affine law there with one polynomial ext-Euclid inversion per step.  A twisted point is
// (X w^2, Y w^3) with X, Y in Fq2 (u = w^6 - 9), so its slopes are lambda * w with lambda the slope on
// the twist curve over Fq2: the point arithmetic stays in Fq2 and the line through T and Q at P = (xp, yp)
//     (yp - Y_T w^3) - lambda w (xp - X_T w^2) = yp - (lambda xp) w + (lambda X_T - Y_T) w^3
// is get_lambda's numerator exactly.  Its denominator (the vertical through T + Q) lies in Fq6 = Fq2[w^2]
// and is removed by the final exponentiation ((p^6 - 1) divides (p^12 - 1)/r), so it is not computed:
// the pairing VALUE is identical to the reference's, coefficient by coefficient in the w basis.
// The Frobenius images Q1 = (x^p, y^p), -Q2 (bn128.rs:164-172) are conj(.) times constants of Fq2.
//
// An Fq12 product is spread over 12 lanes of a warp (lane k owns the coefficients of w^k and w^(k+12),
// 12 base-field products each); the twist arithmetic runs on lane 0.  The same bodies compile for the
// host (tests/emul) with a sequential executor.  Verifier-side work: latency matters, throughput does not.
#pragma once
#include "g1.cuh"
#include "g2.cuh"

namespace mz {

struct F12 {
  Fq c[12];  // sum c[k] w^k
};

MZ_HD void f12_set_one(F12& a) {
#pragma unroll 1
  for (int k = 0; k < 12; k++) a.c[k] = Fq::zero();
  a.c[0] = Fq::one();
}

// n * a for a small constant n >= 1 (double-and-add on the add pipe, from the top set bit)
MZ_HD Fq fe_mul_small(const Fq& a, uint32_t n) {
  int top = 31;
  while (top > 0 && !((n >> top) & 1u)) top--;
  Fq r = a;
  for (int bit = top - 1; bit >= 0; bit--) {
    r = fe_dbl(r);
    if ((n >> bit) & 1u) r = fe_add(r, a);
  }
  return r;
}

// lane k of a product: lo = coefficient of w^k, hi = coefficient of w^(k+12) (k <= 10) of a * b.
// The 12 base-field products of a lane are taken two at a time with one reduction (fe_mul2).
MZ_HD void f12_mul_lane(int k, const Fq* a, const Fq* b, Fq& lo, Fq& hi) {
  lo = Fq::zero();
  hi = Fq::zero();
  int i = 0;
#pragma unroll 1
  for (; i + 1 <= k; i += 2) lo = fe_add(lo, fe_mul2(a[i], b[k - i], a[i + 1], b[k - i - 1]));
  if (i <= k) {  // odd term count: the last low product stands alone
    lo = fe_add(lo, fe_mul(a[i], b[k - i]));
    i++;
  }
#pragma unroll 1
  for (; i + 1 < 12; i += 2) hi = fe_add(hi, fe_mul2(a[i], b[k + 12 - i], a[i + 1], b[k + 11 - i]));
  if (i < 12) hi = fe_add(hi, fe_mul(a[i], b[k + 12 - i]));
}
// w^12 = 18 w^6 - 82 and w^18 = 242 w^6 - 1476:
//   c_k = lo_k - 82 hi_k - 1476 hi_(k+6)        k = 0..5
//   c_k = lo_k + 18 hi_(k-6) + 242 hi_k          k = 6..11        (hi_11 = 0)
MZ_HD Fq f12_reduce_lane(int k, const Fq* lo, const Fq* hi) {
  if (k < 6) {
    Fq r = fe_sub(lo[k], fe_mul_small(hi[k], 82));
    if (k + 6 <= 10) r = fe_sub(r, fe_mul_small(hi[k + 6], 1476));
    return r;
  }
  Fq r = fe_add(lo[k], fe_mul_small(hi[k - 6], 18));
  if (k <= 10) r = fe_add(r, fe_mul_small(hi[k], 242));
  return r;
}

// Executors: how an Fq12 product is carried out and who runs the scalar (lane-0) sections.
struct SeqExec {  // host emulation / single thread
  Fq lo[12], hi[12];
  MZ_HD bool leader() const { return true; }
  MZ_HD void sync() const {}
  MZ_HD void mul(F12& out, const F12& a, const F12& b) {
    for (int k = 0; k < 12; k++) f12_mul_lane(k, a.c, b.c, lo[k], hi[k]);
    F12 r;
    for (int k = 0; k < 12; k++) r.c[k] = f12_reduce_lane(k, lo, hi);
    out = r;
  }
};
#if defined(__CUDACC__)
struct WarpExec {  // one warp; all operands in shared memory
  Fq* lo;
  Fq* hi;
  __device__ bool leader() const { return (threadIdx.x & 31) == 0; }
  __device__ void sync() const { __syncwarp(); }
  __device__ void mul(F12& out, const F12& a, const F12& b) {
    const int k = threadIdx.x & 31;
    if (k < 12) f12_mul_lane(k, a.c, b.c, lo[k], hi[k]);
    __syncwarp();
    Fq r;
    if (k < 12) r = f12_reduce_lane(k, lo, hi);
    __syncwarp();
    if (k < 12) out.c[k] = r;
    __syncwarp();
  }
};
#endif

// Fq2 coefficient e at w^k in the w basis: (e0 - 9 e1) w^k + e1 w^(k+6)
MZ_HD void f12_put_fq2(F12& l, int k, const Fq2& e) {
  l.c[k] = fe_sub(e.c0, fe_mul_small(e.c1, 9));
  l.c[k + 6] = e.c1;
}

struct TwistPt {  // affine point of the twist curve over Fq2; inf <=> point at infinity
  Fq2 x, y;
  bool inf;
};

// line through t and q (tangent when equal) at P, then t <- t + q.  Leader-only.  A vertical line
// (t == -q) is an element of Fq6 like the dropped denominators: the factor is 1 and t becomes infinity.
MZ_HD void pairing_line_step(F12& l, TwistPt& t, const TwistPt& q, const Fq& xp, const Fq& yp) {
  f12_set_one(l);
  if (q.inf) return;
  if (t.inf) { t = q; return; }
  Fq2 lam;
  if (f2_eq(t.x, q.x)) {
    if (!f2_eq(t.y, q.y) || f2_is_zero(t.y)) { t.inf = true; return; }
    Fq2 xx = f2_sqr(t.x);
    lam = f2_mul(f2_add(f2_dbl(xx), xx), f2_inv(f2_dbl(t.y)));
  } else {
    lam = f2_mul(f2_sub(q.y, t.y), f2_inv(f2_sub(q.x, t.x)));
  }
  // yp - (lambda xp) w + (lambda X_T - Y_T) w^3
#pragma unroll 1
  for (int k = 0; k < 12; k++) l.c[k] = Fq::zero();
  l.c[0] = yp;
  Fq2 a1;
  a1.c0 = fe_neg(fe_mul(lam.c0, xp));
  a1.c1 = fe_neg(fe_mul(lam.c1, xp));
  f12_put_fq2(l, 1, a1);
  f12_put_fq2(l, 3, f2_sub(f2_mul(lam, t.x), t.y));
  Fq2 x3 = f2_sub(f2_sub(f2_sqr(lam), t.x), q.x);
  Fq2 y3 = f2_sub(f2_mul(lam, f2_sub(t.x, x3)), t.y);
  t.x = x3;
  t.y = y3;
}

namespace pairing_const {
// xi = 9 + u;  xi^((p-1)/3), xi^((p-1)/2), xi^((p^2-1)/3), xi^((p^2-1)/2) as raw (non-Montgomery) limbs c0 | c1
MZ_HD constexpr uint32_t frob(int which, int i) {
  constexpr uint32_t t[4][16] = {
      {0x176f553du, 0x99e39557u, 0xc2c3330cu, 0xb78cc310u, 0xf559b143u, 0x4c0bec3cu, 0x4f7911f7u, 0x2fb34798u, 0x640fcba2u, 0x1665d51cu, 0x0b7c9dceu, 0x32ae2a1du, 0xd75a0794u, 0x4ba4cc8bu, 0x61ebae20u, 0x16c9e550u},
      {0x71a0135au, 0xdc540146u, 0xa9c95998u, 0xdbaae0edu, 0xb6e2f9b9u, 0xdc5ec698u, 0x489af5dcu, 0x063cf305u, 0x2623b0e3u, 0x82d37f63u, 0x8fa25bd2u, 0x21807dc9u, 0xec796f2bu, 0x0704b5a7u, 0xac41049au, 0x07c03cbcu},
      {0x607cfd48u, 0xe4bd44e5u, 0xbb966e3du, 0xc28f069fu, 0xe0acccb0u, 0x5e6dd9e7u, 0xe131a029u, 0x30644e72u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
      {0xd87cfd46u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}};
  return t[which][i];
}
MZ_HD Fq2 frob_const(int which) {
  Fq2 r;
  Fq a, b;
  for (int i = 0; i < 8; i++) { a.v[i] = frob(which, i); b.v[i] = frob(which, 8 + i); }
  r.c0 = fe_to_mont(a);
  r.c1 = fe_to_mont(b);
  return r;
}
constexpr int kAteBits = 65;  // ATE_LOOP_COUNT = 29793968203157093288 (bn128.rs:26) has 65 bits
MZ_HD constexpr uint32_t ate_limb(int i) { return i == 0 ? 0xbe763ba8u : i == 1 ? 0x9d797039u : 0x00000001u; }
// (p^12 - 1) / r, 2790 bits, little-endian limbs
#define MZ_FINAL_EXP_LIMBS \
    0xca86f120u, 0x86964b64u, 0xe54523a4u, 0x40a4efb7u, 0x96e84abbu, 0x837fa978u, 0xb9b2b918u, 0x361102b6u, \
    0xf35692dau, 0xc0de81deu, 0xa6c3c760u, 0xbe04c7e8u, 0xd570bb7fu, 0xd766f9c9u, 0x83561841u, 0xc230974du, \
    0xc3be69a3u, 0x5bba1668u, 0x10526294u, 0x7f3811c4u, 0xdadda71cu, 0x29baee7du, 0x145da900u, 0xbf813b8du, \
    0x423f9a2cu, 0x641bbadfu, 0x44eacc5eu, 0xa80bb4eau, 0x14fde37cu, 0xcd656648u, 0x580291d2u, 0x4a0364b9u, \
    0x0826f0ddu, 0xee93dfb1u, 0xc5514724u, 0x6b42db8du, 0x0b0f3785u, 0xbb10cf43u, 0x6f804216u, 0x40494e40u, \
    0xacf3aafbu, 0x55cfe107u, 0xe0ebae87u, 0x2088ec80u, 0x11a337a0u, 0x846a3ed0u, 0x1e3a5195u, 0x48a45a4au, \
    0xdfc50e16u, 0xe5664568u, 0x4c0cc4ebu, 0xab6a4129u, 0xd268c7dau, 0x82d0d602u, 0xed3cc48au, 0x6668449au, \
    0xb2015dfcu, 0x5062cd0fu, 0xb1ddb3d1u, 0x7f2940a8u, 0x2a226448u, 0x77f5b63au, 0x61e443aeu, 0xfef07813u, \
    0x88d5c6c8u, 0xf977870eu, 0x1f676baau, 0x790364a6u, 0xceaddea3u, 0x5887e72eu, 0xa09a1b70u, 0x1377e563u, \
    0x1bd8c3b2u, 0x0c54efeeu, 0xd524d8f7u, 0x3ec3d15au, 0xb2383a5du, 0xdaf15466u, 0xbb94fec0u, 0xe1e30a73u, \
    0x5f3f7be2u, 0x6a1c7101u, 0x6369b1ffu, 0x842d43bfu, 0x107d20bcu, 0x20fddadfu, 0x4b6dc970u, 0x0000002fu, \

static const uint32_t h_final_exp[88] = {MZ_FINAL_EXP_LIMBS};
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t d_final_exp[88] = {MZ_FINAL_EXP_LIMBS};
#endif
MZ_HD uint32_t final_exp_limb(int i) {
#if defined(__CUDA_ARCH__)
  return d_final_exp[i];
#else
  return h_final_exp[i];
#endif
}
}  // namespace pairing_const

// Miller function of the optimal ate pairing: f_{6x+2,Q}(P) * l_{[6x+2]Q,Q1}(P) * l_{[6x+2]Q+Q1,-Q2}(P),
// without the Fq6 denominators.  P = (xp, yp) affine G1 (Montgomery), Q affine G2.  Either at infinity: f = 1.
// `f`, `l` are working storage the executor's lanes can all reach; the other arguments are read by the leader.
template <class Exec>
MZ_HD void pairing_miller(Exec& ex, F12& f, F12& l, const Affine& p, const AffineG2& q) {
  const bool trivial = affine_is_inf(p) || g2_affine_is_inf(q);
  if (ex.leader()) f12_set_one(f);
  ex.sync();
  if (trivial) return;
  TwistPt Q, T;
  Q.x = q.x; Q.y = q.y; Q.inf = false;
  T = Q;
  for (int i = pairing_const::kAteBits - 2; i >= 0; i--) {
    ex.mul(f, f, f);
    if (ex.leader()) pairing_line_step(l, T, T, p.x, p.y);
    ex.sync();
    ex.mul(f, f, l);
    if ((pairing_const::ate_limb(i >> 5) >> (i & 31)) & 1u) {
      if (ex.leader()) pairing_line_step(l, T, Q, p.x, p.y);
      ex.sync();
      ex.mul(f, f, l);
    }
  }
  TwistPt Q1, nQ2;
  if (ex.leader()) {
    Fq2 cx = q.x, cy = q.y;
    cx.c1 = fe_neg(cx.c1);
    cy.c1 = fe_neg(cy.c1);
    Q1.x = f2_mul(cx, pairing_const::frob_const(0));
    Q1.y = f2_mul(cy, pairing_const::frob_const(1));
    Q1.inf = false;
    nQ2.x = f2_mul(q.x, pairing_const::frob_const(2));
    nQ2.y = f2_neg(f2_mul(q.y, pairing_const::frob_const(3)));
    nQ2.inf = false;
    pairing_line_step(l, T, Q1, p.x, p.y);
  }
  ex.sync();
  ex.mul(f, f, l);
  if (ex.leader()) pairing_line_step(l, T, nQ2, p.x, p.y);
  ex.sync();
  ex.mul(f, f, l);
}

// f <- f^((p^12-1)/r)   (`base`, `acc`: working storage like f)
template <class Exec>
MZ_HD void pairing_final_exp(Exec& ex, F12& f, F12& base, F12& acc) {
  if (ex.leader()) { base = f; f12_set_one(acc); }
  ex.sync();
  bool started = false;
  for (int i = 2789; i >= 0; i--) {
    if (started) ex.mul(acc, acc, acc);
    if ((pairing_const::final_exp_limb(i >> 5) >> (i & 31)) & 1u) {
      ex.mul(acc, acc, base);
      started = true;
    }
  }
  if (ex.leader()) f = acc;
  ex.sync();
}

// ---------------------------------------------------------------------------------------------------
// Final exponentiation by parts - the same power (p^12-1)/r = (p^6-1)(p^2+1) * (p^4-p^2+1)/r, about 330
// Fq12 products instead of 4 200:
//   easy part  f^(p^6-1) = conj6(f) / f  (conj6 = w -> -w = Frobenius^6), then ^(p^2+1) by one Frobenius;
//   hard part  (p^4-p^2+1)/r = l0 + l1 p + l2 p^2 + p^3 with, for the BN parameter x = 4965661367192848881,
//              l2 = 6x^2+1, l1 = -36x^3-18x^2-12x+1, l0 = -36x^3-30x^2-18x-2  (checked numerically);
//              after the easy part f is unitary, so negative powers are conj6.
// Validated on the host emulation against the plain power (tests/test_emul_cpu.py); the kernels switch to it
// with MZ_PAIRING_FAST_FINAL_EXP (off until it has run on the GPU once).
namespace pairing_const {
// g_j = xi^((p^j - 1)/6), j = 1, 2, 3, raw limbs c0 | c1: the Frobenius p^j maps w^k to w^k g_j^k
MZ_HD constexpr uint32_t frob12(int j, int i) {
  constexpr uint32_t t[3][16] = {
      {0xdcc9e470u, 0xd60b35dau, 0x292f2176u, 0x5c521e08u, 0x76e68b60u, 0xe8b99fddu, 0x2865a7dfu, 0x1284b71cu,
       0x80f362acu, 0xca5cf05fu, 0x8eeec7e5u, 0x74799277u, 0x12150b8eu, 0xa6327cfeu, 0xb4fae7e6u, 0x246996f3u},
      {0x607cfd49u, 0xe4bd44e5u, 0xbb966e3du, 0xc28f069fu, 0xe0acccb0u, 0x5e6dd9e7u, 0xe131a029u, 0x30644e72u,
       0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
      {0x1ed4a67fu, 0xe86f7d39u, 0xbe55d24au, 0x894cb38du, 0xd0acaa90u, 0xefe9608cu, 0xcc82e4bbu, 0x19dc81cfu,
       0xf4c0c101u, 0x7694aa2bu, 0x97d439ecu, 0x7f03a5e3u, 0x3576139du, 0x06cbeee3u, 0x0be77d73u, 0x00abf8b6u},
  };
  return t[j][i];
}
MZ_HD Fq2 frob12_const(int j) {
  Fq2 r;
  Fq a, b;
  for (int i = 0; i < 8; i++) { a.v[i] = frob12(j, i); b.v[i] = frob12(j, 8 + i); }
  r.c0 = fe_to_mont(a);
  r.c1 = fe_to_mont(b);
  return r;
}
constexpr uint64_t kBnX = 4965661367192848881ull;
}  // namespace pairing_const

// leader-only helpers (out must not alias a)
MZ_HD void f12_conj6(F12& out, const F12& a) {
#pragma unroll 1
  for (int k = 0; k < 12; k++) out.c[k] = (k & 1) ? fe_neg(a.c[k]) : a.c[k];
}
// out = a^(p^j), j = 1..3: sum_k c_k w^k g^k with g^k = (A + 9B... ) written as (a - 9b) + b w^6
MZ_HD void f12_frobenius(F12& out, const F12& a, int j) {
  const Fq2 g = pairing_const::frob12_const(j - 1);
  Fq2 gam = f2_one();
#pragma unroll 1
  for (int k = 0; k < 12; k++) out.c[k] = Fq::zero();
#pragma unroll 1
  for (int k = 0; k < 12; k++) {
    const Fq A = fe_sub(gam.c0, fe_mul_small(gam.c1, 9));
    const Fq cb = fe_mul(a.c[k], gam.c1);
    out.c[k] = fe_add(out.c[k], fe_mul(a.c[k], A));
    if (k + 6 < 12) {
      out.c[k + 6] = fe_add(out.c[k + 6], cb);
    } else {  // w^(k+6) = w^(k-6) (18 w^6 - 82)
      out.c[k] = fe_add(out.c[k], fe_mul_small(cb, 18));
      out.c[k - 6] = fe_sub(out.c[k - 6], fe_mul_small(cb, 82));
    }
    gam = f2_mul(gam, g);
  }
}
// out = n^-1 for n in Fq6 = Fq2[v]/(v^3 - xi), v = w^2, given and returned as an F12 with even coefficients only
MZ_HD void f12_inv_fq6(F12& out, const F12& n) {
  Fq2 xi;
  xi.c0 = fe_mul_small(Fq::one(), 9);
  xi.c1 = Fq::one();
  Fq2 m[3];
  for (int j = 0; j < 3; j++) {
    m[j].c0 = fe_add(n.c[2 * j], fe_mul_small(n.c[2 * j + 6], 9));
    m[j].c1 = n.c[2 * j + 6];
  }
  Fq2 t0 = f2_sub(f2_sqr(m[0]), f2_mul(xi, f2_mul(m[1], m[2])));
  Fq2 t1 = f2_sub(f2_mul(xi, f2_sqr(m[2])), f2_mul(m[0], m[1]));
  Fq2 t2 = f2_sub(f2_sqr(m[1]), f2_mul(m[0], m[2]));
  Fq2 d = f2_add(f2_mul(m[0], t0), f2_mul(xi, f2_add(f2_mul(m[2], t1), f2_mul(m[1], t2))));
  Fq2 di = f2_inv(d);
  Fq2 r[3] = {f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di)};
#pragma unroll 1
  for (int k = 0; k < 12; k++) out.c[k] = Fq::zero();
  for (int j = 0; j < 3; j++) {
    out.c[2 * j] = fe_sub(r[j].c0, fe_mul_small(r[j].c1, 9));
    out.c[2 * j + 6] = r[j].c1;
  }
}

// out = a^e for a small public exponent e >= 1 (out must not alias a)
template <class Exec>
MZ_HD void f12_pow_u64(Exec& ex, F12& out, const F12& a, uint64_t e) {
  int top = 63;
  while (top > 0 && !((e >> top) & 1ull)) top--;
  if (ex.leader()) out = a;
  ex.sync();
  for (int bit = top - 1; bit >= 0; bit--) {
    ex.mul(out, out, out);
    if ((e >> bit) & 1ull) ex.mul(out, out, a);
  }
}

// f <- f^((p^12-1)/r); w: nine F12 of working storage every lane of the executor can reach
template <class Exec>
MZ_HD void pairing_final_exp_fast(Exec& ex, F12& f, F12* w) {
  // easy part
  if (ex.leader()) f12_conj6(w[0], f);
  ex.sync();
  ex.mul(w[1], f, w[0]);  // norm to Fq6 (odd coefficients vanish)
  if (ex.leader()) f12_inv_fq6(w[2], w[1]);
  ex.sync();
  ex.mul(w[1], w[0], w[2]);  // 1 / f
  ex.mul(w[1], w[0], w[1]);  // f^(p^6 - 1)
  if (ex.leader()) f12_frobenius(w[0], w[1], 2);
  ex.sync();
  ex.mul(w[1], w[0], w[1]);  // ^(p^2 + 1): unitary from here on; w[1] = g
  // hard part
  f12_pow_u64(ex, w[2], w[1], pairing_const::kBnX);  // g^x
  f12_pow_u64(ex, w[3], w[2], pairing_const::kBnX);  // g^(x^2)
  f12_pow_u64(ex, w[4], w[3], pairing_const::kBnX);  // g^(x^3)
  f12_pow_u64(ex, w[5], w[3], 6);
  ex.mul(w[5], w[5], w[1]);                          // y2 = g^(6x^2 + 1)
  f12_pow_u64(ex, w[7], w[4], 36);                   // g^(36 x^3)
  f12_pow_u64(ex, w[0], w[3], 18);
  ex.mul(w[6], w[7], w[0]);
  f12_pow_u64(ex, w[0], w[2], 12);
  ex.mul(w[6], w[6], w[0]);                          // g^(36x^3 + 18x^2 + 12x)
  if (ex.leader()) f12_conj6(w[0], w[6]);
  ex.sync();
  ex.mul(w[6], w[0], w[1]);                          // y1 = g^(-36x^3 - 18x^2 - 12x + 1)
  f12_pow_u64(ex, w[0], w[3], 30);
  ex.mul(w[7], w[7], w[0]);
  f12_pow_u64(ex, w[0], w[2], 18);
  ex.mul(w[7], w[7], w[0]);
  ex.mul(w[7], w[7], w[1]);
  ex.mul(w[7], w[7], w[1]);                          // g^(36x^3 + 30x^2 + 18x + 2)
  if (ex.leader()) f12_conj6(w[8], w[7]);            // y0
  ex.sync();
  if (ex.leader()) f12_frobenius(w[0], w[6], 1);
  ex.sync();
  ex.mul(w[8], w[8], w[0]);                          // y0 * y1^p
  if (ex.leader()) f12_frobenius(w[0], w[5], 2);
  ex.sync();
  ex.mul(w[8], w[8], w[0]);                          // * y2^(p^2)
  if (ex.leader()) f12_frobenius(w[0], w[1], 3);
  ex.sync();
  ex.mul(w[8], w[8], w[0]);                          // * g^(p^3)
  if (ex.leader()) f = w[8];
  ex.sync();
}

}  // namespace mz

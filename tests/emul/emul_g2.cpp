// Host build of the device G2 header (emulated carry flag). TEST INFRASTRUCTURE ONLY.
#include "../../myzkp_b200/csrc/g2.cuh"
#include <string.h>
using namespace mz;
static Fq load_mont(const uint32_t* raw) { Fq a; for (int i = 0; i < 8; i++) a.v[i] = raw[i]; return fe_to_mont(a); }
static void store_raw(const Fq& a, uint32_t* out) { Fq r = fe_from_mont(a); for (int i = 0; i < 8; i++) out[i] = r.v[i]; }
extern "C" {
// base / out: 32 raw limbs (x.c0, x.c1, y.c0, y.c1), canonical non-Montgomery; k: 8 raw limbs
void emul_g2_scalar_mul(const uint32_t* base, const uint32_t* k, uint32_t* out) {
  AffineG2 b;
  b.x.c0 = load_mont(base); b.x.c1 = load_mont(base + 8); b.y.c0 = load_mont(base + 16); b.y.c1 = load_mont(base + 24);
  AffineG2 r = g2_scalar_mul(b, k);
  store_raw(r.x.c0, out); store_raw(r.x.c1, out + 8); store_raw(r.y.c0, out + 16); store_raw(r.y.c1, out + 24);
}
// acc (Jacobian given as affine or all-zero infinity) += q, returned affine - exercises the special cases
void emul_g2_add(const uint32_t* p, const uint32_t* q, uint32_t* out) {
  AffineG2 a, b;
  a.x.c0 = load_mont(p); a.x.c1 = load_mont(p + 8); a.y.c0 = load_mont(p + 16); a.y.c1 = load_mont(p + 24);
  b.x.c0 = load_mont(q); b.x.c1 = load_mont(q + 8); b.y.c0 = load_mont(q + 16); b.y.c1 = load_mont(q + 24);
  JacG2 acc = g2_jac_inf();
  g2_jac_madd(acc, a);
  // give acc a non-trivial Z so the general formulas are exercised: acc = 2a - a when a is finite
  if (!g2_affine_is_inf(a)) { g2_jac_dbl(acc); AffineG2 na = a; na.y = f2_neg(a.y); g2_jac_madd(acc, na); }
  g2_jac_madd(acc, b);
  AffineG2 r = g2_jac_to_affine(acc);
  store_raw(r.x.c0, out); store_raw(r.x.c1, out + 8); store_raw(r.y.c0, out + 16); store_raw(r.y.c1, out + 24);
}
// [k1] p + [k2] q through the Jacobian + Jacobian addition (k = 0 gives infinity operands)
void emul_g2_lincomb(const uint32_t* p, const uint32_t* k1, const uint32_t* q, const uint32_t* k2, uint32_t* out) {
  AffineG2 a, b;
  a.x.c0 = load_mont(p); a.x.c1 = load_mont(p + 8); a.y.c0 = load_mont(p + 16); a.y.c1 = load_mont(p + 24);
  b.x.c0 = load_mont(q); b.x.c1 = load_mont(q + 8); b.y.c0 = load_mont(q + 16); b.y.c1 = load_mont(q + 24);
  JacG2 x = g2_scalar_mul_jac(a, k1), y = g2_scalar_mul_jac(b, k2);
  g2_jac_add(x, y);
  AffineG2 r = g2_jac_to_affine(x);
  store_raw(r.x.c0, out); store_raw(r.x.c1, out + 8); store_raw(r.y.c0, out + 16); store_raw(r.y.c1, out + 24);
}
void emul_fq2_inv(const uint32_t* a, uint32_t* out) {
  Fq2 x; x.c0 = load_mont(a); x.c1 = load_mont(a + 8);
  Fq2 r = f2_inv(x);
  store_raw(r.c0, out); store_raw(r.c1, out + 8);
}
}

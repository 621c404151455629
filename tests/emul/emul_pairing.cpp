// Host build of the device pairing header (sequential executor). TEST INFRASTRUCTURE ONLY.
#include "../../myzkp_b200/csrc/pairing.cuh"
using namespace mz;
static Fq load_mont(const uint32_t* raw) { Fq a; for (int i = 0; i < 8; i++) a.v[i] = raw[i]; return fe_to_mont(a); }
static void store_raw(const Fq& a, uint32_t* out) { Fq r = fe_from_mont(a); for (int i = 0; i < 8; i++) out[i] = r.v[i]; }
extern "C" {
// a, b, out: 12 x 8 raw limbs (coefficients of w^0..w^11, canonical)
void emul_f12_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  F12 x, y, r;
  for (int k = 0; k < 12; k++) { x.c[k] = load_mont(a + 8 * k); y.c[k] = load_mont(b + 8 * k); }
  SeqExec ex;
  ex.mul(r, x, y);
  for (int k = 0; k < 12; k++) store_raw(r.c[k], out + 8 * k);
}
// g1: 16 raw limbs (x, y; zeros = infinity), g2: 32 raw limbs; out: 96 raw limbs.  final = 0: Miller value only
// out = a^((p^12-1)/r) by the plain power (fast == 0) or by parts (fast == 1); a: 96 raw limbs
void emul_final_exp(const uint32_t* a, int fast, uint32_t* out) {
  F12 f, base, acc;
  for (int k = 0; k < 12; k++) f.c[k] = load_mont(a + 8 * k);
  SeqExec ex;
  if (fast) { F12 w[9]; pairing_final_exp_fast(ex, f, w); }
  else pairing_final_exp(ex, f, base, acc);
  for (int k = 0; k < 12; k++) store_raw(f.c[k], out + 8 * k);
}
void emul_pairing(const uint32_t* g1, const uint32_t* g2, int final, uint32_t* out) {
  Affine p; p.x = load_mont(g1); p.y = load_mont(g1 + 8);
  AffineG2 q; q.x.c0 = load_mont(g2); q.x.c1 = load_mont(g2 + 8); q.y.c0 = load_mont(g2 + 16); q.y.c1 = load_mont(g2 + 24);
  SeqExec ex;
  F12 f, l, base, acc;
  pairing_miller(ex, f, l, p, q);
  if (final == 1) pairing_final_exp(ex, f, base, acc);
  if (final == 2) { F12 w[9]; pairing_final_exp_fast(ex, f, w); }
  for (int k = 0; k < 12; k++) store_raw(f.c[k], out + 8 * k);
}
}

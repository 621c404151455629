// Host build of the device G1 header (emulated carry flag). TEST INFRASTRUCTURE ONLY.
#include "../../myzkp_b200/csrc/g1.cuh"
#include <string.h>
using namespace mz;
extern "C" {
// all points are arrays of u32 limbs in Montgomery form: XYZZ = 32 limbs, Affine = 16 limbs
void emul_g1_madd(uint32_t* acc, const uint32_t* q) { XYZZ a; Affine b; memcpy(&a, acc, 128); memcpy(&b, q, 64); xyzz_madd(a, b); memcpy(acc, &a, 128); }
void emul_g1_add(uint32_t* acc, const uint32_t* q) { XYZZ a, b; memcpy(&a, acc, 128); memcpy(&b, q, 128); xyzz_add(a, b); memcpy(acc, &a, 128); }
void emul_g1_dbl(uint32_t* acc) { XYZZ a; memcpy(&a, acc, 128); xyzz_dbl(a); memcpy(acc, &a, 128); }
void emul_g1_to_affine(const uint32_t* p, uint32_t* out) { XYZZ a; memcpy(&a, p, 128); Affine r = xyzz_to_affine(a); memcpy(out, &r, 64); }
void emul_g1_jac_dbl(uint32_t* p) { Jac a; memcpy(&a, p, 96); jac_dbl(a); memcpy(p, &a, 96); }
void emul_g1_neg(const uint32_t* p, uint32_t* out) { Affine a; memcpy(&a, p, 64); Affine r = affine_neg(a); memcpy(out, &r, 64); }
}

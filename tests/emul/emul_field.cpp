// Host build of the device field/group headers with an emulated carry flag.
// TEST INFRASTRUCTURE ONLY (never linked into libmyzkp_b200.so).
#include "../../myzkp_b200/csrc/field.cuh"
#include "../../myzkp_b200/csrc/inv.cuh"
using namespace mz;
extern "C" {
#define BIN(name, T, fn) void name(const uint32_t* a, const uint32_t* b, uint32_t* o) { T x, y; for (int i=0;i<8;i++){x.v[i]=a[i]; y.v[i]=b[i];} T r = fn(x,y); for(int i=0;i<8;i++) o[i]=r.v[i]; }
#define UN(name, T, fn) void name(const uint32_t* a, uint32_t* o) { T x; for (int i=0;i<8;i++){x.v[i]=a[i];} T r = fn(x); for(int i=0;i<8;i++) o[i]=r.v[i]; }
BIN(emul_fq_mul, Fq, fe_mul) BIN(emul_fq_add, Fq, fe_add) BIN(emul_fq_sub, Fq, fe_sub)
BIN(emul_fr_mul, Fr, fe_mul) BIN(emul_fr_add, Fr, fe_add) BIN(emul_fr_sub, Fr, fe_sub)
UN(emul_fq_sqr, Fq, fe_sqr) UN(emul_fr_sqr, Fr, fe_sqr)
UN(emul_fq_inv_bingcd, Fq, fe_inv_bingcd) UN(emul_fr_inv_bingcd, Fr, fe_inv_bingcd)
UN(emul_fq_inv_safegcd, Fq, fe_inv_safegcd) UN(emul_fr_inv_safegcd, Fr, fe_inv_safegcd)
UN(emul_fq_neg, Fq, fe_neg) UN(emul_fq_inv, Fq, fe_inv) UN(emul_fq_to_mont, Fq, fe_to_mont) UN(emul_fq_from_mont, Fq, fe_from_mont)
UN(emul_fr_neg, Fr, fe_neg) UN(emul_fr_inv, Fr, fe_inv) UN(emul_fr_to_mont, Fr, fe_to_mont) UN(emul_fr_from_mont, Fr, fe_from_mont)
#define QUAD(name, T, fn) void name(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* o) { T x, y, z, w; for (int i=0;i<8;i++){x.v[i]=a[i]; y.v[i]=b[i]; z.v[i]=c[i]; w.v[i]=d[i];} T r = fn(x,y,z,w); for(int i=0;i<8;i++) o[i]=r.v[i]; }
QUAD(emul_fq_mul2, Fq, fe_mul2) QUAD(emul_fr_mul2, Fr, fe_mul2) QUAD(emul_fq_mul_sub_mul, Fq, fe_mul_sub_mul) QUAD(emul_fr_mul_sub_mul, Fr, fe_mul_sub_mul)
BIN(emul_fq_mul_k, Fq, fe_mul_k) BIN(emul_fr_mul_k, Fr, fe_mul_k) QUAD(emul_fq_mul2_k, Fq, fe_mul2_k) QUAD(emul_fr_mul2_k, Fr, fe_mul2_k)
int emul_fq_is_canonical(const uint32_t* a) { Fq x; for (int i=0;i<8;i++) x.v[i]=a[i]; return fe_is_canonical(x); }
}

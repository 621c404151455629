// 254-bit prime-field arithmetic for BN128 (Fq = base field p, Fr = scalar field r)
// as 8 x 32-bit limbs in Montgomery form (R = 2^256), written for sm_100a.
//
// Replaces the reference's BigInt-backed FiniteFieldElement ops
// (myzkp/src/modules/algebra/field.rs:157-183 add/sub/mul + `% modulus`,
//  field.rs:210-237 inverse) on the KZG hot path.  Values are always kept
// canonical in [0, m), so equality of limbs == the reference's equality of
// sanitized values (field.rs:290-294).
//
// The multiply is an even/odd-column CIOS: 64-bit partial products are
// accumulated with mad.lo.cc/madc.hi.cc pairs, which ptxas fuses into
// IMAD.WIDE.U32(.X) with predicate carries on sm_100a (checked with
// cuobjdump -sass: ~120 IMAD.WIDE + 8 IMAD.HI + 9 IMAD per multiply).
//
// The same header compiles for the host (g++) with an emulated carry flag so
// that tests can run the exact device algorithms on the CPU
// (tests/emul/); the product library never uses that path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MZ_HD __host__ __device__ __forceinline__
#define MZ_D __device__ __forceinline__
#else
#define MZ_HD inline
#define MZ_D inline
#endif

namespace mz {

// ---------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define MZ_ASM asm volatile
MZ_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; MZ_ASM("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MZ_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; MZ_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MZ_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; MZ_ASM("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MZ_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; MZ_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MZ_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; MZ_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MZ_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; MZ_ASM("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MZ_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
MZ_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
MZ_D void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
  MZ_ASM("{.reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t;}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
MZ_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MZ_ASM("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MZ_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MZ_ASM("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MZ_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MZ_ASM("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MZ_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MZ_ASM("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MZ_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MZ_ASM("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// host emulation of the PTX condition-code register (tests only)
static thread_local uint32_t g_cf = 0;
inline uint32_t emu_add(uint32_t a, uint32_t b, uint32_t cin, bool set) {
  uint64_t s = (uint64_t)a + b + cin;
  if (set) g_cf = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t emu_sub(uint32_t a, uint32_t b, uint32_t bin, bool set) {
  uint64_t s = (uint64_t)a - b - bin;
  if (set) g_cf = (uint32_t)((s >> 32) & 1);  // borrow
  return (uint32_t)s;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, g_cf, true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add(a, b, g_cf, false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, g_cf, true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub(a, b, g_cf, false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
  uint64_t t = (uint64_t)a * b;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
}
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, g_cf, true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, g_cf, true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, g_cf, false); }
#endif

// ---------------------------------------------------------------------------
// moduli (curve/bn128.rs:19-22, field.rs:428-431)
// ---------------------------------------------------------------------------
#define MZ_LIMBS8(name, a0, a1, a2, a3, a4, a5, a6, a7)                          \
  static MZ_HD constexpr uint32_t name(int i) {                                  \
    return i == 0 ? a0 : i == 1 ? a1 : i == 2 ? a2 : i == 3 ? a3 : i == 4 ? a4  \
         : i == 5 ? a5 : i == 6 ? a6 : a7;                                       \
  }

struct FqParams {  // base field p
  MZ_LIMBS8(mod, 0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
            0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
  static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
  // R = 2^256 mod p (Montgomery one)
  MZ_LIMBS8(one, 0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
            0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
  // R^2 = 2^512 mod p
  MZ_LIMBS8(r2, 0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
            0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u)
};
struct FrParams {  // scalar field r
  MZ_LIMBS8(mod, 0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
            0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
  static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
  MZ_LIMBS8(one, 0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
            0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
  MZ_LIMBS8(r2, 0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
            0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u)
};

template <class PR>
MZ_HD constexpr uint32_t mod_limb(int i) { return PR::mod(i); }

// ---------------------------------------------------------------------------
// field element
// ---------------------------------------------------------------------------
template <class PR>
struct Fe {
  uint32_t v[8];

  static MZ_HD Fe zero() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  static MZ_HD Fe one() {  // Montgomery form of 1
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = PR::one(i);
    return r;
  }
  static MZ_HD Fe r2() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = PR::r2(i);
    return r;
  }
  MZ_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  MZ_HD bool operator==(const Fe& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  MZ_HD bool operator!=(const Fe& b) const { return !(*this == b); }
};

// r = a - m if a >= m else a   (a < 2m)
template <class PR>
MZ_HD void fe_reduce_once(Fe<PR>& a) {
  uint32_t t[8];
  t[0] = sub_cc(a.v[0], mod_limb<PR>(0));
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = subc_cc(a.v[i], mod_limb<PR>(i));
  uint32_t borrow = subc(0, 0);  // 0xffffffff if a < m
#pragma unroll
  for (int i = 0; i < 8; i++) a.v[i] = borrow ? a.v[i] : t[i];
}

// true iff a (any 256-bit value) < modulus
template <class PR>
MZ_HD bool fe_is_canonical(const Fe<PR>& a) {
  (void)sub_cc(a.v[0], mod_limb<PR>(0));
#pragma unroll
  for (int i = 1; i < 8; i++) (void)subc_cc(a.v[i], mod_limb<PR>(i));
  return subc(0, 0) != 0;
}

template <class PR>
MZ_HD Fe<PR> fe_add(const Fe<PR>& a, const Fe<PR>& b) {
  Fe<PR> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[7] = addc(a.v[7], b.v[7]);  // both < 2^254: no carry out
  fe_reduce_once(r);
  return r;
}

template <class PR>
MZ_HD Fe<PR> fe_sub(const Fe<PR>& a, const Fe<PR>& b) {
  Fe<PR> r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0, 0);  // all-ones if a < b
  r.v[0] = add_cc(r.v[0], borrow & mod_limb<PR>(0));
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], borrow & mod_limb<PR>(i));
  r.v[7] = addc(r.v[7], borrow & mod_limb<PR>(7));
  return r;
}

template <class PR>
MZ_HD Fe<PR> fe_neg(const Fe<PR>& a) {
  Fe<PR> r;
  uint32_t nz = a.is_zero() ? 0u : 0xffffffffu;
  r.v[0] = sub_cc(mod_limb<PR>(0) & nz, a.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = subc_cc(mod_limb<PR>(i) & nz, a.v[i]);
  r.v[7] = subc(mod_limb<PR>(7) & nz, a.v[7]);
  return r;
}

template <class PR>
MZ_HD Fe<PR> fe_dbl(const Fe<PR>& a) { return fe_add(a, a); }

// --- Montgomery multiply ----------------------------------------------------
// Accumulators: E[k] holds column k, O[k] holds column k+1 of the running sum.
// One row adds a*b_i and m*modulus, after which column 0 is zero and the
// roles swap (old O becomes the new E; old E shifted by two limbs the new O).
namespace detail {

// acc[j..j+1] += a[j] * bi for even j, one carry chain, carry left in CF
template <class A>
MZ_HD void row_mad_even(uint32_t* acc, const A& a, uint32_t bi) {
  acc[0] = mad_lo_cc(a(0), bi, acc[0]);
  acc[1] = madc_hi_cc(a(0), bi, acc[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    acc[j] = madc_lo_cc(a(j), bi, acc[j]);
    acc[j + 1] = madc_hi_cc(a(j), bi, acc[j + 1]);
  }
}
// acc[j-1..j] += a[j] * bi for odd j (acc is the odd-column array)
template <class A>
MZ_HD void row_mad_odd(uint32_t* acc, const A& a, uint32_t bi) {
  acc[0] = mad_lo_cc(a(1), bi, acc[0]);
  acc[1] = madc_hi_cc(a(1), bi, acc[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    acc[j] = madc_lo_cc(a(j + 1), bi, acc[j]);
    acc[j + 1] = madc_hi_cc(a(j + 1), bi, acc[j + 1]);
  }
}
// new odd-column array from the old even array shifted down two limbs, plus
// the odd-limb products; consumes the incoming CF (from folding old column 1)
template <class A>
MZ_HD void row_shift_mad_odd(uint32_t* acc, const A& a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < 6; j += 2) {
    acc[j] = madc_lo_cc(a(j + 1), bi, acc[j + 2]);
    acc[j + 1] = madc_hi_cc(a(j + 1), bi, acc[j + 3]);
  }
  acc[6] = madc_lo_cc(a(7), bi, 0);
  acc[7] = madc_hi(a(7), bi, 0);
}

template <class PR>
struct ModAcc {
  MZ_HD uint32_t operator()(int i) const { return mod_limb<PR>(i); }
};
struct ArrAcc {
  const uint32_t* p;
  MZ_HD uint32_t operator()(int i) const { return p[i]; }
};

template <class PR>
MZ_HD void mont_row(uint32_t* E, uint32_t* O, const ArrAcc& a, uint32_t bi, bool first) {
  if (first) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      mul_wide(a(j + 1), bi, O[j], O[j + 1]);
      mul_wide(a(j), bi, E[j], E[j + 1]);
    }
  } else {
    // E is the previous row's odd array (already in place); O is the previous
    // even array whose column 0 is zero and column 1 (O[1]) folds into E[0].
    E[0] = add_cc(E[0], O[1]);
    row_shift_mad_odd(O, a, bi);
    row_mad_even(E, a, bi);
    O[7] = addc(O[7], 0);
  }
  uint32_t m = mul_lo(E[0], PR::INV);
  row_mad_odd(O, ModAcc<PR>(), m);  // cannot carry out (sum < 2^256 * 2^32)
  row_mad_even(E, ModAcc<PR>(), m);
  O[7] = addc(O[7], 0);
}

}  // namespace detail

template <class PR>
MZ_HD Fe<PR> fe_mul(const Fe<PR>& a, const Fe<PR>& b) {
  uint32_t E[8], O[8];
  detail::ArrAcc aa{a.v};
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    detail::mont_row<PR>(E, O, aa, b.v[i], i == 0);
    detail::mont_row<PR>(O, E, aa, b.v[i + 1], false);
  }
  // result column k = E[k] + O[k+1]
  Fe<PR> r;
  r.v[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(E[i], O[i + 1]);
  r.v[7] = addc(E[7], 0);
  fe_reduce_once(r);
  return r;
}

// --- sum of two products with ONE reduction -----------------------------------
// (a*b + c*d) R^-1 mod m: every row adds a*b_i, c*d_i and m*modulus to the same accumulators,
// so the pair costs 192 wide multiplies instead of 2 x 128.  Bounds (inputs < modulus < 2^254):
// the running sum after a row's shift is < 3 * 2^254 (1 + 2^-32) < 2^256 and before it
// < 2^288, which is what the (E, O) arrays hold; the final value is
// < (2 m^2 + 2^256 m) / 2^256 < 1.38 m, so one conditional subtraction suffices.
namespace detail {
template <class PR>
MZ_HD void mont_row2(uint32_t* E, uint32_t* O, const ArrAcc& a, uint32_t bi, const ArrAcc& c, uint32_t di,
                     bool first) {
  if (first) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      mul_wide(a(j + 1), bi, O[j], O[j + 1]);
      mul_wide(a(j), bi, E[j], E[j + 1]);
    }
  } else {
    E[0] = add_cc(E[0], O[1]);
    row_shift_mad_odd(O, a, bi);
    row_mad_even(E, a, bi);
    O[7] = addc(O[7], 0);
  }
  row_mad_odd(O, c, di);  // the total stays < 2^288: no carry out of column 8
  row_mad_even(E, c, di);
  O[7] = addc(O[7], 0);
  uint32_t m = mul_lo(E[0], PR::INV);
  row_mad_odd(O, ModAcc<PR>(), m);
  row_mad_even(E, ModAcc<PR>(), m);
  O[7] = addc(O[7], 0);
}
}  // namespace detail

template <class PR>
MZ_HD Fe<PR> fe_mul2(const Fe<PR>& a, const Fe<PR>& b, const Fe<PR>& c, const Fe<PR>& d) {
  uint32_t E[8], O[8];
  detail::ArrAcc aa{a.v}, cc{c.v};
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    detail::mont_row2<PR>(E, O, aa, b.v[i], cc, d.v[i], i == 0);
    detail::mont_row2<PR>(O, E, aa, b.v[i + 1], cc, d.v[i + 1], false);
  }
  Fe<PR> r;
  r.v[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(E[i], O[i + 1]);
  r.v[7] = addc(E[7], 0);
  fe_reduce_once(r);
  return r;
}
// a*b - c*d
template <class PR>
MZ_HD Fe<PR> fe_mul_sub_mul(const Fe<PR>& a, const Fe<PR>& b, const Fe<PR>& c, const Fe<PR>& d) {
  return fe_mul2(a, b, fe_neg(c), d);
}

// --- Montgomery square --------------------------------------------------------
// a^2 costs 36 wide products instead of 64: the 28 products a_i*a_j (i < j) are summed once
// and doubled, the 8 squares a_i^2 are added on top.  The 512-bit result S is then reduced
// by 8 rows of m*modulus (the second half of mont_row; with no a*b_i products the shift of
// the even array rides on the odd-limb products of m*modulus), and S's high half is added
// at the end: 100 wide multiplies against 128 for fe_mul.
//
// Off-diagonal sum T, in two accumulators so every row is one carry chain per parity:
//   X[k] = column k+1: products whose low column i+j is odd   (columns 1..14)
//   Y[k] = column k+2: products whose low column i+j is even  (columns 2..13)
// A chain that ends on a column an earlier row already wrote may carry out; the next
// column is untouched at that point and receives the carry bit.  A chain that ends on
// an untouched column adds hi(product) <= 2^32 - 2 to zero and cannot carry out.
namespace detail {
MZ_HD uint32_t shl1(uint32_t lo, uint32_t hi) { return (hi << 1) | (lo >> 31); }

// one reduction row without multiplicand products: (E, O) are the previous row's (odd, even) arrays
template <class PR>
MZ_HD void redc_row(uint32_t* E, uint32_t* O) {
  E[0] = add_cc(E[0], O[1]);
  uint32_t m = mul_lo(E[0], PR::INV);
  row_shift_mad_odd(O, ModAcc<PR>(), m);
  row_mad_even(E, ModAcc<PR>(), m);
  O[7] = addc(O[7], 0);
}
}  // namespace detail

// Montgomery reduction of a 16-limb value S < 2^256 * modulus: 8 rows of m * modulus on the low half,
// then the high half is added (result < modulus + S / 2^256 < 2 * modulus) and reduced once.
namespace detail {
template <class PR>
MZ_HD Fe<PR> redc_wide(const uint32_t* S) {
  uint32_t E[8], O[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { E[i] = S[i]; O[i] = 0; }
  {
    uint32_t m = mul_lo(E[0], PR::INV);
    row_mad_odd(O, ModAcc<PR>(), m);
    row_mad_even(E, ModAcc<PR>(), m);
    O[7] = addc(O[7], 0);
  }
  redc_row<PR>(O, E);
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    redc_row<PR>(E, O);
    redc_row<PR>(O, E);
  }
  Fe<PR> r;
  r.v[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(E[i], O[i + 1]);
  r.v[7] = addc(E[7], 0);
  r.v[0] = add_cc(r.v[0], S[8]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], S[8 + i]);
  r.v[7] = addc(r.v[7], S[15]);
  fe_reduce_once(r);
  return r;
}

// 4 x 4 limb product (16 wide multiplies).  P[k] = column k, Q[k] = column k + 1: a product whose low
// column is even goes to P, odd to Q, so every row is one carry chain per accumulator.  A chain that
// ends on a column an earlier row wrote may carry out into the next (still untouched) column; a chain
// that ends on an untouched column adds hi(product) <= 2^32 - 2 to a carry bit and cannot carry out.
MZ_HD void mul4(const uint32_t* a, const uint32_t* b, uint32_t* r) {
  uint32_t P[8], Q[7];
  // row 0
  mul_wide(a[0], b[0], P[0], P[1]);
  mul_wide(a[2], b[0], P[2], P[3]);
  mul_wide(a[1], b[0], Q[0], Q[1]);
  mul_wide(a[3], b[0], Q[2], Q[3]);
  // row 1: a0,a2 -> columns 1,3 (Q0..Q3, all written -> carry to Q4); a1,a3 -> columns 2,4 (P2,P3 written; P4,P5 new)
  Q[0] = mad_lo_cc(a[0], b[1], Q[0]);  Q[1] = madc_hi_cc(a[0], b[1], Q[1]);
  Q[2] = madc_lo_cc(a[2], b[1], Q[2]); Q[3] = madc_hi_cc(a[2], b[1], Q[3]);
  Q[4] = addc(0, 0);
  P[2] = mad_lo_cc(a[1], b[1], P[2]);  P[3] = madc_hi_cc(a[1], b[1], P[3]);
  P[4] = madc_lo_cc(a[3], b[1], 0);    P[5] = madc_hi(a[3], b[1], 0);
  // row 2: a0,a2 -> columns 2,4 (P2..P5 written -> carry to P6); a1,a3 -> columns 3,5 (Q2,Q3 written, Q4 carry bit, Q5 new)
  P[2] = mad_lo_cc(a[0], b[2], P[2]);  P[3] = madc_hi_cc(a[0], b[2], P[3]);
  P[4] = madc_lo_cc(a[2], b[2], P[4]); P[5] = madc_hi_cc(a[2], b[2], P[5]);
  P[6] = addc(0, 0);
  Q[2] = mad_lo_cc(a[1], b[2], Q[2]);  Q[3] = madc_hi_cc(a[1], b[2], Q[3]);
  Q[4] = madc_lo_cc(a[3], b[2], Q[4]); Q[5] = madc_hi(a[3], b[2], 0);
  // row 3: a0,a2 -> columns 3,5 (Q2..Q5 written -> carry to Q6); a1,a3 -> columns 4,6 (P4,P5 written, P6 carry bit, P7 new)
  Q[2] = mad_lo_cc(a[0], b[3], Q[2]);  Q[3] = madc_hi_cc(a[0], b[3], Q[3]);
  Q[4] = madc_lo_cc(a[2], b[3], Q[4]); Q[5] = madc_hi_cc(a[2], b[3], Q[5]);
  Q[6] = addc(0, 0);
  P[4] = mad_lo_cc(a[1], b[3], P[4]);  P[5] = madc_hi_cc(a[1], b[3], P[5]);
  P[6] = madc_lo_cc(a[3], b[3], P[6]); P[7] = madc_hi(a[3], b[3], 0);
  // r = P + Q * 2^32 (< 2^256)
  r[0] = P[0];
  r[1] = add_cc(P[1], Q[0]);
#pragma unroll
  for (int k = 2; k < 7; k++) r[k] = addc_cc(P[k], Q[k - 1]);
  r[7] = addc(P[7], Q[6]);
}

// d = |x - y| over 4 limbs; returns all-ones if x < y
MZ_HD uint32_t absdiff4(const uint32_t* x, const uint32_t* y, uint32_t* d) {
  d[0] = sub_cc(x[0], y[0]);
  d[1] = subc_cc(x[1], y[1]);
  d[2] = subc_cc(x[2], y[2]);
  d[3] = subc_cc(x[3], y[3]);
  const uint32_t neg = subc(0, 0);  // all-ones on borrow
  d[0] = sub_cc(d[0] ^ neg, neg);
  d[1] = subc_cc(d[1] ^ neg, neg);
  d[2] = subc_cc(d[2] ^ neg, neg);
  d[3] = subc(d[3] ^ neg, neg);
  return neg;
}

// S (16 limbs) = a * b by one level of subtractive Karatsuba: 3 x 16 wide multiplies instead of 64.
// a = a0 + a1 2^128, b = b0 + b1 2^128:  a0 b1 + a1 b0 = a0 b0 + a1 b1 + (a0 - a1)(b1 - b0).
// The additions run on the integer-add pipe, which the multiply-bound kernels leave idle.
MZ_HD void mul8_karatsuba(const uint32_t* a, const uint32_t* b, uint32_t* S) {
  uint32_t z0[8], z2[8], zm[8], da[4], db[4];
  mul4(a, b, z0);
  mul4(a + 4, b + 4, z2);
  const uint32_t neg = absdiff4(a, a + 4, da) ^ absdiff4(b + 4, b, db);  // sign of (a0 - a1)(b1 - b0)
  mul4(da, db, zm);
  // mid = z0 + z2 +- zm as a 9-limb two's-complement value (true value in [0, 2^257))
  uint32_t mid[9];
  mid[0] = add_cc(z0[0], z2[0]);
#pragma unroll
  for (int k = 1; k < 8; k++) mid[k] = addc_cc(z0[k], z2[k]);
  mid[8] = addc(0, 0);
  (void)add_cc(neg, neg);  // CF = 1 when subtracting: -zm = ~zm + 1
#pragma unroll
  for (int k = 0; k < 8; k++) mid[k] = addc_cc(mid[k], zm[k] ^ neg);
  mid[8] = addc(mid[8], neg);
  // S = z0 + mid 2^128 + z2 2^256
#pragma unroll
  for (int k = 0; k < 4; k++) S[k] = z0[k];
  S[4] = add_cc(z0[4], mid[0]);
  S[5] = addc_cc(z0[5], mid[1]);
  S[6] = addc_cc(z0[6], mid[2]);
  S[7] = addc_cc(z0[7], mid[3]);
  S[8] = addc_cc(z2[0], mid[4]);
  S[9] = addc_cc(z2[1], mid[5]);
  S[10] = addc_cc(z2[2], mid[6]);
  S[11] = addc_cc(z2[3], mid[7]);
  S[12] = addc_cc(z2[4], mid[8]);
  S[13] = addc_cc(z2[5], 0);
  S[14] = addc_cc(z2[6], 0);
  S[15] = addc(z2[7], 0);
}
}  // namespace detail

// Karatsuba variants of the multiply: 48 + 64 wide multiplies (fe_mul: 120 + 8 high halves),
// 96 + 64 for the two-product form (192).
template <class PR>
MZ_HD Fe<PR> fe_mul_k(const Fe<PR>& a, const Fe<PR>& b) {
  uint32_t S[16];
  detail::mul8_karatsuba(a.v, b.v, S);
  return detail::redc_wide<PR>(S);
}
template <class PR>
MZ_HD Fe<PR> fe_mul2_k(const Fe<PR>& a, const Fe<PR>& b, const Fe<PR>& c, const Fe<PR>& d) {
  uint32_t S[16], T[16];
  detail::mul8_karatsuba(a.v, b.v, S);
  detail::mul8_karatsuba(c.v, d.v, T);
  S[0] = add_cc(S[0], T[0]);
#pragma unroll
  for (int k = 1; k < 15; k++) S[k] = addc_cc(S[k], T[k]);
  S[15] = addc(S[15], T[15]);  // a b + c d < 2^509
  return detail::redc_wide<PR>(S);
}

template <class PR>
MZ_HD Fe<PR> fe_sqr(const Fe<PR>& a) {
  const uint32_t* v = a.v;
  uint32_t X[14], Y[12];
  // row 0: a0 * a[1..7]
  mul_wide(v[0], v[1], X[0], X[1]);
  mul_wide(v[0], v[3], X[2], X[3]);
  mul_wide(v[0], v[5], X[4], X[5]);
  mul_wide(v[0], v[7], X[6], X[7]);
  mul_wide(v[0], v[2], Y[0], Y[1]);
  mul_wide(v[0], v[4], Y[2], Y[3]);
  mul_wide(v[0], v[6], Y[4], Y[5]);
  // row 1: a1 * a[2..7]   X: columns (3,4),(5,6),(7,8) all written -> carry to column 9
  X[2] = mad_lo_cc(v[1], v[2], X[2]);  X[3] = madc_hi_cc(v[1], v[2], X[3]);
  X[4] = madc_lo_cc(v[1], v[4], X[4]); X[5] = madc_hi_cc(v[1], v[4], X[5]);
  X[6] = madc_lo_cc(v[1], v[6], X[6]); X[7] = madc_hi_cc(v[1], v[6], X[7]);
  X[8] = addc(0, 0);
  //                        Y: columns (4,5),(6,7),(8,9); (8,9) new
  Y[2] = mad_lo_cc(v[1], v[3], Y[2]);  Y[3] = madc_hi_cc(v[1], v[3], Y[3]);
  Y[4] = madc_lo_cc(v[1], v[5], Y[4]); Y[5] = madc_hi_cc(v[1], v[5], Y[5]);
  Y[6] = madc_lo_cc(v[1], v[7], 0);    Y[7] = madc_hi(v[1], v[7], 0);
  // row 2: a2 * a[3..7]   X: (5,6),(7,8),(9,10); column 9 holds a carry bit, 10 is new
  X[4] = mad_lo_cc(v[2], v[3], X[4]);  X[5] = madc_hi_cc(v[2], v[3], X[5]);
  X[6] = madc_lo_cc(v[2], v[5], X[6]); X[7] = madc_hi_cc(v[2], v[5], X[7]);
  X[8] = madc_lo_cc(v[2], v[7], X[8]); X[9] = madc_hi(v[2], v[7], 0);
  //                        Y: (6,7),(8,9) written -> carry to column 10
  Y[4] = mad_lo_cc(v[2], v[4], Y[4]);  Y[5] = madc_hi_cc(v[2], v[4], Y[5]);
  Y[6] = madc_lo_cc(v[2], v[6], Y[6]); Y[7] = madc_hi_cc(v[2], v[6], Y[7]);
  Y[8] = addc(0, 0);
  // row 3: a3 * a[4..7]   X: (7,8),(9,10) written -> carry to column 11
  X[6] = mad_lo_cc(v[3], v[4], X[6]);  X[7] = madc_hi_cc(v[3], v[4], X[7]);
  X[8] = madc_lo_cc(v[3], v[6], X[8]); X[9] = madc_hi_cc(v[3], v[6], X[9]);
  X[10] = addc(0, 0);
  //                        Y: (8,9),(10,11); column 10 holds a carry bit, 11 is new
  Y[6] = mad_lo_cc(v[3], v[5], Y[6]);  Y[7] = madc_hi_cc(v[3], v[5], Y[7]);
  Y[8] = madc_lo_cc(v[3], v[7], Y[8]); Y[9] = madc_hi(v[3], v[7], 0);
  // row 4: a4 * a[5..7]   X: (9,10),(11,12); 12 new
  X[8] = mad_lo_cc(v[4], v[5], X[8]);    X[9] = madc_hi_cc(v[4], v[5], X[9]);
  X[10] = madc_lo_cc(v[4], v[7], X[10]); X[11] = madc_hi(v[4], v[7], 0);
  //                        Y: (10,11) written -> carry to column 12
  Y[8] = mad_lo_cc(v[4], v[6], Y[8]);  Y[9] = madc_hi_cc(v[4], v[6], Y[9]);
  Y[10] = addc(0, 0);
  // row 5: a5 * a[6..7]   X: (11,12) written -> carry to column 13
  X[10] = mad_lo_cc(v[5], v[6], X[10]); X[11] = madc_hi_cc(v[5], v[6], X[11]);
  X[12] = addc(0, 0);
  //                        Y: (12,13); 13 new
  Y[10] = mad_lo_cc(v[5], v[7], Y[10]); Y[11] = madc_hi(v[5], v[7], 0);
  // row 6: a6 * a7        X: (13,14); 14 new
  X[12] = mad_lo_cc(v[6], v[7], X[12]); X[13] = madc_hi(v[6], v[7], 0);

  // T = X * 2^32 + Y * 2^64 (columns 1..15)
  uint32_t T[16];
  T[0] = 0;
  T[1] = X[0];
  T[2] = add_cc(X[1], Y[0]);
#pragma unroll
  for (int k = 3; k < 14; k++) T[k] = addc_cc(X[k - 1], Y[k - 2]);
  T[14] = addc_cc(X[13], 0);
  T[15] = addc(0, 0);
  // S = 2 T + sum_i a_i^2 2^(64 i): the doubling is a funnel shift, the squares ride one carry chain
  uint32_t S[16];
  S[0] = mad_lo_cc(v[0], v[0], 0);
  S[1] = madc_hi_cc(v[0], v[0], detail::shl1(0, T[1]));
#pragma unroll
  for (int i = 1; i < 8; i++) {
    S[2 * i] = madc_lo_cc(v[i], v[i], detail::shl1(T[2 * i - 1], T[2 * i]));
    S[2 * i + 1] = madc_hi_cc(v[i], v[i], detail::shl1(T[2 * i], T[2 * i + 1]));
  }
  return detail::redc_wide<PR>(S);
}

template <class PR>
MZ_HD Fe<PR> fe_to_mont(const Fe<PR>& a) { return fe_mul(a, Fe<PR>::r2()); }

template <class PR>
MZ_HD Fe<PR> fe_from_mont(const Fe<PR>& a) {
  Fe<PR> one_raw = Fe<PR>::zero();
  one_raw.v[0] = 1;
  return fe_mul(a, one_raw);
}

// a^e for a small (<= 64-bit) public exponent; a in Montgomery form
template <class PR>
MZ_HD Fe<PR> fe_pow_u64(const Fe<PR>& a, uint64_t e) {
  Fe<PR> r = Fe<PR>::one();
  Fe<PR> base = a;
  while (e) {
    if (e & 1) r = fe_mul(r, base);
    base = fe_sqr(base);
    e >>= 1;
  }
  return r;
}

// Fermat inverse a^(m-2); inverse(0) = 0, as the reference's ext-Euclid
// returns 0 for 0 (field.rs:210-237 with utils.rs:52-81).
template <class PR>
MZ_HD Fe<PR> fe_inv(const Fe<PR>& a) {
  // exponent m - 2, scanned MSB first
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = mod_limb<PR>(i);
  e[0] -= 2;  // low limb of both moduli is >= 2
  Fe<PR> r = Fe<PR>::one();
  for (int i = 7; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      r = fe_sqr(r);
      if ((e[i] >> bit) & 1) r = fe_mul(r, a);
    }
  }
  return r;
}

// Latency-oriented inverse for the one-off final to-affine of a commitment: binary
// extended GCD on the raw residue (variable time - the data is public), ~5x fewer
// dependent cycles for a lone thread than the 380-multiply Fermat ladder above.
// Input/output in Montgomery form; inverse(0) = 0.
namespace detail {
template <class PR>
MZ_HD bool limbs_is_one(const uint32_t* a) {
  uint32_t o = a[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < 8; i++) o |= a[i];
  return o == 0;
}
MZ_HD void limbs_shr1(uint32_t* a, uint32_t top) {  // a = (top:a) >> 1
#pragma unroll
  for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
  a[7] = (a[7] >> 1) | (top << 31);
}
// x = x/2 mod m for x in [0, m)
template <class PR>
MZ_HD void limbs_half_mod(uint32_t* x) {
  const uint32_t odd = 0u - (x[0] & 1u);
  x[0] = add_cc(x[0], mod_limb<PR>(0) & odd);
#pragma unroll
  for (int i = 1; i < 8; i++) x[i] = addc_cc(x[i], mod_limb<PR>(i) & odd);
  const uint32_t carry = addc(0, 0);
  limbs_shr1(x, carry);
}
MZ_HD bool limbs_geq(const uint32_t* a, const uint32_t* b) {  // a >= b
  (void)sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) (void)subc_cc(a[i], b[i]);
  return subc(0, 0) == 0;
}
MZ_HD void limbs_sub(uint32_t* a, const uint32_t* b) {  // a -= b (a >= b)
  a[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) a[i] = subc_cc(a[i], b[i]);
  a[7] = subc(a[7], b[7]);
}
}  // namespace detail

template <class PR>
MZ_HD Fe<PR> fe_inv_bingcd(const Fe<PR>& a) {
  if (a.is_zero()) return a;
  uint32_t u[8], v[8];
  Fe<PR> x1, x2;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = a.v[i];
    v[i] = mod_limb<PR>(i);
    x1.v[i] = 0;
    x2.v[i] = 0;
  }
  x1.v[0] = 1;
  // invariants: x1 * a == u, x2 * a == v (mod m)
  while (!detail::limbs_is_one<PR>(u) && !detail::limbs_is_one<PR>(v)) {
    while ((u[0] & 1u) == 0) {
      detail::limbs_shr1(u, 0);
      detail::limbs_half_mod<PR>(x1.v);
    }
    while ((v[0] & 1u) == 0) {
      detail::limbs_shr1(v, 0);
      detail::limbs_half_mod<PR>(x2.v);
    }
    if (detail::limbs_geq(u, v)) {
      detail::limbs_sub(u, v);
      x1 = fe_sub(x1, x2);
    } else {
      detail::limbs_sub(v, u);
      x2 = fe_sub(x2, x1);
    }
  }
  Fe<PR> r = detail::limbs_is_one<PR>(u) ? x1 : x2;  // r = (aR)^-1 = a^-1 R^-1
  return fe_mul(fe_mul(r, Fe<PR>::r2()), Fe<PR>::r2());  // -> a^-1 R
}

typedef Fe<FqParams> Fq;
typedef Fe<FrParams> Fr;

}  // namespace mz

// Batched-affine accumulation rounds (BAA) in front of the XYZZ bucket accumulate.
//
// An affine addition costs one field inversion plus 2M + 1S; with Montgomery's trick the
// inversions of a whole batch collapse into one, leaving ~6 multiplies per addition
// instead of the 10 of an XYZZ mixed add.  Every accumulate thread owns a fixed segment
// of the sorted entry list; in a round it pairs up neighbouring items of the same
// bucket (all pairs of a round are independent), so a round is
//     forward  : per thread, running product of the pair denominators (prefixes kept)
//     invert   : one batched inversion over the per-thread products (separate kernel)
//     backward : per thread, finish every pair with its denominator inverse
// Each round halves the runs; after a few rounds the leftover list is finished by the
// serial XYZZ pass.
//
// STATUS (round 1, measured on B200, profiles/baa_r1.md): bit-exact, but only break-even
// with the plain XYZZ accumulate (round 0 is HBM-bound: two passes x two 128-byte random
// gathers per addition, plus a separate inversion kernel per round), so it is OFF by
// default (myzkp_ctx_set_baa_rounds).  A fused per-thread variant with in-thread
// branch-free inversions was measured 1.5x slower and removed.  Group-law special cases are kept (curve.rs:131-145): infinity
// operands, P + P (the doubling slope 3x^2 / 2y goes through the same batch) and
// P + (-P).
//
// The per-thread bodies below are host/device so tests/emul runs them on the CPU.
#pragma once
#include <stddef.h>

#include "g1.cuh"

namespace mz {

// item source of one thread in one round
// Item i of a thread lives at base[i * stride]: the kernels interleave the threads
// (stride = number of threads) so that a warp walking its lists in step touches
// consecutive addresses; the CPU emulation uses stride 1.
struct BaaSrc {
  // round 0: (transposed) sorted entries + resident table
  const uint32_t* keys_s;   // thread's first sorted key
  const uint32_t* vals_s;   // thread's first sorted val (sign << 31 | table index)
  const Affine* tbl;
  // rounds >= 1: the thread's private list
  const Affine* pts;
  const uint32_t* keys;
  size_t stride;
};

template <bool R0>
MZ_HD uint32_t baa_key(const BaaSrc& s, uint32_t i) { return R0 ? s.keys_s[i * s.stride] : s.keys[i * s.stride]; }

MZ_HD Fq baa_load_fq(const Fq* p) {
  Fq r;
#if defined(__CUDA_ARCH__)
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
#else
  r = *p;
#endif
  return r;
}
MZ_HD void baa_store_fq(Fq* p, const Fq& r) {
#if defined(__CUDA_ARCH__)
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
#else
  *p = r;
#endif
}

template <bool R0>
MZ_HD Fq baa_x(const BaaSrc& s, uint32_t i) {
  return baa_load_fq(R0 ? &s.tbl[s.vals_s[i * s.stride] & 0x7fffffffu].x : &s.pts[i * s.stride].x);
}
template <bool R0>
MZ_HD Fq baa_y(const BaaSrc& s, uint32_t i) {
  if (R0) {
    uint32_t v = s.vals_s[i * s.stride];
    Fq y = baa_load_fq(&s.tbl[v & 0x7fffffffu].y);
    return (v >> 31) ? fe_neg(y) : y;
  }
  return baa_load_fq(&s.pts[i * s.stride].y);
}

// Denominator of the pair (a, b) = (item i, item i+1):
//   kind 0: generic      d = xb - xa
//   kind 1: doubling     d = 2 ya          (xa == xb, ya == yb)
//   kind 2: result is b  (a is infinity)   d = 1
//   kind 3: result is a  (b is infinity)   d = 1
//   kind 4: result is infinity (a == -b)   d = 1
// (0, 0) encodes infinity; no curve point has x == 0 (3 is a non-residue mod p).
template <bool R0>
MZ_HD int baa_classify(const BaaSrc& s, uint32_t i, const Fq& xa, const Fq& xb, Fq& d) {
  if (xa.is_zero() && baa_y<R0>(s, i).is_zero()) { d = Fq::one(); return 2; }
  if (xb.is_zero() && baa_y<R0>(s, i + 1).is_zero()) { d = Fq::one(); return 3; }
  d = fe_sub(xb, xa);
  if (!d.is_zero()) return 0;
  Fq ya = baa_y<R0>(s, i), yb = baa_y<R0>(s, i + 1);
  if (ya == yb && !ya.is_zero()) { d = fe_dbl(ya); return 1; }
  d = Fq::one();
  return 4;
}

// number of non-sentinel entries at the head of a round-0 segment of length len
MZ_HD uint32_t baa_count_valid(const uint32_t* keys_s, size_t stride, uint32_t len, uint32_t sentinel) {
  uint32_t n = 0;
  while (n < len && keys_s[n * stride] < sentinel) n++;
  return n;
}

// Pairing: the thread's list is cut into fixed slots (0,1), (2,3), ...; a slot is a pair
// when both items exist and share the bucket key, otherwise its items pass through.  No
// look-ahead is needed and the lanes of a warp stay in step slot by slot, whatever their
// run boundaries are.
//
// Forward pass: slots right to left; prefix[j] = d_0 * ... * d_j in that order.  Returns the
// number of pairs; prod = product of all denominators (one if none).  Prefix element j lives
// at prefix[j * pstride] (the kernels interleave threads so a warp's accesses coalesce).
// Both passes work on a slot range [q0, q1) of the thread's list.
template <bool R0>
MZ_HD uint32_t baa_forward(const BaaSrc& s, uint32_t n, uint32_t q0, uint32_t q1, Fq* prefix, size_t pstride, Fq& prod) {
  uint32_t cnt = 0;
  prod = Fq::one();
  if (q1 > n / 2) q1 = n / 2;  // a trailing single item is never a pair
  for (int64_t q = (int64_t)q1 - 1; q >= (int64_t)q0; q--) {
    const uint32_t i = 2 * (uint32_t)q;
    if (baa_key<R0>(s, i) == baa_key<R0>(s, i + 1)) {
      Fq xa = baa_x<R0>(s, i), xb = baa_x<R0>(s, i + 1);
      Fq d;
      baa_classify<R0>(s, i, xa, xb, d);
      prod = cnt ? fe_mul(prod, d) : d;
      baa_store_fq(prefix + (size_t)cnt * pstride, prod);
      cnt++;
    }
  }
  return cnt;
}

// Backward pass: slots left to right (pairs are met in reverse forward order), writing the
// next list with the source's stride (dst may alias the private source list: o <= i always).
// inv = inverse of prod.  Returns the new item count.
template <bool R0>
MZ_HD uint32_t baa_backward(const BaaSrc& s, uint32_t n, uint32_t q0, uint32_t q1, const Fq* prefix, size_t pstride,
                            Fq inv, uint32_t cnt, Affine* dst_pts, uint32_t* dst_keys, uint32_t o) {
  uint32_t j = cnt;
  for (uint32_t i = 2 * q0; i < n && i < 2 * q1; i += 2) {
    const uint32_t k = baa_key<R0>(s, i);
    const bool has_b = i + 1 < n;
    const uint32_t kb = has_b ? baa_key<R0>(s, i + 1) : 0;
    if (has_b && kb == k) {
      j--;
      Fq xa = baa_x<R0>(s, i), xb = baa_x<R0>(s, i + 1);
      Fq d;
      int kind = baa_classify<R0>(s, i, xa, xb, d);
      Fq dinv = j ? fe_mul(inv, baa_load_fq(prefix + (size_t)(j - 1) * pstride)) : inv;
      inv = fe_mul(inv, d);
      Affine r;
      if (kind <= 1) {
        Fq ya = baa_y<R0>(s, i);
        Fq lam;
        if (kind == 0) {
          lam = fe_mul(fe_sub(baa_y<R0>(s, i + 1), ya), dinv);
        } else {
          Fq xx = fe_sqr(xa);
          lam = fe_mul(fe_add(fe_dbl(xx), xx), dinv);
        }
        r.x = fe_sub(fe_sub(fe_sqr(lam), xa), xb);
        r.y = fe_sub(fe_mul(lam, fe_sub(xa, r.x)), ya);
      } else if (kind == 2) {
        r.x = xb;
        r.y = baa_y<R0>(s, i + 1);
      } else if (kind == 3) {
        r.x = xa;
        r.y = baa_y<R0>(s, i);
      } else {
        r.x = Fq::zero();
        r.y = Fq::zero();
      }
      baa_store_fq(&dst_pts[o * s.stride].x, r.x);
      baa_store_fq(&dst_pts[o * s.stride].y, r.y);
      dst_keys[o * s.stride] = k;
      o++;
    } else {  // pass through (read both before writing: o may equal i)
      Fq ax = baa_x<R0>(s, i), ay = baa_y<R0>(s, i);
      Fq bx = ax, by = ay;
      if (has_b) {
        bx = baa_x<R0>(s, i + 1);
        by = baa_y<R0>(s, i + 1);
      }
      baa_store_fq(&dst_pts[o * s.stride].x, ax);
      baa_store_fq(&dst_pts[o * s.stride].y, ay);
      dst_keys[o * s.stride] = k;
      o++;
      if (has_b) {
        baa_store_fq(&dst_pts[o * s.stride].x, bx);
        baa_store_fq(&dst_pts[o * s.stride].y, by);
        dst_keys[o * s.stride] = kb;
        o++;
      }
    }
  }
  return o;
}

// Finish: serial XYZZ pass over the thread's private list.  Same contract as
// msm_accumulate: the first run goes to *head (key in *head_key), later runs are the
// unique first writers of their buckets.
MZ_HD void baa_finish(const Affine* pts, const uint32_t* keys, size_t stride, uint32_t n, uint32_t sentinel,
                      XYZZ* buckets, XYZZ* head, uint32_t* head_key) {
  if (n == 0) {
    *head_key = sentinel;
    return;
  }
  uint32_t cur = keys[0];
  *head_key = cur;
  XYZZ acc = xyzz_inf();
  bool first_run = true;
  for (uint32_t i = 0; i < n; i++) {
    Affine p;
    p.x = baa_load_fq(&pts[i * stride].x);
    p.y = baa_load_fq(&pts[i * stride].y);
    xyzz_madd(acc, p);
    uint32_t k_next = (i + 1 < n) ? keys[(i + 1) * stride] : sentinel;
    if (k_next != cur) {
      if (first_run) *head = acc;
      else buckets[cur] = acc;
      first_run = false;
      acc = xyzz_inf();
      cur = k_next;
    }
  }
}

}  // namespace mz

// Host run of the batched-affine accumulation rounds (myzkp_b200/csrc/baa.cuh): every
// "thread" of the accumulate grid is executed in turn on the CPU.  TEST INFRASTRUCTURE ONLY.
#include "../../myzkp_b200/csrc/baa.cuh"
#include <string.h>
#include <vector>
using namespace mz;

extern "C" {
// keys/vals: M sorted entries (sentinel keys at the tail); table: Montgomery affine points.
// Runs `rounds` BAA rounds with segments of L entries, the XYZZ finish, merges heads, and
// returns every bucket as a Montgomery affine point (64 B each, (0,0) = infinity).
int emul_baa_buckets(const uint32_t* keys, const uint32_t* vals, uint32_t M, uint32_t L, uint32_t sentinel,
                     const uint32_t* table, int rounds, uint32_t nb, uint32_t* out_affine, int fused) {
  const Affine* tbl = reinterpret_cast<const Affine*>(table);
  const uint32_t T = (M + L - 1) / L;
  std::vector<XYZZ> buckets(nb, xyzz_inf()), heads(T, xyzz_inf());
  std::vector<uint32_t> head_keys(T, sentinel);
  std::vector<Affine> pts((size_t)T * L);
  std::vector<uint32_t> lkeys((size_t)T * L), nitems(T), cnts(T);
  std::vector<Fq> prefix((size_t)T * (L / 2 + 1)), prods(T);
  for (int r = 0; !fused && r < (rounds < 1 ? 1 : rounds); r++) {
    // forward
    for (uint32_t t = 0; t < T; t++) {
      BaaSrc s;
      uint32_t lo = t * L, len = (lo + L <= M) ? L : M - lo;
      s.keys_s = keys + lo; s.vals_s = vals + lo; s.tbl = tbl;
      s.pts = pts.data() + (size_t)t * L; s.keys = lkeys.data() + (size_t)t * L; s.stride = 1;
      if (r == 0) {
        nitems[t] = baa_count_valid(s.keys_s, 1, len, sentinel);
        cnts[t] = rounds >= 1 ? baa_forward<true>(s, nitems[t], 0, L, prefix.data() + (size_t)t * (L / 2 + 1), 1, prods[t]) : 0;
        if (rounds < 1) prods[t] = Fq::one();
      } else {
        cnts[t] = baa_forward<false>(s, nitems[t], 0, L, prefix.data() + (size_t)t * (L / 2 + 1), 1, prods[t]);
      }
    }
    // invert
    for (uint32_t t = 0; t < T; t++) prods[t] = fe_inv(prods[t]);
    // backward
    for (uint32_t t = 0; t < T; t++) {
      BaaSrc s;
      uint32_t lo = t * L;
      s.keys_s = keys + lo; s.vals_s = vals + lo; s.tbl = tbl;
      s.pts = pts.data() + (size_t)t * L; s.keys = lkeys.data() + (size_t)t * L; s.stride = 1;
      Affine* dp = pts.data() + (size_t)t * L;
      uint32_t* dk = lkeys.data() + (size_t)t * L;
      if (r == 0) nitems[t] = baa_backward<true>(s, nitems[t], 0, L, prefix.data() + (size_t)t * (L / 2 + 1), 1, prods[t], cnts[t], dp, dk, 0);
      else nitems[t] = baa_backward<false>(s, nitems[t], 0, L, prefix.data() + (size_t)t * (L / 2 + 1), 1, prods[t], cnts[t], dp, dk, 0);
    }
  }
  for (uint32_t t = 0; t < T; t++)
    baa_finish(pts.data() + (size_t)t * L, lkeys.data() + (size_t)t * L, 1, nitems[t], sentinel, buckets.data(), &heads[t], &head_keys[t]);
  for (uint32_t t = 0; t < T; t++)
    if (head_keys[t] < sentinel) xyzz_add(buckets[head_keys[t]], heads[t]);
  for (uint32_t b = 0; b < nb; b++) {
    Affine a = xyzz_to_affine(buckets[b]);
    memcpy(out_affine + 16 * (size_t)b, &a, 64);
  }
  return 0;
}
}

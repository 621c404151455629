// Kernels of the batched-affine accumulation rounds (see baa.cuh for the algorithm).
#include "baa.cuh"
#include "ctx.cuh"

namespace mz {

constexpr int kBaaThreads = 128;
constexpr int kBaaInvGroup = 16;  // thread products inverted together by one thread of baa_invert

template <bool R0>
__global__ void __launch_bounds__(kBaaThreads) baa_forward_kernel(const uint32_t* __restrict__ keys_s,
                                                                  const uint32_t* __restrict__ vals_s, uint64_t M,
                                                                  uint32_t L, uint32_t sentinel,
                                                                  const Affine* __restrict__ tbl, const Affine* pts,
                                                                  const uint32_t* lkeys, uint32_t* nitems,
                                                                  uint32_t* cnts, Fq* prefix, Fq* prods, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  BaaSrc s;
  s.keys_s = keys_s + t * L;
  s.vals_s = vals_s + t * L;
  s.tbl = tbl;
  s.pts = pts + t * L;
  s.keys = lkeys + t * L;
  uint32_t n;
  if (R0) {
    uint64_t lo = t * L;
    uint32_t len = (uint32_t)(lo + L <= M ? L : M - lo);
    n = baa_count_valid(s.keys_s, len, sentinel);
    nitems[t] = n;
  } else {
    n = nitems[t];
  }
  Fq prod;
  uint32_t cnt = baa_forward<R0>(s, n, prefix + t, (size_t)T, prod);
  cnts[t] = cnt;
  baa_store_fq(prods + t, prod);
}

// in-place inversion of the T thread products: Montgomery's trick over groups of
// kBaaInvGroup, one binary-GCD inversion per group
__global__ void __launch_bounds__(kBaaThreads) baa_invert_kernel(Fq* prods, uint64_t T) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t lo = g * kBaaInvGroup;
  if (lo >= T) return;
  int cnt = (T - lo) < (uint64_t)kBaaInvGroup ? (int)(T - lo) : kBaaInvGroup;
  Fq pref[kBaaInvGroup];
  Fq acc = baa_load_fq(prods + lo);
  pref[0] = acc;
  for (int k = 1; k < cnt; k++) {
    acc = fe_mul(acc, baa_load_fq(prods + lo + k));
    pref[k] = acc;
  }
  Fq inv = fe_inv_bingcd(acc);
  for (int k = cnt - 1; k >= 1; k--) {
    Fq d = baa_load_fq(prods + lo + k);
    baa_store_fq(prods + lo + k, fe_mul(inv, pref[k - 1]));
    inv = fe_mul(inv, d);
  }
  baa_store_fq(prods + lo, inv);
}

template <bool R0>
__global__ void __launch_bounds__(kBaaThreads, 4) baa_backward_kernel(const uint32_t* __restrict__ keys_s,
                                                                      const uint32_t* __restrict__ vals_s, uint32_t L,
                                                                      const Affine* __restrict__ tbl, Affine* pts,
                                                                      uint32_t* lkeys, uint32_t* nitems,
                                                                      const uint32_t* __restrict__ cnts,
                                                                      const Fq* __restrict__ prefix,
                                                                      const Fq* __restrict__ prods, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  BaaSrc s;
  s.keys_s = keys_s + t * L;
  s.vals_s = vals_s + t * L;
  s.tbl = tbl;
  s.pts = pts + t * L;
  s.keys = lkeys + t * L;
  uint32_t n = nitems[t];
  Fq inv = baa_load_fq(prods + t);
  nitems[t] = baa_backward<R0>(s, n, prefix + t, (size_t)T, inv, cnts[t], pts + t * L, lkeys + t * L);
}

__global__ void __launch_bounds__(kBaaThreads, 4) baa_finish_kernel(const Affine* __restrict__ pts,
                                                                    const uint32_t* __restrict__ lkeys, uint32_t L,
                                                                    const uint32_t* __restrict__ nitems,
                                                                    uint32_t sentinel, XYZZ* __restrict__ buckets,
                                                                    XYZZ* __restrict__ heads,
                                                                    uint32_t* __restrict__ head_keys, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  baa_finish(pts + t * L, lkeys + t * L, nitems[t], sentinel, buckets, heads + t, head_keys + t);
}

// Runs `rounds` (>= 1) batched-affine rounds over T segments of L sorted entries and the
// XYZZ finish; same outputs as msm_accumulate (buckets pre-zeroed by the caller).
int baa_accumulate(myzkp_ctx* ctx, const uint32_t* keys_s, const uint32_t* vals_s, uint64_t M, uint32_t L,
                   uint32_t sentinel, int rounds, XYZZ* buckets, XYZZ* heads, uint32_t* head_keys, uint64_t T) {
  const size_t pcap = (size_t)L / 2 + 1;
  MZ_CUDA_TRY(ctx, ctx->baa_pts.ensure((size_t)T * L * sizeof(Affine)));
  MZ_CUDA_TRY(ctx, ctx->baa_keys.ensure((size_t)T * L * sizeof(uint32_t)));
  MZ_CUDA_TRY(ctx, ctx->baa_prefix.ensure((size_t)T * pcap * sizeof(Fq)));
  MZ_CUDA_TRY(ctx, ctx->baa_meta.ensure((size_t)T * (sizeof(Fq) + 2 * sizeof(uint32_t)) + 256));
  Affine* pts = ctx->baa_pts.as<Affine>();
  uint32_t* lkeys = ctx->baa_keys.as<uint32_t>();
  Fq* prefix = ctx->baa_prefix.as<Fq>();
  Fq* prods = ctx->baa_meta.as<Fq>();
  uint32_t* nitems = reinterpret_cast<uint32_t*>(prods + T);
  uint32_t* cnts = nitems + T;
  const unsigned blocks = (unsigned)((T + kBaaThreads - 1) / kBaaThreads);
  const uint64_t groups = (T + kBaaInvGroup - 1) / kBaaInvGroup;
  const unsigned iblocks = (unsigned)((groups + kBaaThreads - 1) / kBaaThreads);
  for (int r = 0; r < rounds; r++) {
    if (r == 0)
      baa_forward_kernel<true><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_s, vals_s, M, L, sentinel, ctx->table, pts,
                                                                        lkeys, nitems, cnts, prefix, prods, T);
    else
      baa_forward_kernel<false><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_s, vals_s, M, L, sentinel, ctx->table,
                                                                         pts, lkeys, nitems, cnts, prefix, prods, T);
    MZ_LAUNCH_CHECK(ctx);
    baa_invert_kernel<<<iblocks, kBaaThreads, 0, ctx->stream>>>(prods, T);
    MZ_LAUNCH_CHECK(ctx);
    if (r == 0)
      baa_backward_kernel<true><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_s, vals_s, L, ctx->table, pts, lkeys,
                                                                         nitems, cnts, prefix, prods, T);
    else
      baa_backward_kernel<false><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_s, vals_s, L, ctx->table, pts, lkeys,
                                                                          nitems, cnts, prefix, prods, T);
    MZ_LAUNCH_CHECK(ctx);
  }
  baa_finish_kernel<<<blocks, kBaaThreads, 0, ctx->stream>>>(pts, lkeys, L, nitems, sentinel, buckets, heads, head_keys, T);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

}  // namespace mz

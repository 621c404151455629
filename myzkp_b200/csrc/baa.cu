// Kernels of the batched-affine accumulation rounds (see baa.cuh for the algorithm).
#include "baa.cuh"
#include "ctx.cuh"

namespace mz {

constexpr int kBaaThreads = 128;

// keys_t[i * T + t] = keys_s[t * L + i] (sentinel beyond M): a T x L -> L x T transpose through
// shared memory so that both sides are coalesced
__global__ void __launch_bounds__(256) baa_transpose_kernel(const uint32_t* __restrict__ keys_s,
                                                            const uint32_t* __restrict__ vals_s, uint64_t M, uint32_t L,
                                                            uint64_t T, uint32_t sentinel, uint32_t* __restrict__ keys_t,
                                                            uint32_t* __restrict__ vals_t) {
  __shared__ uint32_t tk[32][33], tv[32][33];
  const uint64_t t0 = (uint64_t)blockIdx.x * 32;  // 32 threads (rows of the T x L matrix)
  const uint32_t i0 = blockIdx.y * 32;            // 32 entries (columns)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    uint64_t t = t0 + r;
    uint32_t i = i0 + tx;
    uint64_t src = t * L + i;
    bool ok = t < T && i < L && src < M;
    tk[r][tx] = ok ? keys_s[src] : sentinel;
    tv[r][tx] = ok ? vals_s[src] : 0u;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    uint32_t i = i0 + r;
    uint64_t t = t0 + tx;
    if (i < L && t < T) {
      keys_t[(uint64_t)i * T + t] = tk[tx][r];
      vals_t[(uint64_t)i * T + t] = tv[tx][r];
    }
  }
}

constexpr int kBaaInvGroup = 16;  // thread products inverted together by one thread of baa_invert_kernel

template <bool R0>
__global__ void __launch_bounds__(kBaaThreads) baa_forward_kernel(const uint32_t* __restrict__ keys_t,
                                                                  const uint32_t* __restrict__ vals_t, uint32_t L,
                                                                  uint32_t sentinel, const Affine* __restrict__ tbl,
                                                                  const Affine* pts, const uint32_t* lkeys,
                                                                  uint32_t* nitems, uint32_t* cnts, Fq* prefix, Fq* prods,
                                                                  uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  BaaSrc s;
  s.keys_s = keys_t + t;  // transposed: entry i of thread t at [i * T + t]
  s.vals_s = vals_t + t;
  s.tbl = tbl;
  s.pts = pts + t;
  s.keys = lkeys + t;
  s.stride = (size_t)T;
  uint32_t n;
  if (R0) {
    n = baa_count_valid(s.keys_s, s.stride, L, sentinel);  // the transpose pads short segments with sentinels
    nitems[t] = n;
  } else {
    n = nitems[t];
  }
  Fq prod;
  cnts[t] = baa_forward<R0>(s, n, 0, L, prefix + t, (size_t)T, prod);
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
__global__ void __launch_bounds__(kBaaThreads, 4) baa_backward_kernel(const uint32_t* __restrict__ keys_t,
                                                                      const uint32_t* __restrict__ vals_t, uint32_t L,
                                                                      const Affine* __restrict__ tbl, Affine* pts,
                                                                      uint32_t* lkeys, uint32_t* nitems,
                                                                      const uint32_t* __restrict__ cnts,
                                                                      const Fq* __restrict__ prefix,
                                                                      const Fq* __restrict__ prods, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  BaaSrc s;
  s.keys_s = keys_t + t;
  s.vals_s = vals_t + t;
  s.tbl = tbl;
  s.pts = pts + t;
  s.keys = lkeys + t;
  s.stride = (size_t)T;
  Fq inv = baa_load_fq(prods + t);
  nitems[t] = baa_backward<R0>(s, nitems[t], 0, L, prefix + t, (size_t)T, inv, cnts[t], pts + t, lkeys + t, 0);
}

__global__ void __launch_bounds__(kBaaThreads, 4) baa_finish_kernel(const Affine* __restrict__ pts,
                                                                    const uint32_t* __restrict__ lkeys, uint32_t L,
                                                                    const uint32_t* __restrict__ nitems,
                                                                    uint32_t sentinel, XYZZ* __restrict__ buckets,
                                                                    XYZZ* __restrict__ heads,
                                                                    uint32_t* __restrict__ head_keys, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  baa_finish(pts + t, lkeys + t, (size_t)T, nitems[t], sentinel, buckets, heads + t, head_keys + t);
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
  MZ_CUDA_TRY(ctx, ctx->baa_trans.ensure((size_t)T * L * 2 * sizeof(uint32_t)));
  Affine* pts = ctx->baa_pts.as<Affine>();
  uint32_t* lkeys = ctx->baa_keys.as<uint32_t>();
  Fq* prefix = ctx->baa_prefix.as<Fq>();
  Fq* prods = ctx->baa_meta.as<Fq>();
  uint32_t* nitems = reinterpret_cast<uint32_t*>(prods + T);
  uint32_t* cnts = nitems + T;
  uint32_t* keys_t = ctx->baa_trans.as<uint32_t>();
  uint32_t* vals_t = keys_t + (size_t)T * L;
  {
    dim3 grid((unsigned)((T + 31) / 32), (L + 31) / 32);
    baa_transpose_kernel<<<grid, 256, 0, ctx->stream>>>(keys_s, vals_s, M, L, T, sentinel, keys_t, vals_t);
    MZ_LAUNCH_CHECK(ctx);
  }
  const unsigned blocks = (unsigned)((T + kBaaThreads - 1) / kBaaThreads);
  const uint64_t groups = (T + kBaaInvGroup - 1) / kBaaInvGroup;
  const unsigned iblocks = (unsigned)((groups + kBaaThreads - 1) / kBaaThreads);
  for (int r = 0; r < rounds; r++) {
    if (r == 0)
      baa_forward_kernel<true><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_t, vals_t, L, sentinel, ctx->table, pts, lkeys,
                                                                        nitems, cnts, prefix, prods, T);
    else
      baa_forward_kernel<false><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_t, vals_t, L, sentinel, ctx->table, pts,
                                                                         lkeys, nitems, cnts, prefix, prods, T);
    MZ_LAUNCH_CHECK(ctx);
    baa_invert_kernel<<<iblocks, kBaaThreads, 0, ctx->stream>>>(prods, T);
    MZ_LAUNCH_CHECK(ctx);
    if (r == 0)
      baa_backward_kernel<true><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_t, vals_t, L, ctx->table, pts, lkeys, nitems,
                                                                         cnts, prefix, prods, T);
    else
      baa_backward_kernel<false><<<blocks, kBaaThreads, 0, ctx->stream>>>(keys_t, vals_t, L, ctx->table, pts, lkeys, nitems,
                                                                          cnts, prefix, prods, T);
    MZ_LAUNCH_CHECK(ctx);
  }
  baa_finish_kernel<<<blocks, kBaaThreads, 0, ctx->stream>>>(pts, lkeys, L, nitems, sentinel, buckets, heads, head_keys, T);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

}  // namespace mz

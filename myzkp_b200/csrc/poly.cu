// Scalar-field (Fr) polynomial kernels of open_kzg / Gemini fold.
//
// Replaces, for divisor (x - u):
//   Polynomial::eval                polynomial.rs:120-128  (running-power sum)
//   f - y, from_monomials, Div      polynomial.rs:517-523, 202-212, 583-597
//   div_rem_ref                     polynomial.rs:371-405  (O(d^2) in the reference)
// by one suffix scan: c_i = f_i + u * c_{i+1}; then q_{i-1} = c_i and y = c_0.
// and split_and_fold's level step   gemini.rs:71-98: g[k] = f[2k] + rho * f[2k+1].
//
// Coefficients stay canonical (non-Montgomery) in memory; only u / rho are in
// Montgomery form, since montmul(u*R, x) = u*x is again canonical.
#include "ctx.cuh"

namespace mz {

constexpr int kPolyThreads = 256;
constexpr int kPolyPerThread = 8;
constexpr int kPolyTile = kPolyThreads * kPolyPerThread;

struct FrMap {  // x -> h + m * x ; m Montgomery, h canonical
  Fr m, h;
};

__device__ __forceinline__ Fr load_fr(const uint32_t* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void store_fr(uint32_t* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// compose: apply `hi` (higher indices) first, then `lo`
__device__ __forceinline__ FrMap compose(const FrMap& lo, const FrMap& hi) {
  FrMap r;
  r.m = fe_mul(lo.m, hi.m);
  r.h = fe_add(lo.h, fe_mul(lo.m, hi.h));
  return r;
}

// upw[k] = u^k (Montgomery) for k = 0..kPolyPerThread
__global__ void poly_small_powers(const uint32_t* u_canon, Fr* upw) {
  Fr u = fe_to_mont(load_fr(u_canon));
  Fr p = Fr::one();
  for (int k = 0; k <= kPolyPerThread; k++) {
    upw[k] = p;
    p = fe_mul(p, u);
  }
}

// per-thread map of its <= 8 coefficients, then an in-block inclusive suffix
// scan (Kogge-Stone) of the maps.  On return sm[t] = M_t o M_{t+1} o ... o M_255.
__device__ __forceinline__ void tile_suffix_scan(const uint32_t* coefs, size_t n, size_t tile,
                                                 const Fr* upw, FrMap* sm, int& len_out, size_t& lo_out) {
  const int t = threadIdx.x;
  size_t lo = tile * kPolyTile + (size_t)t * kPolyPerThread;
  int len = 0;
  if (lo < n) len = (n - lo) < (size_t)kPolyPerThread ? (int)(n - lo) : kPolyPerThread;
  const Fr u = upw[1];
  FrMap me;
  me.h = Fr::zero();
  for (int k = len - 1; k >= 0; k--) me.h = fe_add(load_fr(coefs + (lo + k) * 8), fe_mul(u, me.h));
  me.m = upw[len];
  sm[t] = me;
  __syncthreads();
#pragma unroll 1
  for (int d = 1; d < kPolyThreads; d <<= 1) {
    FrMap nw = me;
    if (t + d < kPolyThreads) nw = compose(me, sm[t + d]);
    __syncthreads();
    me = nw;
    sm[t] = me;
    __syncthreads();
  }
  len_out = len;
  lo_out = lo;
}

__global__ void __launch_bounds__(kPolyThreads) poly_tile_maps(const uint32_t* coefs, size_t n, const Fr* upw,
                                                               FrMap* tiles) {
  __shared__ FrMap sm[kPolyThreads];
  int len;
  size_t lo;
  tile_suffix_scan(coefs, n, blockIdx.x, upw, sm, len, lo);
  if (threadIdx.x == 0) tiles[blockIdx.x] = sm[0];
}

// single block: carry entering every tile from above, and the whole-range map
__global__ void __launch_bounds__(kPolyThreads) poly_tiles_scan(const FrMap* tiles, size_t ntiles,
                                                                const uint32_t* carry_in, uint32_t* tile_carry,
                                                                uint32_t* out_h, uint32_t* out_upow) {
  __shared__ FrMap sm[kPolyThreads];
  const int t = threadIdx.x;
  size_t per = (ntiles + kPolyThreads - 1) / kPolyThreads;
  size_t first = (size_t)t * per;
  size_t last = first + per < ntiles ? first + per : ntiles;  // exclusive
  FrMap me;
  me.m = Fr::one();
  me.h = Fr::zero();
  for (size_t b = last; b > first; b--) me = compose(tiles[b - 1], me);
  sm[t] = me;
  __syncthreads();
#pragma unroll 1
  for (int d = 1; d < kPolyThreads; d <<= 1) {
    FrMap nw = me;
    if (t + d < kPolyThreads) nw = compose(me, sm[t + d]);
    __syncthreads();
    me = nw;
    sm[t] = me;
    __syncthreads();
  }
  Fr cin = load_fr(carry_in);
  if (tile_carry) {
    Fr x = cin;
    if (t + 1 < kPolyThreads) {
      FrMap s = sm[t + 1];
      x = fe_add(s.h, fe_mul(s.m, cin));
    }
    for (size_t b = last; b > first; b--) {
      store_fr(tile_carry + (b - 1) * 8, x);
      FrMap mb = tiles[b - 1];
      x = fe_add(mb.h, fe_mul(mb.m, x));
    }
  }
  if (t == 0) {
    FrMap s = sm[0];
    if (out_h) store_fr(out_h, fe_add(s.h, fe_mul(s.m, cin)));
    if (out_upow) store_fr(out_upow, fe_from_mont(s.m));
  }
}

// q[i] = c_{i+1} for the range (q[n-1] = carry entering the range), c0 = c_0
__global__ void __launch_bounds__(kPolyThreads) poly_tile_quotient(const uint32_t* coefs, size_t n, const Fr* upw,
                                                                   const uint32_t* tile_carry, uint32_t* q,
                                                                   uint32_t* c0) {
  __shared__ FrMap sm[kPolyThreads];
  int len;
  size_t lo;
  tile_suffix_scan(coefs, n, blockIdx.x, upw, sm, len, lo);
  if (len == 0) return;
  const int t = threadIdx.x;
  Fr x = load_fr(tile_carry + (size_t)blockIdx.x * 8);
  if (t + 1 < kPolyThreads) {
    FrMap s = sm[t + 1];
    x = fe_add(s.h, fe_mul(s.m, x));
  }
  const Fr u = upw[1];
  // x = c_{lo+len}: belongs to q[lo+len-1]
  for (int k = len - 1; k >= 0; k--) {
    size_t i = lo + k;
    store_fr(q + i * 8, x);  // q[i] = c_{i+1}
    x = fe_add(load_fr(coefs + i * 8), fe_mul(u, x));
  }
  if (lo == 0) store_fr(c0, x);
}

__global__ void poly_fold(const uint32_t* in, size_t n_out, const uint32_t* rho_canon, uint32_t* out) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_out) return;
  Fr rho = fe_to_mont(load_fr(rho_canon));
  Fr e = load_fr(in + (2 * k) * 8);
  Fr o = load_fr(in + (2 * k + 1) * 8);
  store_fr(out + k * 8, fe_add(e, fe_mul(rho, o)));
}

__global__ void poly_check_canonical(const uint32_t* in, size_t n, int* flag) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (!fe_is_canonical(load_fr(in + k * 8))) atomicOr(flag, 1);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// small-buffer layout (ctx->small, 4 KiB): see also msm.cu
//   [0,32)    u / rho staged        [32,64)  carry staged
//   [64,...)  upw[0..8] (9 * 32 B)  [512,..) misc outputs
static int stage_small(myzkp_ctx* ctx, const uint8_t u_le[32], const uint8_t carry_le[32],
                       const uint32_t* d_carry = nullptr) {
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  uint8_t* s = ctx->small.as<uint8_t>();
  // pageable 32-byte sources: cudaMemcpyAsync stages them before returning
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s, u_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  if (d_carry) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s + 32, d_carry, 32, cudaMemcpyDeviceToDevice, ctx->stream));
  else if (carry_le) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s + 32, carry_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  else MZ_CUDA_TRY(ctx, cudaMemsetAsync(s + 32, 0, 32, ctx->stream));
  poly_small_powers<<<1, 1, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(s), reinterpret_cast<Fr*>(s + 64));
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int fr_range_eval(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32], uint32_t* d_h,
                  uint32_t* d_upow) {
  MZ_TRY(stage_small(ctx, u_le, nullptr));
  uint8_t* s = ctx->small.as<uint8_t>();
  size_t ntiles = (n + kPolyTile - 1) / kPolyTile;
  if (ntiles == 0) ntiles = 1;
  MZ_CUDA_TRY(ctx, ctx->poly_tiles.ensure(ntiles * (sizeof(FrMap) + 32)));
  FrMap* tiles = ctx->poly_tiles.as<FrMap>();
  poly_tile_maps<<<(unsigned)ntiles, kPolyThreads, 0, ctx->stream>>>(d_coefs, n, reinterpret_cast<Fr*>(s + 64), tiles);
  MZ_LAUNCH_CHECK(ctx);
  poly_tiles_scan<<<1, kPolyThreads, 0, ctx->stream>>>(tiles, ntiles, reinterpret_cast<uint32_t*>(s + 32), nullptr,
                                                       d_h, d_upow);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int fr_range_quotient(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32],
                      const uint8_t carry_le[32], uint32_t* d_q, uint32_t* d_c0, const uint32_t* d_carry) {
  if (n == 0) return MYZKP_OK;
  MZ_TRY(stage_small(ctx, u_le, carry_le, d_carry));
  uint8_t* s = ctx->small.as<uint8_t>();
  size_t ntiles = (n + kPolyTile - 1) / kPolyTile;
  MZ_CUDA_TRY(ctx, ctx->poly_tiles.ensure(ntiles * (sizeof(FrMap) + 32)));
  FrMap* tiles = ctx->poly_tiles.as<FrMap>();
  uint32_t* tile_carry = reinterpret_cast<uint32_t*>(tiles + ntiles);
  const Fr* upw = reinterpret_cast<Fr*>(s + 64);
  poly_tile_maps<<<(unsigned)ntiles, kPolyThreads, 0, ctx->stream>>>(d_coefs, n, upw, tiles);
  MZ_LAUNCH_CHECK(ctx);
  poly_tiles_scan<<<1, kPolyThreads, 0, ctx->stream>>>(tiles, ntiles, reinterpret_cast<uint32_t*>(s + 32),
                                                       tile_carry, nullptr, nullptr);
  MZ_LAUNCH_CHECK(ctx);
  poly_tile_quotient<<<(unsigned)ntiles, kPolyThreads, 0, ctx->stream>>>(d_coefs, n, upw, tile_carry, d_q, d_c0);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int fr_fold(myzkp_ctx* ctx, const uint32_t* d_in, size_t n_out, const uint32_t* d_rho, uint32_t* d_out) {
  if (n_out == 0) return MYZKP_OK;
  poly_fold<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(d_in, n_out, d_rho, d_out);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int fr_check_canonical(myzkp_ctx* ctx, const uint32_t* d_in, size_t n, int* d_flag) {
  if (n == 0) return MYZKP_OK;
  poly_check_canonical<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_in, n, d_flag);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

}  // namespace mz

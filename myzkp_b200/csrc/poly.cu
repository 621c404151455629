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
constexpr int kPolyPerThread = 16;
constexpr int kPolyTile = kPolyThreads * kPolyPerThread;  // 4096 coefficients
constexpr int kPolyBlockLevels = 8;                        // log2(kPolyThreads)
constexpr int kPolyTileLevels = 8;                         // log2(threads of the tile-carry scan)

// Powers of u the scan needs (Montgomery form):
//   [0]                u
//   [1 + l], l < 8     u^(K * 2^l)      in-block Kogge-Stone strides (K = kPolyPerThread)
//   [9]                u^TILE
//   [10]               u^(TILE * per)   per = tiles handled serially by one thread of poly_tiles_scan
//   [11 + l], l < 8    u^(TILE * per * 2^l)
constexpr int kPolyPowCount = 11 + kPolyTileLevels;

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

__device__ Fr fr_pow_u64(Fr base, uint64_t e) {
  Fr r = Fr::one();
  while (e) {
    if (e & 1) r = fe_mul(r, base);
    base = fe_sqr(base);
    e >>= 1;
  }
  return r;
}

__global__ void poly_powers(const uint32_t* u_canon, uint64_t per, Fr* pw) {
  Fr u = fe_to_mont(load_fr(u_canon));
  pw[0] = u;
  Fr p = fr_pow_u64(u, kPolyPerThread);
  for (int l = 0; l < kPolyBlockLevels; l++) {
    pw[1 + l] = p;
    p = fe_sqr(p);
  }
  pw[9] = p;  // u^(K * 256) = u^TILE
  p = fr_pow_u64(p, per);
  pw[10] = p;
  for (int l = 0; l < kPolyTileLevels; l++) {
    pw[11 + l] = p;
    p = fe_sqr(p);
  }
}

// tpw[t] = (u^K)^t for t < kPolyThreads (Montgomery): weight of the tile carry at thread t
__global__ void poly_thread_powers(const Fr* pw, Fr* tpw) { tpw[threadIdx.x] = fr_pow_u64(pw[1], (uint64_t)threadIdx.x); }

// The scan runs from the top coefficient down: position p = n - 1 - i, c(p) = f(p) + u c(p - 1),
// c(-1) = carry entering the range.  Every thread owns K consecutive positions, so every
// thread's map is x -> h_t + u^K x with the SAME multiplier: composing maps only needs the
// h parts and a table of u^(K 2^l).  Positions past the bottom (i < 0) read as zero.
__device__ __forceinline__ Fr poly_coef(const uint32_t* coefs, size_t n, size_t p) {
  return p < n ? load_fr(coefs + (n - 1 - p) * 8) : Fr::zero();
}

// inclusive in-block scan of the per-thread h values: on return h = carry leaving thread t
// when the carry entering the tile is zero.  sm: kPolyThreads field elements.
__device__ __forceinline__ Fr poly_block_scan(Fr h, const Fr* pw, Fr* sm) {
  const int t = threadIdx.x;
  sm[t] = h;
  __syncthreads();
#pragma unroll 1
  for (int l = 0; l < kPolyBlockLevels; l++) {
    const int d = 1 << l;
    Fr nw = h;
    if (t >= d) nw = fe_add(h, fe_mul(pw[1 + l], sm[t - d]));
    __syncthreads();
    h = nw;
    sm[t] = h;
    __syncthreads();
  }
  return h;
}

__device__ __forceinline__ Fr poly_thread_h(const uint32_t* coefs, size_t n, size_t p0, const Fr& u) {
  Fr h = poly_coef(coefs, n, p0);
#pragma unroll 4
  for (int k = 1; k < kPolyPerThread; k++) h = fe_add(poly_coef(coefs, n, p0 + k), fe_mul(u, h));
  return h;
}

// tile_h[b] = carry leaving tile b when the carry entering it is zero
// (and thread_incl[b * 256 + t] = carry leaving thread t of tile b under the same condition)
__global__ void __launch_bounds__(kPolyThreads) poly_tile_maps(const uint32_t* coefs, size_t n, const Fr* pw, Fr* tile_h,
                                                               Fr* thread_incl) {
  __shared__ Fr sm[kPolyThreads];
  const size_t p0 = ((size_t)blockIdx.x * kPolyThreads + threadIdx.x) * kPolyPerThread;
  Fr h = poly_block_scan(poly_thread_h(coefs, n, p0, pw[0]), pw, sm);
  thread_incl[(size_t)blockIdx.x * kPolyThreads + threadIdx.x] = h;
  if (threadIdx.x == kPolyThreads - 1) tile_h[blockIdx.x] = h;
}

// single block: carry entering every tile (tile_carry, canonical); out_h = carry leaving the last
// tile = c at the bottom of the padded range
__global__ void __launch_bounds__(kPolyThreads) poly_tiles_scan(const Fr* tile_h, size_t ntiles, uint64_t per,
                                                                const uint32_t* carry_in, const Fr* pw,
                                                                uint32_t* tile_carry, uint32_t* out_last) {
  __shared__ Fr sm[kPolyThreads];
  const int t = threadIdx.x;
  const size_t first = (size_t)t * per;
  const size_t last = first + per < ntiles ? first + per : ntiles;
  const Fr mt = pw[9];  // u^TILE
  // this thread's `per` tiles as one map x -> h + (u^TILE)^per x (missing tiles act as zero tiles)
  Fr h = Fr::zero();
  const Fr cin = load_fr(carry_in);
  if (t == 0) h = cin;  // the carry entering the range rides through the scan as thread 0's starting value
  for (size_t b = first; b < first + per; b++) h = fe_add(b < ntiles ? tile_h[b] : Fr::zero(), fe_mul(mt, h));
  // inclusive scan over threads with stride powers pw[11 + l]
  sm[t] = h;
  __syncthreads();
#pragma unroll 1
  for (int l = 0; l < kPolyTileLevels; l++) {
    const int d = 1 << l;
    Fr nw = h;
    if (t >= d) nw = fe_add(h, fe_mul(pw[11 + l], sm[t - d]));
    __syncthreads();
    h = nw;
    sm[t] = h;
    __syncthreads();
  }
  // carry entering thread t's first tile = carry leaving thread t-1 (cin already folded in)
  Fr x = t > 0 ? sm[t - 1] : cin;
  for (size_t b = first; b < last; b++) {
    if (tile_carry) store_fr(tile_carry + b * 8, x);
    x = fe_add(tile_h[b], fe_mul(mt, x));
  }
  if (out_last && last == ntiles && first < ntiles) store_fr(out_last, x);
}

// q[i] = c_{i+1} for the range (q[n-1] = carry entering the range), c0 = c_0
__global__ void __launch_bounds__(kPolyThreads) poly_tile_quotient(const uint32_t* coefs, size_t n, const Fr* pw,
                                                                   const Fr* tpw, const Fr* thread_incl,
                                                                   const uint32_t* tile_carry, uint32_t* q, uint32_t* c0) {
  const int t = threadIdx.x;
  const size_t p0 = ((size_t)blockIdx.x * kPolyThreads + t) * kPolyPerThread;
  if (p0 >= n) return;
  const Fr u = pw[0];
  // carry entering this thread = (u^K)^t * tile carry + carry leaving thread t-1 under a zero tile carry
  Fr x = load_fr(tile_carry + (size_t)blockIdx.x * 8);
  if (t > 0) x = fe_add(thread_incl[(size_t)blockIdx.x * kPolyThreads + t - 1], fe_mul(tpw[t], x));
#pragma unroll 4
  for (int k = 0; k < kPolyPerThread; k++) {
    const size_t p = p0 + k;
    if (p >= n) break;
    const size_t i = n - 1 - p;
    store_fr(q + i * 8, x);  // x = c_{i+1}
    x = fe_add(load_fr(coefs + i * 8), fe_mul(u, x));
    if (i == 0) store_fr(c0, x);
  }
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
// small-buffer layout (ctx->small, 4 KiB): see also capi.cu (flag 512, y 640, XYZZ 1024.., point 3072)
//   [0,32)    u staged     [32,64)  carry staged     [3200, 3200 + 19*32)  powers of u
constexpr size_t kPolyPowOffset = 3200;
static_assert(kPolyPowOffset + kPolyPowCount * sizeof(Fr) <= 4096, "powers must fit the small buffer");
static int stage_small(myzkp_ctx* ctx, const uint8_t u_le[32], const uint8_t carry_le[32], const uint32_t* d_carry,
                       uint64_t per) {
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  uint8_t* s = ctx->small.as<uint8_t>();
  // pageable 32-byte sources: cudaMemcpyAsync stages them before returning
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s, u_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  if (d_carry) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s + 32, d_carry, 32, cudaMemcpyDeviceToDevice, ctx->stream));
  else if (carry_le) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s + 32, carry_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  else MZ_CUDA_TRY(ctx, cudaMemsetAsync(s + 32, 0, 32, ctx->stream));
  poly_powers<<<1, 1, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(s), per, reinterpret_cast<Fr*>(s + kPolyPowOffset));
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

// u^n (canonical) by square-and-multiply on one thread
__global__ void poly_upow(const Fr* pw, uint64_t n, uint32_t* out) { store_fr(out, fe_from_mont(fr_pow_u64(pw[0], n))); }

static int run_scan(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32], const uint8_t carry_le[32],
                    const uint32_t* d_carry, uint32_t* d_q, uint32_t* d_c0, uint32_t* d_h) {
  const size_t ntiles = (n + kPolyTile - 1) / kPolyTile;
  const uint64_t per = (ntiles + kPolyThreads - 1) / kPolyThreads;
  MZ_TRY(stage_small(ctx, u_le, carry_le, d_carry, per));
  uint8_t* s = ctx->small.as<uint8_t>();
  const Fr* pw = reinterpret_cast<Fr*>(s + kPolyPowOffset);
  // scratch: tile_h[ntiles] | tile_carry[ntiles] | tpw[256] | thread_incl[ntiles * 256]
  MZ_CUDA_TRY(ctx, ctx->poly_tiles.ensure((2 * ntiles + kPolyThreads + ntiles * kPolyThreads) * sizeof(Fr) + 64));
  Fr* tile_h = ctx->poly_tiles.as<Fr>();
  uint32_t* tile_carry = reinterpret_cast<uint32_t*>(tile_h + ntiles);
  Fr* tpw = tile_h + 2 * ntiles;
  Fr* thread_incl = tpw + kPolyThreads;
  poly_thread_powers<<<1, kPolyThreads, 0, ctx->stream>>>(pw, tpw);
  MZ_LAUNCH_CHECK(ctx);
  poly_tile_maps<<<(unsigned)ntiles, kPolyThreads, 0, ctx->stream>>>(d_coefs, n, pw, tile_h, thread_incl);
  MZ_LAUNCH_CHECK(ctx);
  // without d_q only the value at the bottom is wanted; with padding below index 0 that is NOT the
  // carry leaving the last tile, so the quotient kernel (which knows where i == 0 is) is always used
  poly_tiles_scan<<<1, kPolyThreads, 0, ctx->stream>>>(tile_h, ntiles, per, reinterpret_cast<uint32_t*>(s + 32), pw,
                                                       tile_carry, nullptr);
  MZ_LAUNCH_CHECK(ctx);
  (void)d_h;
  poly_tile_quotient<<<(unsigned)ntiles, kPolyThreads, 0, ctx->stream>>>(d_coefs, n, pw, tpw, thread_incl, tile_carry,
                                                                         d_q, d_c0);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int fr_range_eval(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32], uint32_t* d_h,
                  uint32_t* d_upow) {
  if (n == 0) {
    MZ_TRY(stage_small(ctx, u_le, nullptr, nullptr, 1));
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(d_h, 0, 32, ctx->stream));
  } else {
    // the evaluation is c_0 of the scan; the quotient coefficients go to scratch
    MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 32));
    MZ_TRY(run_scan(ctx, d_coefs, n, u_le, nullptr, nullptr, ctx->scalars2.as<uint32_t>(), d_h, nullptr));
  }
  poly_upow<<<1, 1, 0, ctx->stream>>>(reinterpret_cast<Fr*>(ctx->small.as<uint8_t>() + kPolyPowOffset), (uint64_t)n, d_upow);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int fr_range_quotient(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32],
                      const uint8_t carry_le[32], uint32_t* d_q, uint32_t* d_c0, const uint32_t* d_carry) {
  if (n == 0) return MYZKP_OK;
  return run_scan(ctx, d_coefs, n, u_le, carry_le, d_carry, d_q, d_c0, nullptr);
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

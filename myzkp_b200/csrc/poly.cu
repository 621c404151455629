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

// ---------------------------------------------------------------------------
// single-pass suffix scan with decoupled look-back
// ---------------------------------------------------------------------------
// Position p' counts from the TOP of the (zero-padded) range down: the range is padded with zero
// coefficients ABOVE its top coefficient up to a whole number of tiles (leading zeros change
// nothing), so the last tile ends exactly at index 0 and its carry-out is c_0.  With
// c(p') = f(p') + u c(p' - 1) the quotient coefficient at index i is the carry ENTERING i's
// position.  The carry entering the range (sharded open: the value handed down by the range above)
// is injected as the virtual coefficient just above the top one (or as tile 0's carry when there is
// no padding).
//
// One block owns one tile of kScanTile coefficients (ticket order, so a tile only ever waits for
// tiles that are already running):
//   load      coalesced 16-byte loads into shared memory, one padded chunk per thread
//   thread    zero-carry Horner over the thread's K coefficients (two half chains)  K-1 multiplies
//   warp      Kogge-Stone over lanes by shuffles, multipliers (u^K)^(2^l)  5 multiplies
//   block     scan of the warp aggregates; the tile's zero-carry aggregate is published
//   look-back the block reads 128 predecessor tiles at a time (aggregate or inclusive carry),
//             weights them by (u^TILE)^l and sums them; publishes the inclusive carry
//   thread    true carry = exclusive scan value + (u^K)^lane * warp carry    1 multiply
//   output    second Horner pass with the true carry (two half chains), in place  K+1 multiplies
//   store     coalesced copy-out of the tile's quotient coefficients
// HBM traffic: 32 B read + 32 B written per coefficient (the algorithmic 64 B); 2.3 Fr multiplies
// per coefficient.  Without d_q (evaluation only) nothing but the 32-byte result is written.
constexpr int kScanThr = 128;
constexpr int kScanK = 16;
constexpr int kScanTile = kScanThr * kScanK;  // 2048 coefficients
constexpr int kScanWarps = kScanThr / 32;
constexpr int kScanChunkBytes = kScanK * 32 + 16;  // +16: consecutive threads start in different 16-byte bank groups
constexpr int kScanSmemBytes = kScanThr * kScanChunkBytes;

// Montgomery-form powers of u used by the scan
struct ScanPow {
  Fr u;
  Fr half;                // u^(K/2)
  Fr lane[32];            // (u^K)^l
  Fr warp[kScanWarps];    // (u^(32 K))^w
  Fr tile[kScanThr + 1];  // (u^TILE)^l
};
// per-call look-back state: flags[ntiles] (0 empty, 1 aggregate, 2 inclusive), ticket, then the values
struct ScanState {
  uint32_t* flags;
  uint32_t* ticket;
  Fr* agg;
  Fr* incl;
};

__device__ Fr fr_pow_u64(Fr base, uint64_t e) {
  Fr r = Fr::one();
  while (e) {
    if (e & 1) r = fe_mul(r, base);
    base = fe_sqr(base);
    e >>= 1;
  }
  return r;
}

// one thread per table entry; also u^n (canonical) for the range evaluations
__global__ void poly_scan_powers(const uint32_t* u_canon, ScanPow* pw, uint64_t n, uint32_t* d_upow) {
  const Fr u = fe_to_mont(load_fr(u_canon));
  const int t = threadIdx.x;
  if (t == 0) pw->u = u;
  if (t == 1) pw->half = fr_pow_u64(u, kScanK / 2);
  if (t < 32) pw->lane[t] = fr_pow_u64(u, (uint64_t)kScanK * t);
  else if (t < 32 + kScanWarps) pw->warp[t - 32] = fr_pow_u64(u, (uint64_t)32 * kScanK * (t - 32));
  else if (t < 32 + kScanWarps + kScanThr + 1) pw->tile[t - 32 - kScanWarps] = fr_pow_u64(u, (uint64_t)kScanTile * (t - 32 - kScanWarps));
  else if (t == 32 + kScanWarps + kScanThr + 1 && d_upow) store_fr(d_upow, fe_from_mont(fr_pow_u64(u, n)));
}
constexpr int kScanPowThreads = 32 + kScanWarps + kScanThr + 1 + 1;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ Fr shfl_up_fr(const Fr& a, int d) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], d);
  return r;
}
__device__ __forceinline__ Fr shfl_xor_fr(const Fr& a, int m) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], m);
  return r;
}
__device__ __forceinline__ Fr shfl_fr(const Fr& a, int src) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}
__device__ __forceinline__ Fr ldcg_fr(const Fr* p) {  // L2-coherent load of a value another block published
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldcg(q), b = __ldcg(q + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Fr lds_fr(const uint8_t* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void sts_fr(uint8_t* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// kWriteQ: q[i] = carry entering index i's position for the whole range (q[n-1] = carry entering the
// range); otherwise only *d_c0 is produced.  d_flag (optional): set when a coefficient is >= r.
template <bool kWriteQ>
__global__ void __launch_bounds__(kScanThr, 3)
    poly_scan_tiles(const uint32_t* __restrict__ coefs, size_t n, const ScanPow* __restrict__ pw,
                    const uint32_t* __restrict__ carry_in, ScanState st, uint32_t ntiles, uint32_t* __restrict__ q,
                    uint32_t* __restrict__ d_c0, int* __restrict__ d_flag) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ Fr s_warp[kScanWarps];  // warp aggregates
  __shared__ Fr s_part[kScanWarps];  // look-back: weighted sum of each warp's window
  __shared__ int s_found[kScanWarps];
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(st.ticket, 1u);
  __syncthreads();
  const uint32_t b = s_tile;
  const size_t padded = (size_t)ntiles * kScanTile;
  const size_t pad = padded - n;  // zero coefficients above the top one (tile 0 only)
  // tile b covers indices [i_lo, i_lo + TILE), top position first: p' = padded - 1 - i
  const size_t i_lo = (size_t)(ntiles - 1 - b) * kScanTile;

  // ---- load: global 16-byte word W of the tile -> element e = W / 2 -> in-tile position T-1-e ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(coefs + i_lo * 8);
#pragma unroll 8
    for (int r = 0; r < 2 * kScanK; r++) {
      const int W = r * kScanThr + tid;
      const int e = W >> 1;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (i_lo + e < n) val = __ldg(src + W);
      const int pos = kScanTile - 1 - e;
      *reinterpret_cast<uint4*>(smem + (pos / kScanK) * kScanChunkBytes + (pos % kScanK) * 32 + (W & 1) * 16) = val;
    }
  }
  __syncthreads();
  if (b == 0 && pad > 0 && tid == (int)((pad - 1) / kScanK)) {
    // the carry entering the range rides in as the virtual coefficient right above the top one
    sts_fr(smem + tid * kScanChunkBytes + ((pad - 1) % kScanK) * 32, load_fr(carry_in));
  }
  __syncthreads();

  // ---- thread: zero-carry Horner over its K coefficients, as two independent half chains ----
  // (a lone Horner chain leaves the multiply pipe waiting on its own carry chains; two interleaved
  // chains double the instruction-level parallelism for one extra multiply by u^(K/2))
  const Fr u = pw->u;
  const Fr uh = pw->half;  // u^(K/2)
  const uint8_t* mine = smem + tid * kScanChunkBytes;
  Fr h1 = lds_fr(mine);
  Fr h2 = lds_fr(mine + (kScanK / 2) * 32);
  bool bad = !fe_is_canonical(h1) || !fe_is_canonical(h2);
#pragma unroll 2
  for (int k = 1; k < kScanK / 2; k++) {
    const Fr f1 = lds_fr(mine + k * 32);
    const Fr f2 = lds_fr(mine + (kScanK / 2 + k) * 32);
    bad |= !fe_is_canonical(f1) || !fe_is_canonical(f2);
    h1 = fe_add(f1, fe_mul(u, h1));
    h2 = fe_add(f2, fe_mul(u, h2));
  }
  const Fr h = fe_add(h2, fe_mul(uh, h1));
  if (d_flag && bad) atomicOr(d_flag, 1);

  // ---- warp: inclusive Kogge-Stone over lanes ----
  Fr s = h;
#pragma unroll 1
  for (int l = 0; l < 5; l++) {
    const int d = 1 << l;
    const Fr o = shfl_up_fr(s, d);
    const Fr t = fe_add(s, fe_mul(pw->lane[d], o));
    if (lane >= d) s = t;
  }
  if (lane == 31) s_warp[w] = s;
  __syncthreads();

  // ---- block: every warp scans the warp aggregates (lanes 0..kScanWarps-1), so all know H and B ----
  Fr B = lane < kScanWarps ? s_warp[lane] : Fr::zero();
#pragma unroll 1
  for (int d = 1; d < kScanWarps; d <<= 1) {
    const Fr o = shfl_up_fr(B, d);
    const Fr t = fe_add(B, fe_mul(pw->warp[d], o));
    if (lane >= d && lane < kScanWarps) B = t;
  }
  const Fr H = shfl_fr(B, kScanWarps - 1);  // the tile's zero-carry aggregate
  // ---- look-back: the whole block reads kScanThr predecessor tiles per round (thread t: tile j - t).
  // A warp's 32 tiles are resolved once every tile in front of its nearest inclusive carry has at least
  // an aggregate; each warp sums its weighted values, the nearest warp that met an inclusive carry ends
  // the walk.  (A one-warp window could not keep up: the carries then advance 32 tiles per ~3.5 us.)
  Fr X = Fr::zero();  // carry entering the tile
  if (b == 0) {
    if (pad == 0) X = load_fr(carry_in);
  } else {
    if (tid == 0) {
      st.agg[b] = H;
      __threadfence();
      st_release_u32(st.flags + b, 1u);
    }
    Fr scale = Fr::one();
    bool scaled = false;
    for (long long j = (long long)b - 1;; j -= kScanThr) {
      const long long tj = j - tid;
      uint32_t f;
      int Lw;
      while (true) {
        f = tj >= 0 ? ld_relaxed_u32(st.flags + tj) : 0u;
        const uint32_t have_incl = __ballot_sync(0xffffffffu, f == 2u);
        Lw = have_incl ? __ffs(have_incl) - 1 : 32;
        const uint32_t pending = __ballot_sync(0xffffffffu, tj >= 0 && lane < Lw && f == 0u);
        if (!pending) break;
      }
      __threadfence();  // the values were published before the flags
      Fr v = Fr::zero();
      if (tj >= 0 && lane <= Lw) v = fe_mul(pw->tile[tid], ldcg_fr((lane == Lw) ? st.incl + tj : st.agg + tj));
#pragma unroll 1
      for (int m = 16; m > 0; m >>= 1) v = fe_add(v, shfl_xor_fr(v, m));
      if (lane == 0) {
        s_part[w] = v;
        s_found[w] = Lw < 32;
      }
      __syncthreads();
      Fr r = s_part[0];
      bool found = s_found[0] != 0;
#pragma unroll 1
      for (int w2 = 1; w2 < kScanWarps && !found; w2++) {
        r = fe_add(r, s_part[w2]);
        found = s_found[w2] != 0;
      }
      __syncthreads();
      if (scaled) r = fe_mul(scale, r);
      X = fe_add(X, r);
      if (found || j - kScanThr < 0) break;
      scale = scaled ? fe_mul(scale, pw->tile[kScanThr]) : pw->tile[kScanThr];
      scaled = true;
    }
  }
  if (tid == 0) {
    const Fr incl = fe_add(H, fe_mul(pw->tile[1], X));
    st.incl[b] = incl;
    __threadfence();
    st_release_u32(st.flags + b, 2u);
    if (b == ntiles - 1 && !kWriteQ && d_c0) store_fr(d_c0, incl);
  }
  if (!kWriteQ) return;

  // ---- thread: true carry, second Horner pass in place ----
  // carry entering warp w: B_{w-1} + (u^(32 K))^w X; entering the thread: s_{lane-1} + (u^K)^lane * that
  Fr x;
  {
    Fr y = fe_mul(pw->warp[w], X);
    const Fr Bm1 = shfl_fr(B, w > 0 ? w - 1 : 0);
    if (w > 0) y = fe_add(y, Bm1);
    x = fe_mul(pw->lane[lane], y);
  }
  {
    const Fr sm1 = shfl_up_fr(s, 1);
    if (lane > 0) x = fe_add(x, sm1);
  }
  uint8_t* mine_w = smem + tid * kScanChunkBytes;
  // carry entering the second half = zero-carry value of the first half + u^(K/2) x
  Fr x2 = fe_add(h1, fe_mul(uh, x));
#pragma unroll 2
  for (int k = 0; k < kScanK / 2; k++) {
    const Fr f1 = lds_fr(mine_w + k * 32);
    const Fr f2 = lds_fr(mine_w + (kScanK / 2 + k) * 32);
    sts_fr(mine_w + k * 32, x);
    sts_fr(mine_w + (kScanK / 2 + k) * 32, x2);
    x = fe_add(f1, fe_mul(u, x));
    x2 = fe_add(f2, fe_mul(u, x2));
  }
  x = x2;  // carry leaving the thread
  if (b == ntiles - 1 && tid == kScanThr - 1 && d_c0) store_fr(d_c0, x);
  __syncthreads();

  // ---- store ----
  {
    uint4* dst = reinterpret_cast<uint4*>(q + i_lo * 8);
#pragma unroll 8
    for (int r = 0; r < 2 * kScanK; r++) {
      const int W = r * kScanThr + tid;
      const int e = W >> 1;
      if (i_lo + e < n) {
        const int pos = kScanTile - 1 - e;
        dst[W] = *reinterpret_cast<const uint4*>(smem + (pos / kScanK) * kScanChunkBytes + (pos % kScanK) * 32 + (W & 1) * 16);
      }
    }
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
//   [0,32)    u staged     [32,64)  carry staged
// The power table and the look-back state live in ctx->poly_tiles.
static int run_scan(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32], const uint8_t carry_le[32],
                    const uint32_t* d_carry, uint32_t* d_q, uint32_t* d_c0, uint32_t* d_upow, int* d_flag) {
  static_assert(sizeof(ScanPow) % 32 == 0, "table of field elements");
  const size_t ntiles = (n + kScanTile - 1) / kScanTile;
  if (ntiles >= (1ull << 31)) return fail(ctx, MYZKP_ERR_INVALID_ARG, "polynomial too long for the scan");
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  uint8_t* s = ctx->small.as<uint8_t>();
  // pageable 32-byte sources: cudaMemcpyAsync stages them before returning
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s, u_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  if (d_carry) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s + 32, d_carry, 32, cudaMemcpyDeviceToDevice, ctx->stream));
  else if (carry_le) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s + 32, carry_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  else MZ_CUDA_TRY(ctx, cudaMemsetAsync(s + 32, 0, 32, ctx->stream));
  // scratch: ScanPow | agg[ntiles] | incl[ntiles] | flags[ntiles] | ticket
  const size_t flags_off = sizeof(ScanPow) + 2 * ntiles * sizeof(Fr);
  MZ_CUDA_TRY(ctx, ctx->poly_tiles.ensure(flags_off + (ntiles + 1) * sizeof(uint32_t) + 64));
  uint8_t* base = ctx->poly_tiles.as<uint8_t>();
  ScanPow* pw = reinterpret_cast<ScanPow*>(base);
  ScanState st;
  st.agg = reinterpret_cast<Fr*>(base + sizeof(ScanPow));
  st.incl = st.agg + ntiles;
  st.flags = reinterpret_cast<uint32_t*>(base + flags_off);
  st.ticket = st.flags + ntiles;
  MZ_CUDA_TRY(ctx, cudaMemsetAsync(st.flags, 0, (ntiles + 1) * sizeof(uint32_t), ctx->stream));
  poly_scan_powers<<<1, kScanPowThreads, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(s), pw, (uint64_t)n, d_upow);
  MZ_LAUNCH_CHECK(ctx);
  if (n == 0) return MYZKP_OK;
  static bool attr_set[64] = {};
  if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
    MZ_CUDA_TRY(ctx, cudaFuncSetAttribute(poly_scan_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanSmemBytes));
    MZ_CUDA_TRY(ctx, cudaFuncSetAttribute(poly_scan_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanSmemBytes));
    attr_set[ctx->device] = true;
  }
  const uint32_t* cin = reinterpret_cast<const uint32_t*>(s + 32);
  if (d_q)
    poly_scan_tiles<true><<<(unsigned)ntiles, kScanThr, kScanSmemBytes, ctx->stream>>>(d_coefs, n, pw, cin, st, (uint32_t)ntiles,
                                                                                        d_q, d_c0, d_flag);
  else
    poly_scan_tiles<false><<<(unsigned)ntiles, kScanThr, kScanSmemBytes, ctx->stream>>>(d_coefs, n, pw, cin, st,
                                                                                         (uint32_t)ntiles, nullptr, d_c0, d_flag);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

// (h, u^n) of a coefficient range: a reduction - nothing but the two 32-byte results is written
int fr_range_eval(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32], uint32_t* d_h,
                  uint32_t* d_upow, int* d_flag) {
  if (n == 0) {  // empty range: h = 0, u^0 = 1; no scratch is touched (an empty rank must not allocate)
    static const uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(d_h, 0, 32, ctx->stream));
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(d_upow, one, 32, cudaMemcpyHostToDevice, ctx->stream));
    return MYZKP_OK;
  }
  return run_scan(ctx, d_coefs, n, u_le, nullptr, nullptr, nullptr, d_h, d_upow, d_flag);
}

int fr_range_quotient(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32],
                      const uint8_t carry_le[32], uint32_t* d_q, uint32_t* d_c0, const uint32_t* d_carry, int* d_flag) {
  if (n == 0) return MYZKP_OK;
  return run_scan(ctx, d_coefs, n, u_le, carry_le, d_carry, d_q, d_c0, nullptr, d_flag);
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

// Resident SRS table (PublicKeyKZG.powers_1, kzg.rs:8-11) and its generation.
//
// setup_kzg (kzg.rs:27-40) computes powers_1[i] = [alpha^i]G with one
// double-and-add scalar multiplication per point.  Here: alpha^i on the device
// (Fr), a fixed-base comb of G (32 x 255 precomputed multiples, 8-bit digits,
// <= 32 mixed adds per point), batched to-affine, and then the table rows
// row[j][i] = 2^(b_j) * P_i that let every MSM window reuse one bucket set.
#include <cstring>

#include "ctx.cuh"

namespace mz {

__device__ __forceinline__ Fq load_fq(const uint32_t* p) {
  Fq r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void store_fq(uint32_t* p, const Fq& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// canonical LE bytes -> Montgomery affine (row 0); flags non-canonical input
__global__ void srs_import(const uint32_t* in, size_t n, Affine* row0, int* flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = load_fq(in + i * 16), y = load_fq(in + i * 16 + 8);
  if (!fe_is_canonical(x) || !fe_is_canonical(y)) atomicOr(flag, 1);
  Affine a;
  a.x = fe_to_mont(x);
  a.y = fe_to_mont(y);
  row0[i] = a;
}

// Montgomery affine -> canonical LE bytes
__global__ void srs_export(const Affine* row0, size_t n, uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine a = row0[i];
  store_fq(out + i * 16, fe_from_mont(a.x));
  store_fq(out + i * 16 + 8, fe_from_mont(a.y));
}

// rows 1..rows-1 from row 0: Jacobian doubling chain, one batched inversion per point
__global__ void __launch_bounds__(128) srs_build_rows(Affine* tbl, size_t n, int rows,
                                                      const int* __restrict__ row_bits) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine p = tbl[i];
  if (affine_is_inf(p)) {
    for (int j = 1; j < rows; j++) tbl[(size_t)j * n + i] = p;
    return;
  }
  Jac cur;
  cur.x = p.x; cur.y = p.y; cur.z = Fq::one();
  Fq zs[kMaxTableRows - 1];
  Fq pref[kMaxTableRows - 1];
  for (int j = 1; j < rows; j++) {
#pragma unroll 1
    for (int d = row_bits[j - 1]; d < row_bits[j]; d++) jac_dbl(cur);
    Affine un;  // unnormalised X, Y parked in the table slot
    un.x = cur.x; un.y = cur.y;
    tbl[(size_t)j * n + i] = un;
    zs[j - 1] = cur.z;
    pref[j - 1] = (j == 1) ? cur.z : fe_mul(pref[j - 2], cur.z);
  }
  Fq inv = fe_inv_bingcd(pref[rows - 2]);
  for (int k = rows - 2; k >= 0; k--) {
    Fq zinv = (k > 0) ? fe_mul(inv, pref[k - 1]) : inv;
    inv = fe_mul(inv, zs[k]);
    Fq z2 = fe_sqr(zinv);
    Affine un = tbl[(size_t)(k + 1) * n + i];
    Affine o;
    o.x = fe_mul(un.x, z2);
    o.y = fe_mul(un.y, fe_mul(z2, zinv));
    tbl[(size_t)(k + 1) * n + i] = o;
  }
}

// XYZZ -> Montgomery affine, 8 points per thread share one inversion
constexpr int kBatchAffine = 8;
__global__ void __launch_bounds__(128) batch_to_affine(const XYZZ* in, size_t n, Affine* out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * kBatchAffine;
  if (lo >= n) return;
  int cnt = (n - lo) < (size_t)kBatchAffine ? (int)(n - lo) : kBatchAffine;
  Fq pref[kBatchAffine];
  Fq acc = Fq::one();
  for (int k = 0; k < cnt; k++) {
    Fq z = in[lo + k].zzz;
    if (z.is_zero()) z = Fq::one();  // infinity: keep the product invertible
    acc = fe_mul(acc, z);
    pref[k] = acc;
  }
  Fq inv = fe_inv_bingcd(acc);
  for (int k = cnt - 1; k >= 0; k--) {
    XYZZ p = in[lo + k];
    Affine o;
    if (xyzz_is_inf(p)) {
      o.x = Fq::zero(); o.y = Fq::zero();
    } else {
      Fq zi = (k > 0) ? fe_mul(inv, pref[k - 1]) : inv;
      inv = fe_mul(inv, p.zzz);
      o = xyzz_to_affine_with_inv(p, zi);
    }
    out[lo + k] = o;
  }
}

// comb[j*256 + d] = d * base2[j]  (base2[j] = 2^(8j) G, affine), d = 0..255 as XYZZ
constexpr int kCombRows = 32;
__global__ void srs_comb_multiples(const Affine* base2, XYZZ* out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kCombRows * 256) return;
  int j = idx >> 8, d = idx & 255;
  Affine b = base2[j];
  XYZZ acc = xyzz_inf();
  for (int bit = 7; bit >= 0; bit--) {
    xyzz_dbl(acc);
    if ((d >> bit) & 1) xyzz_madd(acc, b);
  }
  out[idx] = acc;
}

// scalars[i] = alpha^(first+i) canonical
__global__ void srs_alpha_powers(const uint32_t* alpha_canon, size_t first, size_t n, uint32_t* scalars) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr a;
#pragma unroll
  for (int k = 0; k < 8; k++) a.v[k] = alpha_canon[k];
  a = fe_to_mont(a);
  Fr r = fe_from_mont(fe_pow_u64(a, (uint64_t)(first + i)));
  uint4* q = reinterpret_cast<uint4*>(scalars + i * 8);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// out[i] = [scalars[i]] G via the comb (unsigned 8-bit digits)
__global__ void __launch_bounds__(128) srs_fixed_base(const uint32_t* scalars, size_t n, const Affine* comb,
                                                      XYZZ* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
#pragma unroll
  for (int k = 0; k < 8; k++) s[k] = scalars[i * 8 + k];
  XYZZ acc = xyzz_inf();
#pragma unroll 1
  for (int j = 0; j < kCombRows; j++) {
    uint32_t d = (s[j >> 2] >> ((j & 3) * 8)) & 255u;
    if (d) xyzz_madd(acc, comb[j * 256 + d]);
  }
  out[i] = acc;
}

__global__ void srs_set_generator(Affine* out) {
  Affine g;
  Fq one = Fq::one();
  g.x = one;                // G = (1, 2)  (bn128.rs:185-188)
  g.y = fe_dbl(one);
  out[0] = g;
}

// ---------------------------------------------------------------------------
// choose the supported windows / table rows for an SRS of n points
static void plan_rows(myzkp_ctx* ctx, uint32_t windows) {
  bool used[256] = {};
  for (int c = 1; c <= 24; c++) {
    if (!((windows >> c) & 1)) continue;
    const int W = (255 + c - 1) / c;
    for (int w = 0; w < W; w++) used[c * w] = true;  // c*w <= 252 for every c here
  }
  ctx->windows = windows;
  ctx->table_rows = 0;
  for (int b = 0; b < 256; b++) {
    ctx->row_of_bit[b] = 0xff;
    if (used[b]) {
      ctx->row_of_bit[b] = (uint8_t)ctx->table_rows;
      ctx->row_bits[ctx->table_rows++] = b;
    }
  }
}

int srs_alloc(myzkp_ctx* ctx, size_t n) {
  if (ctx->table) {
    cudaFree(ctx->table);
    ctx->table = nullptr;
    ctx->srs_n = 0;
  }
  if (n == 0) return MYZKP_OK;
  size_t free_b = 0, total_b = 0;
  MZ_CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
  const uint32_t full = (1u << 4) | (1u << 8) | (1u << 12) | (1u << 16) | (1u << 20) | (1u << 22) | (1u << 24);
  const uint32_t lean = (1u << 8) | (1u << 16) | (1u << 24);
  plan_rows(ctx, full);
  if ((uint64_t)n * ctx->table_rows >= (1ull << 31) ||
      (double)n * ctx->table_rows * sizeof(Affine) > 0.55 * (double)free_b)
    plan_rows(ctx, lean);
  if (ctx->table_windows) plan_rows(ctx, ctx->table_windows);  // caller's choice (myzkp_ctx_set_table_windows)
  if ((uint64_t)n * ctx->table_rows >= (1ull << 31)) return fail(ctx, MYZKP_ERR_INVALID_ARG, "SRS too large (n * rows must be < 2^31)");
  if (!ctx->d_row_of_bit) {
    MZ_CUDA_TRY(ctx, cudaMalloc(&ctx->d_row_of_bit, 256));
    MZ_CUDA_TRY(ctx, cudaMalloc(&ctx->d_row_bits, sizeof(int) * kMaxTableRows));
  }
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_row_of_bit, ctx->row_of_bit, 256, cudaMemcpyHostToDevice, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_row_bits, ctx->row_bits, sizeof(int) * kMaxTableRows, cudaMemcpyHostToDevice, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMalloc(&ctx->table, n * ctx->table_rows * sizeof(Affine)));
  ctx->srs_n = n;
  return MYZKP_OK;
}

int srs_build_from_row0(myzkp_ctx* ctx) {
  size_t n = ctx->srs_n;
  if (n == 0) return MYZKP_OK;
  srs_build_rows<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->table, n, ctx->table_rows, ctx->d_row_bits);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int points_import(myzkp_ctx* ctx, const uint32_t* d_in, size_t n, Affine* d_out, int* d_flag) {
  if (n == 0) return MYZKP_OK;
  srs_import<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_in, n, d_out, d_flag);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

static int ensure_gcomb(myzkp_ctx* ctx) {
  if (ctx->gcomb) return MYZKP_OK;
  // base2[j] = 2^(8j) G through the same row builder (n = 1)
  Affine* base2 = nullptr;
  MZ_CUDA_TRY(ctx, cudaMalloc(&base2, kCombRows * sizeof(Affine)));
  srs_set_generator<<<1, 1, 0, ctx->stream>>>(base2);
  MZ_LAUNCH_CHECK(ctx);
  int comb_bits[kCombRows];
  for (int j = 0; j < kCombRows; j++) comb_bits[j] = 8 * j;
  int* d_comb_bits = nullptr;
  MZ_CUDA_TRY(ctx, cudaMalloc(&d_comb_bits, sizeof(comb_bits)));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(d_comb_bits, comb_bits, sizeof(comb_bits), cudaMemcpyHostToDevice, ctx->stream));
  srs_build_rows<<<1, 128, 0, ctx->stream>>>(base2, 1, kCombRows, d_comb_bits);
  MZ_LAUNCH_CHECK(ctx);
  const int total = kCombRows * 256;
  MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure((size_t)total * sizeof(XYZZ)));
  srs_comb_multiples<<<(total + 127) / 128, 128, 0, ctx->stream>>>(base2, ctx->xyzz_tmp.as<XYZZ>());
  MZ_LAUNCH_CHECK(ctx);
  MZ_CUDA_TRY(ctx, cudaMalloc(&ctx->gcomb, (size_t)total * sizeof(Affine)));
  unsigned blocks = (unsigned)((total + kBatchAffine * 128 - 1) / (kBatchAffine * 128));
  batch_to_affine<<<blocks, 128, 0, ctx->stream>>>(ctx->xyzz_tmp.as<XYZZ>(), total, ctx->gcomb);
  MZ_LAUNCH_CHECK(ctx);
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(base2);
  cudaFree(d_comb_bits);
  return MYZKP_OK;
}

}  // namespace mz

using namespace mz;

namespace {
// Loading / generating an SRS is set-up work (it allocates the table anyway, which synchronises the device): the
// scratch freeze of same-device peers (DevBuf::frozen) is lifted for its duration.
struct SetupScope {
  myzkp_ctx* ctx;
  bool was;
  explicit SetupScope(myzkp_ctx* c) : ctx(c), was(c->peer_same_device) { c->peer_same_device = false; }
  ~SetupScope() { ctx->peer_same_device = was; }
};
}  // namespace

extern "C" int myzkp_srs_load_g1(myzkp_ctx* ctx, const uint8_t* affine_xy_le, size_t n) {
  if (!ctx || (!affine_xy_le && n)) return MYZKP_ERR_INVALID_ARG;
  SetupScope setup(ctx);
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_TRY(srs_alloc(ctx, n));
  if (n == 0) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 64));
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  int* flag = reinterpret_cast<int*>(ctx->small.as<uint8_t>() + 528);  // own word: 512 is the sticky scalar flag
  MZ_CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, affine_xy_le, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  srs_import<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->scalars.as<uint32_t>(), n, ctx->table, flag);
  MZ_LAUNCH_CHECK(ctx);
  MZ_TRY(srs_build_from_row0(ctx));
  int h_flag = 0;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flag) {
    srs_alloc(ctx, 0);
    return fail(ctx, MYZKP_ERR_NONCANONICAL, "SRS coordinate >= p");
  }
  return MYZKP_OK;
}

extern "C" int myzkp_srs_generate_g1(myzkp_ctx* ctx, const uint8_t alpha_le[32], size_t first, size_t n) {
  if (!ctx || !alpha_le) return MYZKP_ERR_INVALID_ARG;
  SetupScope setup(ctx);
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_TRY(ensure_gcomb(ctx));
  MZ_TRY(srs_alloc(ctx, n));
  if (n == 0) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  uint8_t* s = ctx->small.as<uint8_t>();
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s, alpha_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  // chunked so the XYZZ temporaries stay bounded
  const size_t chunk = (size_t)1 << 22;
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure((n < chunk ? n : chunk) * 32));
  MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure((n < chunk ? n : chunk) * sizeof(XYZZ)));
  for (size_t off = 0; off < n; off += chunk) {
    size_t m = n - off < chunk ? n - off : chunk;
    unsigned blocks = (unsigned)((m + 127) / 128);
    srs_alpha_powers<<<blocks, 128, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(s), first + off, m,
                                                      ctx->scalars.as<uint32_t>());
    MZ_LAUNCH_CHECK(ctx);
    srs_fixed_base<<<blocks, 128, 0, ctx->stream>>>(ctx->scalars.as<uint32_t>(), m, ctx->gcomb,
                                                    ctx->xyzz_tmp.as<XYZZ>());
    MZ_LAUNCH_CHECK(ctx);
    unsigned bblocks = (unsigned)((m + kBatchAffine * 128 - 1) / (kBatchAffine * 128));
    batch_to_affine<<<bblocks, 128, 0, ctx->stream>>>(ctx->xyzz_tmp.as<XYZZ>(), m, ctx->table + off);
    MZ_LAUNCH_CHECK(ctx);
  }
  MZ_TRY(srs_build_from_row0(ctx));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MYZKP_OK;
}

extern "C" int myzkp_srs_read_g1(myzkp_ctx* ctx, size_t off, size_t n, uint8_t* out) {
  if (!ctx || (!out && n)) return MYZKP_ERR_INVALID_ARG;
  if (off + n > ctx->srs_n) return fail(ctx, MYZKP_ERR_INVALID_ARG, "srs_read out of range");
  if (n == 0) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 64));
  srs_export<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->table + off, n, ctx->scalars.as<uint32_t>());
  MZ_LAUNCH_CHECK(ctx);
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->scalars.p, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MYZKP_OK;
}

extern "C" size_t myzkp_srs_len(const myzkp_ctx* ctx) { return ctx ? ctx->srs_n : 0; }

extern "C" int myzkp_ctx_set_table_windows(myzkp_ctx* ctx, uint32_t window_mask) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  // bit c set <=> window c is to be supported; 1 <= c <= 24; the row count must fit kMaxTableRows
  if (window_mask & ~0x01fffffeu) return fail(ctx, MYZKP_ERR_INVALID_ARG, "windows are 1..24 bits");
  if (window_mask) {
    bool used[256] = {};
    int rows = 0;
    for (int c = 1; c <= 24; c++)
      if ((window_mask >> c) & 1)
        for (int w = 0; w < (255 + c - 1) / c; w++)
          if (!used[c * w]) { used[c * w] = true; rows++; }
    if (rows > kMaxTableRows) return fail(ctx, MYZKP_ERR_INVALID_ARG, "window set needs too many table rows");
  }
  ctx->table_windows = window_mask;
  return MYZKP_OK;
}

extern "C" int myzkp_srs_table_info(const myzkp_ctx* ctx, int* out_rows, uint64_t* out_bytes, uint32_t* out_windows) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  if (out_rows) *out_rows = ctx->table ? ctx->table_rows : 0;
  if (out_bytes) *out_bytes = ctx->table ? (uint64_t)ctx->srs_n * ctx->table_rows * sizeof(Affine) : 0;
  if (out_windows) *out_windows = ctx->table ? ctx->windows : 0;
  return MYZKP_OK;
}

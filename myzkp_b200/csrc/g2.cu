// G2 entry points: the G2 half of the KZG public key (setup_kzg powers_2, kzg.rs:37; setup_kzg_with_full_g2,
// kzg.rs:42-55) and the G2 MSM of the accumulate_curve_points call sites (zksnark/utils.rs:83-93).
// Device arithmetic in g2.cuh.  Its own translation unit: these kernels inline long Fq2 formula chains
// and dominate the build time.
#include <cstring>

#include "ctx.cuh"
#include "g2.cuh"

// ---------------------------------------------------------------------------
// G2 half of the public key: out[i] = [alpha^(first+i)] base  (kzg.rs:37, 47-52)
// ---------------------------------------------------------------------------
namespace mz {
// io layout per point: x.c0 | x.c1 | y.c0 | y.c1, 8 canonical little-endian limbs each; infinity = zeros
__device__ __forceinline__ Fq g2_load_fq(const uint32_t* raw, int* flag) {
  Fq a;
#pragma unroll
  for (int k = 0; k < 8; k++) a.v[k] = raw[k];
  if (!fe_is_canonical(a)) atomicOr(flag, 1);
  return fe_to_mont(a);
}
__global__ void __launch_bounds__(64) srs_g2_powers(const uint32_t* alpha_canon, const uint32_t* base_raw, size_t first,
                                                    size_t n, uint32_t* out, int* flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr a;
#pragma unroll
  for (int k = 0; k < 8; k++) a.v[k] = alpha_canon[k];
  if (!fe_is_canonical(a)) atomicOr(flag, 1);
  Fr e = fe_from_mont(fe_pow_u64(fe_to_mont(a), (uint64_t)(first + i)));
  AffineG2 b;
  b.x.c0 = g2_load_fq(base_raw, flag);
  b.x.c1 = g2_load_fq(base_raw + 8, flag);
  b.y.c0 = g2_load_fq(base_raw + 16, flag);
  b.y.c1 = g2_load_fq(base_raw + 24, flag);
  AffineG2 r = g2_scalar_mul(b, e.v);
  const Fq c[4] = {fe_from_mont(r.x.c0), fe_from_mont(r.x.c1), fe_from_mont(r.y.c0), fe_from_mont(r.y.c1)};
#pragma unroll
  for (int q = 0; q < 4; q++)
#pragma unroll
    for (int k = 0; k < 8; k++) out[i * 32 + q * 8 + k] = c[q].v[k];
}

// G2 MSM for the small accumulate_curve_points call sites over G2 (zksnark/utils.rs:83-93, e.g.
// tutorial_snark/protocol_2.rs:68): one thread per term (double-and-add, Jacobian), then a block tree
// sum.  part[blockIdx.x] = sum of the block's terms.
constexpr int kG2Threads = 64;
__global__ void __launch_bounds__(kG2Threads) g2_msm_terms(const uint32_t* scalars, const uint32_t* points_raw, size_t n,
                                                           JacG2* part, int* flag) {
  __shared__ JacG2 sm[kG2Threads];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  JacG2 acc = g2_jac_inf();
  if (i < n) {
    Fr k;
#pragma unroll
    for (int q = 0; q < 8; q++) k.v[q] = scalars[i * 8 + q];
    if (!fe_is_canonical(k)) atomicOr(flag, 1);
    AffineG2 b;
    b.x.c0 = g2_load_fq(points_raw + i * 32, flag);
    b.x.c1 = g2_load_fq(points_raw + i * 32 + 8, flag);
    b.y.c0 = g2_load_fq(points_raw + i * 32 + 16, flag);
    b.y.c1 = g2_load_fq(points_raw + i * 32 + 24, flag);
    acc = g2_scalar_mul_jac(b, k.v);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
#pragma unroll 1
  for (int d = kG2Threads / 2; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) {
      JacG2 o = sm[threadIdx.x + d];
      g2_jac_add(acc, o);
      sm[threadIdx.x] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
// single block: sums `count` partials and writes the affine result in wire form
__global__ void __launch_bounds__(kG2Threads) g2_sum_partials(const JacG2* part, size_t count, uint32_t* out) {
  __shared__ JacG2 sm[kG2Threads];
  JacG2 acc = g2_jac_inf();
  for (size_t i = threadIdx.x; i < count; i += kG2Threads) {
    JacG2 o = part[i];
    g2_jac_add(acc, o);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
#pragma unroll 1
  for (int d = kG2Threads / 2; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) {
      JacG2 o = sm[threadIdx.x + d];
      g2_jac_add(acc, o);
      sm[threadIdx.x] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    AffineG2 r = g2_jac_to_affine(acc);
    const Fq c[4] = {fe_from_mont(r.x.c0), fe_from_mont(r.x.c1), fe_from_mont(r.y.c0), fe_from_mont(r.y.c1)};
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int k = 0; k < 8; k++) out[q * 8 + k] = c[q].v[k];
  }
}
}  // namespace mz

// BN128::generator_g2() (bn128.rs:190-205): x.c0, x.c1, y.c0, y.c1 as little-endian u32 limbs
static const uint32_t kG2Generator[32] = {
    0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu,
    0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u,
    0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u,
    0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u,
};

extern "C" int myzkp_srs_generate_g2(myzkp_ctx* ctx, const uint8_t alpha_le[32], const uint8_t* base_or_null, size_t first,
                                     size_t n, uint8_t* out) {
  using namespace mz;
  if (!ctx || !alpha_le || (!out && n)) return MYZKP_ERR_INVALID_ARG;
  if (n == 0) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  uint8_t* s = ctx->small.as<uint8_t>();
  int* flag = reinterpret_cast<int*>(s + 512);
  MZ_CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s, alpha_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(s + 128, base_or_null ? (const void*)base_or_null : (const void*)kG2Generator, 128,
                                   cudaMemcpyHostToDevice, ctx->stream));
  const size_t chunk = (size_t)1 << 20;  // staging buffer of 128 MiB at most
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure((n < chunk ? n : chunk) * 128));
  int h_flag = 0;
  for (size_t off = 0; off < n; off += chunk) {
    const size_t m = n - off < chunk ? n - off : chunk;
    srs_g2_powers<<<(unsigned)((m + 63) / 64), 64, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(s),
                                                                    reinterpret_cast<const uint32_t*>(s + 128), first + off, m,
                                                                    ctx->scalars.as<uint32_t>(), flag);
    MZ_LAUNCH_CHECK(ctx);
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out + off * 128, ctx->scalars.p, m * 128, cudaMemcpyDeviceToHost, ctx->stream));
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_flag) return fail(ctx, MYZKP_ERR_NONCANONICAL, "alpha >= r or G2 base coordinate >= p");
  }
  return MYZKP_OK;
}

extern "C" int myzkp_g2_msm(myzkp_ctx* ctx, const uint8_t* scalars_le, const uint8_t* points /* n*128 */, size_t n,
                            uint8_t out[128]) {
  using namespace mz;
  if (!ctx || !out || (n && (!scalars_le || !points))) return MYZKP_ERR_INVALID_ARG;
  if (n == 0) {  // empty sum = point at infinity (the fold's initial value, zksnark/utils.rs:89)
    memset(out, 0, 128);
    return MYZKP_OK;
  }
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  uint8_t* s = ctx->small.as<uint8_t>();
  int* flag = reinterpret_cast<int*>(s + 512);
  const size_t blocks = (n + kG2Threads - 1) / kG2Threads;
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
  MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 128));
  MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure(blocks * sizeof(JacG2)));
  MZ_CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, scalars_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars2.p, points, n * 128, cudaMemcpyHostToDevice, ctx->stream));
  g2_msm_terms<<<(unsigned)blocks, kG2Threads, 0, ctx->stream>>>(ctx->scalars.as<uint32_t>(), ctx->scalars2.as<uint32_t>(), n,
                                                                 ctx->xyzz_tmp.as<JacG2>(), flag);
  MZ_LAUNCH_CHECK(ctx);
  g2_sum_partials<<<1, kG2Threads, 0, ctx->stream>>>(ctx->xyzz_tmp.as<JacG2>(), blocks, reinterpret_cast<uint32_t*>(s + 1024));
  MZ_LAUNCH_CHECK(ctx);
  int h_flag = 0;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out, s + 1024, 128, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flag) return fail(ctx, MYZKP_ERR_NONCANONICAL, "scalar >= r or G2 coordinate >= p");
  return MYZKP_OK;
}

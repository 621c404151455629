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

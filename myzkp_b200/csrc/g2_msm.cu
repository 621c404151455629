// G2 MSM of the accumulate_curve_points call sites over G2 (zksnark/utils.rs:83-93).  Device arithmetic in g2.cuh.
// Split from g2.cu so the two long compilations run side by side.
#include <cstring>

#include "ctx.cuh"
#include "g2.cuh"

namespace mz {
// io layout per point: x.c0 | x.c1 | y.c0 | y.c1, 8 canonical little-endian limbs each; infinity = zeros
__device__ __forceinline__ Fq g2_load_fq(const uint32_t* raw, int* flag) {
  Fq a;
#pragma unroll
  for (int k = 0; k < 8; k++) a.v[k] = raw[k];
  if (!fe_is_canonical(a)) atomicOr(flag, 1);
  return fe_to_mont(a);
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

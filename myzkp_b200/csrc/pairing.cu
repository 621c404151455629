// Pairing kernels behind myzkp_pairing / myzkp_pairing_product_is_one: the verifier half of KZG
// (verify_kzg kzg.rs:90-102, batch_verify_kzg :104-119, verify_degree_bound :136-144), i.e. the
// reference's optimal_ate_pairing (curve/bn128.rs:147-181).  One warp per pairing; see pairing.cuh.
#include <cstring>

#include "ctx.cuh"
#include "pairing.cuh"

namespace mz {

// Final exponentiation by parts (pairing.cuh, ~13x fewer Fq12 products): validated on the host emulation against the
// plain power and on a B200 against the golden Fq12 values and the verifier tests (round 2); 0 selects the plain power.
#ifndef MZ_PAIRING_FAST_FINAL_EXP
#define MZ_PAIRING_FAST_FINAL_EXP 1
#endif

struct PairingSmem {
  F12 f, l, base, acc;
  Fq lo[12], hi[12];
#if MZ_PAIRING_FAST_FINAL_EXP
  F12 w[9];
#endif
};

template <class Exec>
__device__ __forceinline__ void final_exp(Exec& ex, PairingSmem& sm) {
#if MZ_PAIRING_FAST_FINAL_EXP
  pairing_final_exp_fast(ex, sm.f, sm.w);
#else
  pairing_final_exp(ex, sm.f, sm.base, sm.acc);
#endif
}

__device__ __forceinline__ Fq pairing_load_fq(const uint32_t* raw, int* flag) {
  Fq a;
#pragma unroll
  for (int k = 0; k < 8; k++) a.v[k] = raw[k];
  if (!fe_is_canonical(a)) atomicOr(flag, 1);
  return fe_to_mont(a);
}

// blockIdx.x = pairing index.  full != 0: out[i] = e(P_i, Q_i) as 12 canonical coefficients of w^k;
// full == 0: out[i] = the Miller value in Montgomery form (for pairing_product_check).
__global__ void __launch_bounds__(32) pairing_kernel(const uint32_t* g1_raw, const uint32_t* g2_raw, int full, uint32_t* out,
                                                     int* flag) {
  __shared__ PairingSmem sm;
  __shared__ Affine p;
  __shared__ AffineG2 q;
  const int lane = threadIdx.x;
  const size_t i = blockIdx.x;
  if (lane == 0) {
    p.x = pairing_load_fq(g1_raw + i * 16, flag);
    p.y = pairing_load_fq(g1_raw + i * 16 + 8, flag);
    q.x.c0 = pairing_load_fq(g2_raw + i * 32, flag);
    q.x.c1 = pairing_load_fq(g2_raw + i * 32 + 8, flag);
    q.y.c0 = pairing_load_fq(g2_raw + i * 32 + 16, flag);
    q.y.c1 = pairing_load_fq(g2_raw + i * 32 + 24, flag);
  }
  __syncwarp();
  WarpExec ex{sm.lo, sm.hi};
  pairing_miller(ex, sm.f, sm.l, p, q);
  if (full) final_exp(ex, sm);
  if (lane < 12) {
    Fq c = full ? fe_from_mont(sm.f.c[lane]) : sm.f.c[lane];
#pragma unroll
    for (int k = 0; k < 8; k++) out[i * 96 + lane * 8 + k] = c.v[k];
  }
}

// one warp: *result = (prod_i miller[i])^((p^12-1)/r) == 1
__global__ void __launch_bounds__(32) pairing_product_check(const uint32_t* miller, size_t n, int* result) {
  __shared__ PairingSmem sm;
  const int lane = threadIdx.x;
  WarpExec ex{sm.lo, sm.hi};
  if (lane == 0) f12_set_one(sm.f);
  __syncwarp();
  for (size_t i = 0; i < n; i++) {
    if (lane < 12) {
#pragma unroll
      for (int k = 0; k < 8; k++) sm.l.c[lane].v[k] = miller[i * 96 + lane * 8 + k];
    }
    __syncwarp();
    ex.mul(sm.f, sm.f, sm.l);
  }
  final_exp(ex, sm);
  if (lane == 0) {
    bool one = sm.f.c[0] == Fq::one();
    for (int k = 1; k < 12; k++) one = one && sm.f.c[k].is_zero();
    *result = one ? 1 : 0;
  }
}

static int run_pairings(myzkp_ctx* ctx, const uint8_t* g1, const uint8_t* g2, size_t n, int full, uint32_t** d_out) {
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  int* flag = reinterpret_cast<int*>(ctx->small.as<uint8_t>() + 512);
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 64));
  MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 128));
  MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure(n * 384));
  MZ_CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, g1, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars2.p, g2, n * 128, cudaMemcpyHostToDevice, ctx->stream));
  pairing_kernel<<<(unsigned)n, 32, 0, ctx->stream>>>(ctx->scalars.as<uint32_t>(), ctx->scalars2.as<uint32_t>(), full,
                                                      ctx->xyzz_tmp.as<uint32_t>(), flag);
  MZ_LAUNCH_CHECK(ctx);
  *d_out = ctx->xyzz_tmp.as<uint32_t>();
  return MYZKP_OK;
}

static int finish_flag(myzkp_ctx* ctx) {
  int* flag = reinterpret_cast<int*>(ctx->small.as<uint8_t>() + 512);
  int h_flag = 0;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flag) return fail(ctx, MYZKP_ERR_NONCANONICAL, "pairing input coordinate >= p");
  return MYZKP_OK;
}

}  // namespace mz

extern "C" int myzkp_pairing(myzkp_ctx* ctx, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out) {
  using namespace mz;
  if (!ctx || (n && (!g1 || !g2 || !out))) return MYZKP_ERR_INVALID_ARG;
  if (n == 0) return MYZKP_OK;
  if (n > 65535) return fail(ctx, MYZKP_ERR_INVALID_ARG, "at most 65535 pairings per call");
  uint32_t* d_out = nullptr;
  MZ_TRY(run_pairings(ctx, g1, g2, n, 1, &d_out));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_out, n * 384, cudaMemcpyDeviceToHost, ctx->stream));
  return finish_flag(ctx);
}

extern "C" int myzkp_pairing_product_is_one(myzkp_ctx* ctx, const uint8_t* g1, const uint8_t* g2, size_t n, int* out_is_one) {
  using namespace mz;
  if (!ctx || !out_is_one || (n && (!g1 || !g2))) return MYZKP_ERR_INVALID_ARG;
  if (n == 0) {  // empty product
    *out_is_one = 1;
    return MYZKP_OK;
  }
  if (n > 65535) return fail(ctx, MYZKP_ERR_INVALID_ARG, "at most 65535 pairings per call");
  uint32_t* d_miller = nullptr;
  MZ_TRY(run_pairings(ctx, g1, g2, n, 0, &d_miller));
  int* d_res = reinterpret_cast<int*>(ctx->small.as<uint8_t>() + 768);
  pairing_product_check<<<1, 32, 0, ctx->stream>>>(d_miller, n, d_res);
  MZ_LAUNCH_CHECK(ctx);
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_is_one, d_res, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  return finish_flag(ctx);
}

// Batched field / group operations exposed for the parity tests
// (tests/test_gpu_field_group.py): the same device functions the MSM uses,
// driven directly so they can be compared with the oracle element by element.
#include "ctx.cuh"

using namespace mz;

namespace {

template <class F>
__device__ __forceinline__ F load_fe(const uint32_t* p) {
  F r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = p[i];
  return r;
}
template <class F>
__device__ __forceinline__ void store_fe(uint32_t* p, const F& r) {
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = r.v[i];
}

template <class F>
__global__ void test_field_kernel(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = fe_to_mont(load_fe<F>(a + i * 8));
  F y = b ? fe_to_mont(load_fe<F>(b + i * 8)) : F::zero();
  F r;
  switch (op) {
    case 0: r = fe_add(x, y); break;
    case 1: r = fe_sub(x, y); break;
    case 2: r = fe_mul(x, y); break;
    case 3: r = fe_inv(x); break;
    case 5: r = fe_inv_bingcd(x); break;
    default: r = fe_neg(x); break;
  }
  store_fe(out + i * 8, fe_from_mont(r));
}

__device__ __forceinline__ Affine load_point_bytes(const uint32_t* p) {
  Affine r;
  r.x = fe_to_mont(load_fe<Fq>(p));
  r.y = fe_to_mont(load_fe<Fq>(p + 8));
  return r;
}
__device__ __forceinline__ void store_point_bytes(uint32_t* p, const XYZZ& v) {
  Affine a = xyzz_to_affine(v);
  store_fe(p, fe_from_mont(a.x));
  store_fe(p + 8, fe_from_mont(a.y));
}

__global__ void test_g1_kernel(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine pa = load_point_bytes(a + i * 16);
  XYZZ acc = xyzz_from_affine(pa);
  if (op == 0) {
    xyzz_madd(acc, load_point_bytes(b + i * 16));
  } else if (op == 1) {
    xyzz_dbl(acc);
  } else if (op == 2) {
    // [k]a with k = the 256-bit integer in the first 32 bytes of b[i]; MSB first
    XYZZ r = xyzz_inf();
    for (int limb = 7; limb >= 0; limb--) {
      uint32_t w = b[i * 16 + limb];
      for (int bit = 31; bit >= 0; bit--) {
        xyzz_dbl(r);
        if ((w >> bit) & 1) xyzz_madd(r, pa);
      }
    }
    acc = r;
  } else {
    // XYZZ + XYZZ with non-trivial denominators on both sides
    Affine pb = load_point_bytes(b + i * 16);
    XYZZ q = xyzz_from_affine(pb);
    Fq t = fe_add(pa.x, Fq::one());  // arbitrary non-zero-ish scale factors
    Fq s = fe_add(pb.y, fe_dbl(Fq::one()));
    if (!xyzz_is_inf(acc) && !t.is_zero()) {
      Fq t2 = fe_sqr(t), t3 = fe_mul(t2, t);
      acc.x = fe_mul(acc.x, t2); acc.y = fe_mul(acc.y, t3); acc.zz = t2; acc.zzz = t3;
    }
    if (!xyzz_is_inf(q) && !s.is_zero()) {
      Fq s2 = fe_sqr(s), s3 = fe_mul(s2, s);
      q.x = fe_mul(q.x, s2); q.y = fe_mul(q.y, s3); q.zz = s2; q.zzz = s3;
    }
    xyzz_add(acc, q);
  }
  store_point_bytes(out + i * 16, acc);
}

}  // namespace

extern "C" int myzkp_test_field_op(myzkp_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out,
                                   size_t n) {
  if (!ctx || !a || !out || op < 0 || op > 5 || ((op <= 2) && !b)) return MYZKP_ERR_INVALID_ARG;
  if (n == 0) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32 * 3));
  uint32_t* da = ctx->scalars.as<uint32_t>();
  uint32_t* db = da + n * 8;
  uint32_t* dout = db + n * 8;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (b) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (field == 0) test_field_kernel<Fq><<<blocks, 128, 0, ctx->stream>>>(op, da, b ? db : nullptr, dout, n);
  else test_field_kernel<Fr><<<blocks, 128, 0, ctx->stream>>>(op, da, b ? db : nullptr, dout, n);
  MZ_LAUNCH_CHECK(ctx);
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MYZKP_OK;
}

extern "C" int myzkp_test_g1_op(myzkp_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (!ctx || !a || !out || op < 0 || op > 3 || (op != 1 && !b)) return MYZKP_ERR_INVALID_ARG;
  if (n == 0) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 64 * 3));
  uint32_t* da = ctx->scalars.as<uint32_t>();
  uint32_t* db = da + n * 16;
  uint32_t* dout = db + n * 16;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(da, a, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  if (b) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(db, b, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  test_g1_kernel<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(op, da, b ? db : nullptr, dout, n);
  MZ_LAUNCH_CHECK(ctx);
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out, dout, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MYZKP_OK;
}

extern "C" int myzkp_test_set_sort_group_cap(int cap) {
  mz::sort_set_group_cap(cap);
  return MYZKP_OK;
}

// INT32 multiply-add (IMAD) pipe peak on this GPU, measured the way the MSM uses it:
//   mad.lo.u32 / mad.hi.u32 / mad.wide.u32 streams with 8-way ILP per thread, all SMs
//   full, plus the library's own Montgomery multiply (fe_mul) as a dependent chain.
// Prints one JSON object.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdio.h>

#include "../myzkp_b200/csrc/field.cuh"

#define ILP 8
template <int MODE>
__global__ void __launch_bounds__(256) imad_kernel(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a[ILP], b = seed | 1u, c = seed * 3u + 7u;
  uint64_t w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x + i * 77u + seed; w[i] = a[i]; }
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (MODE == 3) {
          uint32_t hi = (uint32_t)(w[i] >> 32);
          asm volatile("{.reg .u32 t; mov.u32 t, %0; mad.lo.cc.u32 %0, t, %2, %0; madc.hi.u32 %1, t, %2, %1;}"
                       : "+r"(a[i]), "+r"(hi) : "r"(b));
          w[i] = ((uint64_t)hi << 32) | (uint32_t)w[i];
        }
        if (MODE == 2) {
          uint32_t lo = (uint32_t)w[i];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(lo), "r"(b));
        }
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// other pipes, for planning: FP64 FMA (52-bit-limb multiplication candidates), integer adds, and
// DFMA + mad.wide issued together (do the FP64 and IMAD pipes overlap?)
//   MODE 0: fma.rn.f64   MODE 1: add.u32 (IADD3)   MODE 2: one fma.rn.f64 + one mad.wide.u32 per slot
template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t* out, uint32_t seed, int iters) {
  double d[ILP], e = 1.0000001 + seed * 1e-12, f = 1e-9;
  uint32_t a[ILP], b = seed | 1u;
  uint64_t w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { d[i] = 1.0 + threadIdx.x * 1e-6 + i; a[i] = threadIdx.x + i + seed; w[i] = a[i]; }
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (MODE == 0 || MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
        if (MODE == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        if (MODE == 2) {
          uint32_t lo = (uint32_t)w[i];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(lo), "r"(b));
        }
      }
    }
  }
  double sd = 0;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) { sd += d[i]; s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)__double2ll_rz(sd);
}

// dependent chain of Montgomery squarings
__global__ void __launch_bounds__(128) fesqr_kernel(uint32_t* out, uint32_t seed, int iters) {
  mz::Fq x;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = seed * (i + 3) + threadIdx.x;
  x.v[7] &= 0x0fffffff;
  for (int k = 0; k < iters; k++) x = mz::fe_sqr(x);
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= x.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chain(s) of Montgomery multiplies: CH independent chains per thread
template <int CH>
__global__ void __launch_bounds__(128) femul_kernel(uint32_t* out, uint32_t seed, int iters) {
  mz::Fq x[CH], y;
#pragma unroll
  for (int i = 0; i < 8; i++) y.v[i] = seed * (i + 3) + threadIdx.x;
  y.v[7] &= 0x0fffffff;
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = y; x[c].v[0] += c; }
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = mz::fe_mul(x[c], y);
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++)
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[c].v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  uint32_t* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"attr_clock_mhz\": %.0f", p.name, sms, clk_khz / 1000.0);
  const int iters = 4096;
  const char* names[4] = {"mad_lo", "mad_hi", "mad_wide", "mad_lohi_cc_pair"};
  for (int mode = 0; mode < 4; mode++) {
    for (int bps = 2; bps <= 8; bps *= 2) {  // 256-thread blocks per SM: 512, 1024, 2048 threads
      int blocks = sms * bps;
      float ms = 0;
      if (mode == 0) ms = time_ms([&] { imad_kernel<0><<<blocks, 256>>>(out, 12345u, iters); }, 5);
      if (mode == 1) ms = time_ms([&] { imad_kernel<1><<<blocks, 256>>>(out, 12345u, iters); }, 5);
      if (mode == 2) ms = time_ms([&] { imad_kernel<2><<<blocks, 256>>>(out, 12345u, iters); }, 5);
      if (mode == 3) ms = time_ms([&] { imad_kernel<3><<<blocks, 256>>>(out, 12345u, iters); }, 5);
      double ops = (double)blocks * 256 * iters * 8.0 * ILP;
      printf(", \"%s_thr%d_Tops\": %.3f", names[mode], bps * 256, ops / (ms * 1e-3) / 1e12);
    }
  }
  for (int bps = 1; bps <= 4; bps *= 2) {  // 128-thread blocks: 4 per SM = the accumulate kernel's occupancy
    int blocks = sms * bps * 4;
    const int it = 2000;
    float ms1 = time_ms([&] { femul_kernel<1><<<blocks, 128>>>(out, 999u, it); }, 5);
    float ms2 = time_ms([&] { femul_kernel<2><<<blocks, 128>>>(out, 999u, it); }, 5);
    printf(", \"femul_ch1_thr%d_Gmul\": %.2f", bps * 512, (double)blocks * 128 * it / (ms1 * 1e-3) / 1e9);
    printf(", \"femul_ch2_thr%d_Gmul\": %.2f", bps * 512, (double)blocks * 128 * it * 2 / (ms2 * 1e-3) / 1e9);
  }
  {
    const char* pn[3] = {"dfma", "iadd", "dfma_plus_mad_wide_pairs"};
    for (int mode = 0; mode < 3; mode++) {
      int blocks = sms * 8;
      float ms = 0;
      if (mode == 0) ms = time_ms([&] { pipe_kernel<0><<<blocks, 256>>>(out, 12345u, iters); }, 5);
      if (mode == 1) ms = time_ms([&] { pipe_kernel<1><<<blocks, 256>>>(out, 12345u, iters); }, 5);
      if (mode == 2) ms = time_ms([&] { pipe_kernel<2><<<blocks, 256>>>(out, 12345u, iters); }, 5);
      double ops = (double)blocks * 256 * iters * 8.0 * ILP;
      printf(", \"%s_thr2048_Tops\": %.3f", pn[mode], ops / (ms * 1e-3) / 1e12);
    }
    int blocks = sms * 16;
    const int it = 2000;
    float ms = time_ms([&] { fesqr_kernel<<<blocks, 128>>>(out, 999u, it); }, 5);
    printf(", \"fesqr_ch1_thr2048_Gsqr\": %.2f", (double)blocks * 128 * it / (ms * 1e-3) / 1e9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf(", \"status\": \"%s\"}\n", cudaGetErrorString(e));
  return 0;
}

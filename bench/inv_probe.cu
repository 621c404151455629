// Planning probe for a fused batched-affine accumulate (profiles/baa_r1.md): what does a warp that INVERTS cost the
// warps that MULTIPLY on the same SM sub-partition?
//
// One block of 512 threads per SM (16 warps; warp w runs on sub-partition w % 4).  Warps 4..15 ("multipliers", three
// per sub-partition) each run a dependent chain of `muls` Montgomery multiplications - the accumulate kernel's diet.
// Warps 0..3 ("inverters", one per sub-partition) do, depending on the mode,
//   0: nothing (baseline),
//   1: `invs` binary-GCD inversions of 32 DIFFERENT values per warp (lanes diverge),
//   2: `invs` binary-GCD inversions of ONE value shared by the warp (no divergence),
//   3: `invs` Fermat inversions (all on the multiply pipe).
// Reported: time of the whole block (the multipliers' chains are sized to dominate), i.e. how much the inverter's
// instruction stream slows the multiply-bound warps, and the inverter's own duration via clock64.
// Prints one JSON object.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o inv_probe inv_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

#include "../myzkp_b200/csrc/field.cuh"

using namespace mz;

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(uint32_t* out, unsigned long long* inv_cycles, uint32_t seed, int muls, int invs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Fq a, b;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a.v[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
    b.v[i] = (blockIdx.x * 97u + i * 7919u + seed * 3u + 1u) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
  }
  uint32_t acc = 0;
  if (warp >= 4) {
    for (int k = 0; k < muls; k++) a = fe_mul(a, b);
  } else if (MODE != 0) {
    if (MODE == 2) {  // one value for the whole warp
#pragma unroll
      for (int i = 0; i < 8; i++) a.v[i] = __shfl_sync(0xffffffffu, a.v[i], 0);
    }
    const long long t0 = clock64();
    for (int k = 0; k < invs; k++) {
      a = (MODE == 3) ? fe_inv(a) : fe_inv_bingcd(a);
      a.v[0] ^= (uint32_t)k + 1u;  // next input depends on this output
      a.v[7] &= 0x0fffffffu;
    }
    if (lane == 0 && blockIdx.x == 0) inv_cycles[warp] = (unsigned long long)(clock64() - t0);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) acc += a.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static float run(uint32_t* d_out, unsigned long long* d_cyc, int sms, int muls, int invs) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<sms, 512>>>(d_out, d_cyc, 1u, muls / 8, invs ? 1 : 0);  // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<sms, 512>>>(d_out, d_cyc, 7u, muls, invs);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main(int argc, char** argv) {
  const int muls = argc > 1 ? atoi(argv[1]) : 20000;
  const int invs = argc > 2 ? atoi(argv[2]) : 40;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
    printf("{\"error\": \"no CUDA device\"}\n");
    return 1;
  }
  const int sms = prop.multiProcessorCount;
  uint32_t* d_out;
  unsigned long long* d_cyc;
  cudaMalloc(&d_out, (size_t)sms * 512 * 4);
  cudaMalloc(&d_cyc, 4 * sizeof(unsigned long long));
  cudaMemset(d_cyc, 0, 4 * sizeof(unsigned long long));
  unsigned long long cyc[4][4] = {};
  float ms[4];
  ms[0] = run<0>(d_out, d_cyc, sms, muls, 0);
  ms[1] = run<1>(d_out, d_cyc, sms, muls, invs);
  cudaMemcpy(cyc[1], d_cyc, sizeof(cyc[1]), cudaMemcpyDeviceToHost);
  ms[2] = run<2>(d_out, d_cyc, sms, muls, invs);
  cudaMemcpy(cyc[2], d_cyc, sizeof(cyc[2]), cudaMemcpyDeviceToHost);
  ms[3] = run<3>(d_out, d_cyc, sms, muls, invs);
  cudaMemcpy(cyc[3], d_cyc, sizeof(cyc[3]), cudaMemcpyDeviceToHost);
  const double mul_total = (double)sms * 12 * 32 * muls;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"muls_per_chain\": %d, \"inversions_per_inverter_warp\": %d,\n", prop.name, sms, muls, invs);
  printf(" \"multipliers_only_ms\": %.4f, \"mul_per_s\": %.4g,\n", ms[0], mul_total / (ms[0] * 1e-3));
  const char* names[4] = {"", "gcd_32_values", "gcd_shared_value", "fermat"};
  for (int m = 1; m <= 3; m++)
    printf(" \"with_%s_ms\": %.4f, \"slowdown_%s\": %.4f, \"cycles_per_inversion_%s\": %.0f%s\n", names[m], ms[m], names[m],
           ms[m] / ms[0], names[m], invs ? (double)cyc[m][0] / invs : 0.0, m == 3 ? "}" : ",");
  return 0;
}

// Microbenchmark: fp64 FMA issue rate per SM as a function of warps per scheduler and independent chains per thread
// (answers: how much ILP x TLP the AO sweep needs to keep the B200 fp64 pipe busy).  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fp64_chains tools/ubench/fp64_chains.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_chain(double* out, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  if (s == 12345.678) out[0] = s;
}

template <int CH>
void run(int warps, int iters) {
  double* d;
  cudaMalloc(&d, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_chain<CH><<<148, warps * 32>>>(d, 10, 0.999, 1e-3);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_chain<CH><<<148, warps * 32>>>(d, iters, 0.999, 1e-3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  const double fma_per_sm = (double)warps * 32 * iters * 8 * CH;
  printf("warps/SM %2d (per scheduler %d) chains %d : %.2f DFMA/clk/SM, %.1f cycles per dependent DFMA step per warp\n", warps, warps / 4, CH,
         fma_per_sm / cycles, cycles / (iters * 8.0));
  cudaFree(d);
}

int main() {
  const int it = 20000;
  for (int w : {4, 8, 16, 32}) {
    run<1>(w, it);
    run<2>(w, it);
    run<4>(w, it);
    run<8>(w, it);
  }
  return 0;
}

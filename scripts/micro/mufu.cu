// Microbenchmark: MUFU.EX2 issue interval per warp / per SM sub-partition on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* clk) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = -0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 4 << 20); cudaMallocManaged(&clk, 8);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    k<<<148, warps * 32>>>(out, iters, clk); cudaDeviceSynchronize();
    k<<<148, warps * 32>>>(out, iters, clk); cudaDeviceSynchronize();
    double per_warp_instr = (double)*clk / (iters * 16.0);
    int per_smsp = (warps + 3) / 4;
    printf("warps/CTA %2d (per SMSP %d): %.2f clk per MUFU warp-instruction per warp -> %.2f clk per instr per SMSP, lanes/clk/SM = %.1f\n",
           warps, per_smsp, per_warp_instr, per_warp_instr / per_smsp, warps * 32 / per_warp_instr);
  }
  return 0;
}

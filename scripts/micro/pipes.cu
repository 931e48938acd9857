// Microbenchmark: per-warp issue cost of the instructions in the softmax exp pass on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float a[32];
  uint32_t pk[16];
  float s0 = 0.f, s1 = 0.f, m = 0.5f;
  for (int i = 0; i < 32; ++i) a[i] = -0.001f * (threadIdx.x + i);
  for (int i = 0; i < 16; ++i) pk[i] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x = a[2 * j] - m, y = a[2 * j + 1] - m;
      if (MODE == 0 || MODE == 3) {  // MUFU
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y));
      }
      if (MODE == 1 || MODE == 3) {  // pack
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
        pk[j] ^= r;
      }
      if (MODE == 2 || MODE == 3) {  // sums
        s0 += x;
        s1 += y;
      }
      a[2 * j] = x * 0.999f;
      a[2 * j + 1] = y * 0.999f;
    }
  }
  long long t1 = clock64();
  float s = s0 + s1;
  for (int i = 0; i < 32; ++i) s += a[i];
  for (int i = 0; i < 16; ++i) s += __uint_as_float(pk[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 4 << 20); cudaMallocManaged(&clk, 8);
  const int iters = 2000;
  const char* names[4] = {"sub+ex2+mul", "sub+pack+mul", "sub+sum+mul", "sub+ex2+pack+sum+mul (the exp pass)"};
  for (int warps : {4, 8}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, clk);
        cudaDeviceSynchronize();
      }
      printf("%d warps/SMSP  %-40s %.1f clk per 32 elements per warp (128-key block: %.0f clk)\n", warps / 4, names[mode],
             (double)*clk / iters, (double)*clk / iters * 4);
    }
  }
  return 0;
}

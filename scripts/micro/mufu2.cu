// Microbenchmark: packed half-precision exp2 on the MUFU (ex2.approx.f16x2 / .ftz.bf16x2) vs fp32.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  uint32_t a[16];
  float f[16];
  for (int i = 0; i < 16; ++i) { a[i] = 0xB800B400u + threadIdx.x + i; f[i] = -0.001f * (threadIdx.x + i); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += f[i] + __uint_as_float(a[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 4 << 20); cudaMallocManaged(&clk, 8);
  const int iters = 2000;
  const char* names[3] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 256>>>(out, iters, clk);
      if (mode == 1) k<1><<<148, 256>>>(out, iters, clk);
      if (mode == 2) k<2><<<148, 256>>>(out, iters, clk);
      cudaDeviceSynchronize();
    }
    double per = (double)*clk / (iters * 16.0) / 2.0;  // 2 warps per SMSP
    printf("%-24s %.2f clk per warp instruction per SMSP -> %.1f exps/clk/SM\n", names[mode], per,
           (mode == 0 ? 32.0 : 64.0) / per * 4);
  }
  return 0;
}

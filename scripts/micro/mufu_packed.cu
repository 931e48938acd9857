// Microbenchmark: issue interval of packed MUFU forms on sm_100a — ex2.approx.ftz.bf16x2 / f16x2, tanh.approx.bf16x2 /
// f16x2 — against the scalar f32 forms.  If a packed instruction occupies the unit as long as a scalar one, it
// delivers two results per slot.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_packed mufu_packed.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int iters, long long* clk) {
  unsigned a[16];
  float f[16];
  for (int i = 0; i < 16; ++i) { a[i] = 0xBC00BC00u + threadIdx.x + i; f[i] = -0.001f * (threadIdx.x + i); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 5) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
  for (int i = 0; i < 16; ++i) s += a[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE>
void run(const char* name, unsigned* out, long long* clk) {
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    k<MODE><<<148, warps * 32>>>(out, iters, clk); cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32>>>(out, iters, clk); cudaDeviceSynchronize();
    double per = (double)*clk / (iters * 16.0);
    printf("%-24s warps/SM %2d: %.2f clk per warp-instruction per warp -> %.2f clk per instruction per sub-partition\n", name, warps, per,
           per / ((warps + 3) / 4));
  }
}
int main() {
  unsigned* out; long long* clk; cudaMalloc(&out, 4 << 20); cudaMallocManaged(&clk, 8);
  run<0>("ex2.approx.ftz.f32", out, clk);
  run<1>("ex2.approx.ftz.bf16x2", out, clk);
  run<2>("ex2.approx.f16x2", out, clk);
  run<3>("tanh.approx.f32", out, clk);
  run<4>("tanh.approx.bf16x2", out, clk);
  run<5>("tanh.approx.f16x2", out, clk);
  return 0;
}

// Microbenchmark: which pipe does the bf16 pack (cvt.rn.bf16x2.f32 -> F2FP.BF16.F32.PACK_AB) use on sm_100a?  If it shares
// the XU with MUFU.EX2, the softmax exp pass (2 MUFU + 1 pack per score pair) is bound by their SUM.  Also times an
// integer-only pack (round-half-up: two IADD + one PRMT) as the alternative.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>  // 0: MUFU only, 1: F2FP only, 2: MUFU + F2FP, 3: integer pack only, 4: MUFU + integer pack
__global__ void k(uint32_t* out, int iters, long long* clk) {
  float f[16]; float g[16]; uint32_t acc[8];
  for (int i = 0; i < 16; ++i) { f[i] = -0.001f * (threadIdx.x + i); g[i] = 1.0f + 0.01f * i + threadIdx.x; }
  for (int i = 0; i < 8; ++i) acc[i] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0 || MODE == 2 || MODE == 4) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i + 1]));
      }
      if (MODE == 1 || MODE == 2) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(g[i + 1]), "f"(g[i]));
        acc[i >> 1] += r;
      }
      if (MODE == 3 || MODE == 4) {
        uint32_t a = __float_as_uint(g[i]) + 0x8000u, b = __float_as_uint(g[i + 1]) + 0x8000u, r;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
        acc[i >> 1] += r;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] += 1.0f;
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  for (int i = 0; i < 16; ++i) s += __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  uint32_t* out; long long* clk; cudaMalloc(&out, 4 << 20); cudaMallocManaged(&clk, 8);
  const int iters = 2000;
  const char* names[5] = {"16 MUFU.EX2", "8 F2FP pack", "16 MUFU.EX2 + 8 F2FP pack", "8 integer pack (2 IADD + PRMT)", "16 MUFU.EX2 + 8 integer pack"};
  for (int warps : {4, 8}) {
    for (int mode = 0; mode < 5; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, clk);
        if (mode == 4) k<4><<<148, warps * 32>>>(out, iters, clk);
        cudaDeviceSynchronize();
      }
      printf("%d warp(s)/SMSP  %-34s %.1f clk per iteration per warp (+16 FADD)\n", warps / 4, names[mode], (double)*clk / iters);
    }
  }
  return 0;
}

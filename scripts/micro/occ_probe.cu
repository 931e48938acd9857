// Probe: what keeps a 384-thread, 80-register kernel at ONE resident CTA per SM on sm_100a — setmaxnreg (register
// reconfiguration) or the use of tensor memory?  Prints cudaOccupancyMaxActiveBlocksPerMultiprocessor for four variants.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <bool REGCFG, bool TMEM>
__global__ void __launch_bounds__(384, 2) k(uint32_t* out) {
  __shared__ uint32_t holder;
  uint32_t v = threadIdx.x;
  if (TMEM) {
    if (threadIdx.x < 32) {
      uint32_t a = (uint32_t)__cvta_generic_to_shared(&holder);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(a) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    v += holder;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(holder) : "memory");
  }
  if (REGCFG) {
    if (threadIdx.x < 128) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
template <bool R, bool T> void probe(const char* name) {
  int nb = 0;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k<R, T>);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<R, T>, 384, 0);
  printf("%-28s regs %3d -> %d resident CTA(s) per SM (%s)\n", name, fa.numRegs, nb, cudaGetErrorString(e));
}
int main() {
  probe<false, false>("plain");
  probe<true, false>("setmaxnreg");
  probe<false, true>("tcgen05.alloc");
  probe<true, true>("setmaxnreg + tcgen05.alloc");
  return 0;
}

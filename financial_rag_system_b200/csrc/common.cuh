// common.cuh — sm_100a PTX wrappers (mbarrier, TMA, tcgen05/TMEM) and small helpers
// shared by the scan and encoder kernels.  Hand-written; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

namespace frs {

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident (and run its prologue: barrier init, TMEM allocation, descriptor prefetch, weight staging) while the previous
// kernel of the stream is still draining.  pdl_wait() blocks until that previous grid has COMPLETED and its memory is
// visible; everything that reads the previous kernel's output or overwrites its inputs comes after it.
// pdl_launch_dependents() lets the next kernel of the stream start becoming resident as SMs free up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Monotone map float -> uint32 (larger float => larger key; -0 < +0; NaNs sort at the ends).
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_ordered(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (launch error) rather than hang the GPU.
#ifndef FRS_MBAR_SPIN_LIMIT
#define FRS_MBAR_SPIN_LIMIT (1u << 26)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > FRS_MBAR_SPIN_LIMIT) __trap();  // no printf: it would give the kernel a stack frame
  }
}

// ---------------------------------------------------------------------------------------------
// thread-block clusters: rank, peer shared-memory addresses, cluster-scope mbarrier traffic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
// address of the same shared-memory location in CTA `rank` of this cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// arrive on a peer CTA's mbarrier WITHOUT a cluster-scope release of this thread's memory operations: for signals
// whose payload is not generic memory (a drained TMEM accumulator, ordered by tcgen05.wait::ld + tcgen05.fence).
// (The .release.cluster form above costs ~1 us per call: measured 1056 ns against 128 ns for a local arrive.)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// named barriers (sub-CTA sync among the epilogue warps)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// producer side of a named barrier: counts this thread in and does not wait
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// barrier + OR-reduction of a predicate over the participating threads
__device__ __forceinline__ bool named_bar_or(uint32_t id, uint32_t nthreads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %3, 0;\n\t"
      "barrier.cta.red.or.pred q, %1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"(id), "r"(nthreads), "r"((uint32_t)pred)
      : "memory");
  return r != 0;
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// CTA-pair load: the box lands in THIS CTA's shared memory, the bytes complete on an mbarrier that may live in
// the peer (leader) CTA — bar_cluster_addr is a shared::cluster address (mapa)
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// shared -> global tile store (bulk async group); the issuing thread owns the group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N groups have finished READING shared memory (the staging buffer may be rewritten)
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one warp of EACH CTA of the pair allocates / frees; the MMA is issued by the
// leader CTA (rank 0) for both: M = 256 (128 rows per CTA), each CTA holds its own A rows and HALF of the B rows at
// the same shared-memory offsets, and its own half of the accumulator at the same TMEM columns.
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// commit all prior tcgen05.mma of this thread; arrives (count 1) on the mbarrier when they finish
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor, K-major operand tile in the canonical SWIZZLE_128B layout:
// rows are 128 B wide, 8-row groups are 1024 B apart (SBO), the tile is 1024 B aligned.
//   bits [0,14)  start address >> 4      bits [16,30) LBO >> 4 (=1, unused for swizzled K-major)
//   bits [32,46) SBO >> 4 (=64)          bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate, both operands K-major.
//   fmt: 1 = bf16, 0 = f16, 2 = tf32
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4)            // c_format = F32
         | (fmt << 7)         // a_format
         | (fmt << 10)        // b_format
         | (0u << 15)         // a_major = K
         | (0u << 16)         // b_major = K
         | ((N >> 3) << 17)   // n_dim
         | ((M >> 4) << 24);  // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread for the CTA
template <bool kTF32>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                       uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// 1 in exactly one lane of the (converged) warp, chosen by the hardware: elect.sync
__device__ __forceinline__ uint32_t elect_one_pred() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}
// Predicated forms for an issuer WARP that stays converged: every lane runs the (uniform) control flow and
// address arithmetic — so the compiler keeps descriptors in uniform registers instead of the
// ELECT / R2UR.BROADCAST loops it emits inside an `if (lane == 0)` region (~70 clk per MMA) — and only the
// lane with issue != 0 executes the instruction.
__device__ __forceinline__ void tc_mma_f16_pred(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand is read from tensor memory (lane = row, two bf16 K-values
// per 32-bit column), so a matrix produced by the threads (softmax probabilities) never touches shared memory
__device__ __forceinline__ void tc_mma_ts_pred(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pred(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(issue)
      : "memory");
}
// packed fp32 pair add (one FADD2 on sm_100): (a, b) += (c, d)
__device__ __forceinline__ void fadd2(float& a, float& b, float c, float d) {
  asm("{\n\t.reg .b64 x, y;\n\t"
      "mov.b64 x, {%0, %1};\n\t"
      "mov.b64 y, {%2, %3};\n\t"
      "add.f32x2 x, x, y;\n\t"
      "mov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a), "+f"(b)
      : "f"(c), "f"(d));
}
// packed fp32 pairs: one FMUL2 / FFMA2 issue slot per two elements (sm_100 f32x2 arithmetic)
__device__ __forceinline__ void fmul2(float& a, float& b, float c, float d) {
  asm("{\n\t.reg .b64 x, y;\n\t"
      "mov.b64 x, {%0, %1};\n\t"
      "mov.b64 y, {%2, %3};\n\t"
      "mul.f32x2 x, x, y;\n\t"
      "mov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a), "+f"(b)
      : "f"(c), "f"(d));
}
// (a, b) = (a, b) * (c, d) + (e, f)
__device__ __forceinline__ void ffma2(float& a, float& b, float c, float d, float e, float f) {
  asm("{\n\t.reg .b64 x, y, z;\n\t"
      "mov.b64 x, {%0, %1};\n\t"
      "mov.b64 y, {%2, %3};\n\t"
      "mov.b64 z, {%4, %5};\n\t"
      "fma.rn.f32x2 x, x, y, z;\n\t"
      "mov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a), "+f"(b)
      : "f"(c), "f"(d), "f"(e), "f"(f));
}
// (acc0, acc1) += (a, b) * (c, d)
__device__ __forceinline__ void ffma2_acc(float& acc0, float& acc1, float a, float b, float c, float d) {
  asm("{\n\t.reg .b64 x, y, z;\n\t"
      "mov.b64 x, {%2, %3};\n\t"
      "mov.b64 y, {%4, %5};\n\t"
      "mov.b64 z, {%0, %1};\n\t"
      "fma.rn.f32x2 z, x, y, z;\n\t"
      "mov.b64 {%0, %1}, z;\n\t}"
      : "+f"(acc0), "+f"(acc1)
      : "f"(a), "f"(b), "f"(c), "f"(d));
}
// 2^a, 2^b on the FMA pipes (no MUFU): x is clamped to >= -125, split by the magic-number addition into
// n = round(x) (low mantissa bits of r = x + 1.5 * 2^23) and f = x - n in [-0.5, 0.5], 2^f is a cubic (minimax,
// relative error 7.5e-5 — the result is rounded to bf16 anyway) and n is added to the exponent field.
// Per PAIR: 2 FMNMX + 3 FADD2/FFMA2 + 3 FFMA2 + 2 integer ops = 10 issue slots; two MUFU.EX2 hold their own
// warp for 16 clk.  Valid for x <= 120.
__device__ __forceinline__ void exp2_pair_fma(float& a, float& b) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
  a = fmaxf(a, -125.0f);
  b = fmaxf(b, -125.0f);
  float ra = a, rb = b;
  fadd2(ra, rb, kMagic, kMagic);
  float na = ra, nb = rb;
  fadd2(na, nb, -kMagic, -kMagic);
  ffma2(na, nb, -1.0f, -1.0f, a, b);  // f = x - n
  float pa = 0.0551716685295105f, pb = 0.0551716685295105f;
  ffma2(pa, pb, na, nb, 0.2426111251115799f, 0.2426111251115799f);
  ffma2(pa, pb, na, nb, 0.6932609677314758f, 0.6932609677314758f);
  ffma2(pa, pb, na, nb, 0.9999280571937561f, 0.9999280571937561f);
  a = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ra) << 23));
  b = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(rb) << 23));
}
__device__ __forceinline__ void tc_mma2_f16_pred(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
// commit of the pair's MMAs: one arrival on the mbarrier at this offset of every CTA in cta_mask
__device__ __forceinline__ void tc_commit2_mc_pred(uint64_t* bar, uint16_t cta_mask, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask), "r"(issue)
      : "memory");
}
// three-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// 32 lanes x 32 columns of fp32: thread i of the warp receives columns [col, col+32) of TMEM lane
// (lane_base + i); lane_base must be 32 * (warp_id % 4).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 columns
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr)
      : "memory");
}
// N columns (8 or 16) selected at compile time
template <int N>
__device__ __forceinline__ void tmem_ld_32xN(uint32_t taddr, uint32_t (&v)[N]) {
  static_assert(N == 8 || N == 16 || N == 32, "unsupported TMEM load width");
  if constexpr (N == 8) tmem_ld_32x8(taddr, v);
  else if constexpr (N == 16) tmem_ld_32x16(taddr, v);
  else tmem_ld_32x32(taddr, v);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// registers -> TMEM, 32 lanes x 32 columns (same addressing as tmem_ld_32x32)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
      "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
      "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// SFU approximations used by the encoder epilogues
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two fp32 -> packed bf16x2 (lo = a, hi = b), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// Function attributes (the opt-in dynamic shared memory limit) live in a device's context: a launcher configures its
// kernel once per (call site, device).  `done` is the call site's static table; a second thread that races the first
// configures the kernel twice, which is harmless.
constexpr int kMaxDevices = 64;
struct DeviceOnce {
  std::atomic<unsigned char> done[kMaxDevices];
};
template <typename F>
inline cudaError_t once_per_device(DeviceOnce& o, F&& configure) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  if (tracked && o.done[dev].load(std::memory_order_acquire)) return cudaSuccess;
  e = configure();
  if (e == cudaSuccess && tracked) o.done[dev].store(1, std::memory_order_release);
  return e;
}

}  // namespace frs

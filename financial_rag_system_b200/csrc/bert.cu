// bert.cu — sm_100a kernels of the two BERT encoders (bge-small-en-v1.5, ms-marco-MiniLM-L-6-v2).
//
// What they replace, in transformers/models/bert/modeling_bert.py (the arithmetic that
// SentenceTransformer.encode / CrossEncoder.predict run for reference main.py:148,213,245 and
// main2.py:166,171):
//   embed_ln_kernel        BertEmbeddings.forward            :102-112
//   gemm_kernel<QKV>       BertSelfAttention q/k/v Linear    :179-181   (one fused [1152,384] GEMM)
//   attention_kernel       softmax(QK^T/sqrt(32) + mask) V   :115-140, :192-205  (varlen, no padding)
//   gemm_kernel<ResLN>     BertSelfOutput / BertOutput       :294-298, :352-356  (bias+residual+LayerNorm)
//   gemm_kernel<Gelu>      BertIntermediate (erf GELU)       :339-342
//   pool_normalize_kernel  sentence-transformers Pooling(cls|mean) + Normalize
//   (BertPooler + classifier :462-468, :1111-1124 run in fp32 from the fp32 [CLS] rows: bert_fp32.cu ce_head_f32_kernel)
//
// Every GEMM is tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation in TMEM) fed by TMA into
// SWIZZLE_128B shared-memory slabs; warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM
// allocation; the warp runs converged and the lane elected by elect.sync issues), warps 2..9 epilogue /
// softmax (a thread owns one token row = one TMEM lane), GEMM warps 10..11 TMA stores.
#include <math.h>
#include <stdlib.h>

#ifndef FRS_BG_SLEEP_NS
#define FRS_BG_SLEEP_NS 64
#endif
#include "bert.cuh"
#include "common.cuh"

namespace frs {

// ---------------------------------------------------------------------------------------------
// protocol-bug diagnostics: a barrier wait that times out records WHICH wait it was in mapped host
// memory (readable after the trap has killed the context) — see frs_debug_trap_info()
// ---------------------------------------------------------------------------------------------
__device__ uint32_t* g_trap_info = nullptr;  // [4]: code, blockIdx.x, parity, threadIdx.x
static uint32_t* g_trap_info_host = nullptr;

__device__ __noinline__ void trap_with_code(uint32_t code, uint32_t parity) {
  if (g_trap_info) {
    if (atomicCAS(g_trap_info, 0u, code) == 0u) {
      g_trap_info[1] = blockIdx.x;
      g_trap_info[2] = parity;
      g_trap_info[3] = threadIdx.x;
      __threadfence_system();
    }
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait_c(uint64_t* bar, uint32_t parity, uint32_t code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) trap_with_code(code, parity);
  }
}
// Same, for the single-thread producer / MMA-issuer roles: they share an SM sub-partition with two
// softmax (or epilogue) warps, and a tight polling loop would take issue slots from them.
__device__ __forceinline__ void mbar_wait_bg(uint64_t* bar, uint32_t parity, uint32_t code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (FRS_BG_SLEEP_NS > 0) __nanosleep(FRS_BG_SLEEP_NS);
    if (++spins > (1u << 22)) trap_with_code(code, parity);
  }
}

uint32_t* bert_trap_info_host() {
  if (!g_trap_info_host) {
    uint32_t* h = nullptr;
    if (cudaHostAlloc(&h, 16, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(h, 0, 16);
    uint32_t* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return nullptr;
    if (cudaMemcpyToSymbol(g_trap_info, &d, sizeof(d)) != cudaSuccess) return nullptr;
    g_trap_info_host = h;
  }
  return g_trap_info_host;
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}
// erf(x) as an odd rational function x P(x^2) / Q(x^2) on [-4, 4] (|erf| = 1 beyond in fp32): the
// single-precision fit used by Eigen/XLA, max abs error 6.5e-8 against scipy.special.erf — far below
// fp32 rounding of the result.  13 FMA + 1 reciprocal, no branches (erff() is ~2x the instructions,
// and this epilogue is issue-bound).
__device__ __forceinline__ float erf_rational(float x) {
  x = fminf(fmaxf(x, -4.0f), 4.0f);
  const float x2 = x * x;
  float p = fmaf(x2, -2.72614225801306e-10f, 2.77068142495902e-08f);
  p = fmaf(x2, p, -2.10102402082508e-06f);
  p = fmaf(x2, p, -5.69250639462346e-05f);
  p = fmaf(x2, p, -7.34990630326855e-04f);
  p = fmaf(x2, p, -2.95459980854025e-03f);
  p = fmaf(x2, p, -1.60960333262415e-02f);
  float q = fmaf(x2, -1.45660718464996e-05f, -2.13374055278905e-04f);
  q = fmaf(x2, q, -1.68282697438203e-03f);
  q = fmaf(x2, q, -7.37332916720468e-03f);
  q = fmaf(x2, q, -1.42647390514189e-02f);
  return __fdividef(x * p, q);
}
// GELU(x) = x Phi(x)  (modeling_bert.py:339-342, hidden_act="gelu" = the erf form).
// Default: Phi(x) = (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) / 2 with the inner polynomial fitted to the
// exact erf form (max |error| 2.5e-5 over the whole real line, scripts/fit_gelu.py) and MUFU tanh
// (relative error 2^-11): at most ~|x| 2.7e-4 in total, an eighth of the bf16 rounding step of the
// stored result — 8 instructions per element instead of 20.  -DFRS_EXACT_GELU selects the erf form.
__device__ __forceinline__ float gelu_erf(float x) {
  const float hx = 0.5f * x;
#ifdef FRS_EXACT_GELU
  return fmaf(hx, erf_rational(x * 0.70710678118654752f), hx);
#else
  const float x2 = x * x;
  float u = fmaf(x2, -0.00035151678934123415f, 0.03700564602521096f);
  u = fmaf(x2, u, 0.7975078842799184f);
  return fmaf(hx, tanh_approx(x * u), hx);
#endif
}
// two elements at once on the packed fp32 pipes (FMUL2 / FFMA2): 7 packed ops + 2 MUFU per pair instead of
// 16 + 2 — the GELU epilogue is issue-bound (bit-identical to gelu_erf: same operations, same rounding)
__device__ __forceinline__ void gelu_erf2(float& a, float& b) {
#ifdef FRS_EXACT_GELU
  a = gelu_erf(a);
  b = gelu_erf(b);
#else
  float x2a = a, x2b = b;
  fmul2(x2a, x2b, a, b);                                                  // x^2
  float ua = x2a, ub = x2b;
  ffma2(ua, ub, -0.00035151678934123415f, -0.00035151678934123415f, 0.03700564602521096f, 0.03700564602521096f);
  ffma2(ua, ub, x2a, x2b, 0.7975078842799184f, 0.7975078842799184f);      // c0 + x^2 (c1 + c2 x^2)
  fmul2(ua, ub, a, b);                                                    // x u
  ua = tanh_approx(ua);
  ub = tanh_approx(ub);
  fmul2(a, b, 0.5f, 0.5f);                                                // x / 2
  ffma2(ua, ub, a, b, a, b);                                              // (x/2) t + x/2
  a = ua;
  b = ub;
#endif
}
// GELU without the MUFU: z = sat(x / 8 + 1/2) - 1/2 clamps x to [-4, 4] in one FFMA.SAT, Phi(x) - 1/2 = z Q(z^2)
// with a degree-6 Q fitted to the erf form (scripts/fit_gelu.py --poly; max |error| of x Phi 1.9e-4 in fp32
// evaluation over the whole real line, against <= |x| 2.7e-4 for the tanh form).  12 issue slots per PAIR on the FMA
// pipes.  -DFRS_GELU_MIX alternates the two forms by column pair (with tanh alone the two epilogue warps of a
// sub-partition queue on the MUFU: 260 ns of a 448 ns chunk); it measured slower and is off by default.
__device__ __forceinline__ void gelu_poly2(float& a, float& b) {
  float za, zb;
  asm("fma.rn.sat.f32 %0, %1, 0f3E000000, 0f3F000000;" : "=f"(za) : "f"(a));  // sat(x * 0.125 + 0.5)
  asm("fma.rn.sat.f32 %0, %1, 0f3E000000, 0f3F000000;" : "=f"(zb) : "f"(b));
  fadd2(za, zb, -0.5f, -0.5f);
  float ua = za, ub = zb;
  fmul2(ua, ub, za, zb);  // z^2
  float qa = 12524.2783203125f, qb = 12524.2783203125f;
  ffma2(qa, qb, ua, ub, -13731.9296875f, -13731.9296875f);
  ffma2(qa, qb, ua, ub, 6436.49365234375f, 6436.49365234375f);
  ffma2(qa, qb, ua, ub, -1707.1162109375f, -1707.1162109375f);
  ffma2(qa, qb, ua, ub, 287.4537658691406f, 287.4537658691406f);
  ffma2(qa, qb, ua, ub, -33.061439514160156f, -33.061439514160156f);
  ffma2(qa, qb, ua, ub, 3.1830668449401855f, 3.1830668449401855f);
  ffma2(qa, qb, za, zb, 0.5f, 0.5f);  // Phi = z Q + 1/2
  fmul2(a, b, qa, qb);                // x Phi
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// positions + embeddings
// ---------------------------------------------------------------------------------------------
__global__ void row_map_kernel(const int32_t* __restrict__ cu, const int32_t* __restrict__ row_start, int n_seqs,
                               int32_t* __restrict__ src_tok, int32_t* __restrict__ pos_of_row,
                               int32_t* __restrict__ row_of_tok, int32_t* __restrict__ cls_slot) {
  const int s = blockIdx.x;
  if (s >= n_seqs) return;
  const int t0 = cu[s], len = cu[s + 1] - cu[s], r0 = row_start[s], slot = row_start[s + 1] - r0;
  for (int i = threadIdx.x; i < slot; i += blockDim.x) {
    src_tok[r0 + i] = i < len ? t0 + i : -1;
    pos_of_row[r0 + i] = i;
    cls_slot[r0 + i] = i == 0 ? s : -1;  // the [CLS] row of sequence s (see the ResLN epilogue)
    if (i < len) row_of_tok[t0 + i] = r0 + i;
  }
}

// one warp per internal row: 384 = 3 x (32 lanes x float4)
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ type_ids,
                const int32_t* __restrict__ src_tok, const int32_t* __restrict__ pos_of_row, int M, int vocab,
                const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                __nv_bfloat16* __restrict__ x) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int tok = src_tok[row];
  if (tok < 0) {  // alignment row between two sequences: keep it finite
    uint2* out = reinterpret_cast<uint2*>(x + (size_t)row * kHid);
#pragma unroll
    for (int i = 0; i < 3; ++i) out[lane + 32 * i] = make_uint2(0u, 0u);
    return;
  }
  int id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const int tt = type_ids ? (type_ids[tok] != 0) : 0;
  int ps = pos_of_row[row];
  ps = ps < 0 ? 0 : (ps >= kMaxSeq ? kMaxSeq - 1 : ps);
  const float4* w = reinterpret_cast<const float4*>(word + (size_t)id * kHid);
  const float4* pp = reinterpret_cast<const float4*>(pos + (size_t)ps * kHid);
  const float4* ty = reinterpret_cast<const float4*>(type + (size_t)tt * kHid);
  float v[12];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 a = __ldg(w + lane + 32 * i), b = __ldg(pp + lane + 32 * i), c = __ldg(ty + lane + 32 * i);
    v[4 * i + 0] = a.x + b.x + c.x;
    v[4 * i + 1] = a.y + b.y + c.y;
    v[4 * i + 2] = a.z + b.z + c.z;
    v[4 * i + 3] = a.w + b.w + c.w;
    sum += (v[4 * i] + v[4 * i + 1]) + (v[4 * i + 2] + v[4 * i + 3]);
  }
  const float mean = warp_sum(sum) * (1.0f / kHid);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float d = v[i] - mean;
    sq += d * d;
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / kHid) + eps);
  uint2* out = reinterpret_cast<uint2*>(x + (size_t)row * kHid);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    const float y0 = (v[4 * i + 0] - mean) * rstd * g.x + b.x;
    const float y1 = (v[4 * i + 1] - mean) * rstd * g.y + b.y;
    const float y2 = (v[4 * i + 2] - mean) * rstd * g.z + b.z;
    const float y3 = (v[4 * i + 3] - mean) * rstd * g.w + b.w;
    out[lane + 32 * i] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
  }
}

// ---------------------------------------------------------------------------------------------
// GEMM  C[M,N] = A[M,K] . W[N,K]^T  with fused epilogues
// ---------------------------------------------------------------------------------------------
constexpr int kGemmEpiWarps = 8;
constexpr int kGemmEpiThreads = 32 * kGemmEpiWarps;
constexpr int kGemmThreads = 64 + kGemmEpiThreads + 64;  // + one TMA-store warp per column-half group
constexpr int kNSub = 192;  // N of one tcgen05.mma / rows of one weight TMA box

// Every GEMM runs on CTA pairs with 2-SM MMAs; a CTA holds its own 128 A rows and HALF of the weight rows of a
// K-step.  LN = the ResLN kernels: the pair's tile is 256 rows x 384 columns (two N = 192 MMAs per K = 16), each
// CTA holds the whole 384-wide row of its 128 rows in ONE 384-column accumulator (LayerNorm needs the full row).
template <int BN, bool LN>
struct GemmCfg {
  static_assert(BN == 192, "tile N is 192");
  static constexpr int kStageA = kBM * 128;    // 128 rows x 64 bf16
  static constexpr int kStageB = (LN ? BN : BN / 2) * 128;  // this CTA's weight rows x 64 bf16: 2 x 96 (LN) or 96
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = LN ? 4 : 6;   // 40 KB / 28 KB stages
  static constexpr int kAcc = LN ? 1 : 2;      // TMEM accumulator stages (LN: 384 columns, the others 2 x 192)
  static constexpr int kTileN = LN ? 2 * BN : BN;        // columns of a CTA's tile
  static constexpr int kColsPerThread = kTileN / 2;      // two epilogue warps share a TMEM lane quarter
  static constexpr int kRing = kStages * kStage;
  static constexpr int kOut = kRing;                   // output staging for the TMA stores
  static constexpr int kOutBytes = kBM * kNSub * 2;    // 48 KB: 2 groups x 3 boxes of 128 rows x 32 bf16
  static constexpr int kParF = kOut + kOutBytes;       // fp32 params: bias[1536] | gamma[384] | beta[384]
  static constexpr int kStat = kParF + (1536 + 768) * 4;  // LN: per-row partial statistics of the two column halves
  static constexpr int kBars = kStat + (LN ? 2 * 2 * 128 * 16 : 0);  // [2 tile parities][2 halves][128 rows] float4 (sum, sumsq, shift, -)
  static constexpr int kHolder = kBars + (2 * kStages + 2 * kAcc) * 8;
  static constexpr int kTotal = kHolder + 16;
  static_assert(kTotal + 1024 <= 232448, "shared memory budget");
};

// Output path of every epilogue: registers -> swizzled staging box in shared memory -> TMA store.
// (Direct per-thread row stores cost one 16-byte LSU transaction per lane; a staged box leaves as full
// lines.)  The four warps that own one half of the tile's columns form a group with its own ring of three
// 8 KB boxes (one 32-column chunk of all 128 rows each) and its own store-issuing thread: a chunk is
// stored the moment it is staged, and its box is only reused three chunks later — the stores never sit
// on the epilogue's critical path.  Named barrier 2 + group.
constexpr int kChunkBox = kBM * 64;  // 128 rows x 32 bf16
constexpr uint32_t kBarStaged = 2;   // named barriers: staged(group, box) = 2 + 6 group + box, free = + 3
constexpr uint32_t kBarFree = 5;

// ResLN (bias + residual + LayerNorm over the 384-wide row): each CTA of the pair owns the WHOLE row of its 128
// rows (one 384-column accumulator: no double buffering, the epilogue of a tile is exposed, but a K-step now carries
// 768 clk of MMA per 40 KB stage instead of 384 and the main loop no longer waits for operands).  The residual
// rides through the tensor core as six extra K-steps against a 16 x 16 identity.
// -DFRS_GEMM_TRACE: event timeline (clock64 << 8 | id) of CTA 0 of one launch: role 0 = MMA issuer, role 1 = lane 0
// of the first epilogue warp, role 2 = TMA producer (dumped to gpurun_out/gemm_trace_<epi>.txt)
#ifdef FRS_GEMM_TRACE
constexpr int kGTraceCap = 4096;
__device__ __forceinline__ long long gt_now() {  // ns, comparable across SMs (the two CTAs of a pair)
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FRS_GT(id) do { if (gtr && gtn < kGTraceCap) gtr[gtn++] = (gt_now() << 8) | (long long)(id); } while (0)
#else
#define FRS_GT(id) do { } while (0)
#endif
template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
            const GemmParams p) {
  constexpr bool kLN = EPI == kEpiResLN;
  using C = GemmCfg<BN, kLN>;
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the pointer in the shared address space: rounding the
  // pointer through uintptr_t made every staging access a generic LD.E / ST.E instead of LDS / STS)
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = sm;
  uint8_t* sout0 = sm + C::kOut;
  float* sbias = reinterpret_cast<float*>(sm + C::kParF);
  float* sgamma = sbias + 1536;
  float* sbeta = sgamma + 384;
  float4* sstat = reinterpret_cast<float4*>(sm + C::kStat);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + C::kBars);
  uint64_t* empty = full + C::kStages;
  uint64_t* tfull = empty + C::kStages;
  uint64_t* tempty = tfull + C::kAcc;
  uint32_t* holder = reinterpret_cast<uint32_t*>(sm + C::kHolder);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  // CTA PAIRS with 2-SM MMAs (cta_group::2, M = 256): the pair owns two consecutive 128-row tiles; every CTA loads
  // its own A rows and HALF of the weight rows of a K-step into its own shared memory (the boxes complete on the
  // LEADER's mbarrier), and the leader CTA issues the MMAs for both.  A QKV / GELU stage is 28 KB instead of the
  // 40 KB of a 1-SM 128 x 192 tile (six stages in flight instead of four); a ResLN stage carries twice the MMA
  // work per byte.  The K-step time of the 1-SM kernels was the stage round trip (~2400 clk under load) divided
  // by the ring depth, not the tensor pipe.
  constexpr int kResSteps = kLN ? C::kTileN / 64 : 0;  // ResLN: K-steps that add the residual
  constexpr int kEyeOff = 2048;  // ResLN: this CTA's 8 rows of the 16 x 16 identity, byte offset inside the bias area
  const int nt_count = kLN ? 1 : p.N / BN;
  const int ksteps = p.K / 64;
  const int num_tiles = ((p.num_mtiles + 1) / 2) * nt_count;  // tiles of a cluster
  const uint32_t crank = cluster_ctarank();
  const int tile0 = (int)cluster_id_x();
  const int tile_step = (int)num_clusters_x();
  // (row tile, column tile) of this CTA for a cluster tile.  A phantom row tile (odd num_mtiles) loads zeros and
  // its stores are clipped by the tensor map.
  auto tile_mt = [&](int tile) { return p.mtile0 + 2 * (tile / nt_count) + (int)crank; };
  auto tile_nt = [&](int tile) { return tile % nt_count; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < C::kAcc; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * kGemmEpiWarps);  // the leader's counts the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) sbias[i] = p.bias[i];
  if constexpr (kLN) {
    for (int i = threadIdx.x; i < kHid; i += blockDim.x) {
      sgamma[i] = p.gamma[i];
      sbeta[i] = p.beta[i];
    }
    // I16: the 16 x 16 bf16 identity is the B operand (N = 16) of the residual MMAs; in a 2-SM MMA every CTA holds
    // half of the B rows, so this CTA keeps rows 8 crank .. 8 crank + 7 as one K-major SWIZZLE_128B atom (8 rows x
    // 128 B, columns 0..15 used): row i has its 1.0 at column 8 crank + i.  It lives in the part of the bias area a
    // 384-wide GEMM does not use.
    uint8_t* eye = sm + C::kParF + kEyeOff;
    for (int i = threadIdx.x; i < 8 * 8; i += blockDim.x) {  // 16-byte chunk (row i, chunk c)
      const int r = i >> 3, c = i & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c == (int)crank) {  // column 8 crank + r sits in chunk `crank`, element r of the chunk
        const uint32_t one = 0x3F80u << (16 * (r & 1));
        const int w = r >> 1;
        v.x = w == 0 ? one : 0u;
        v.y = w == 1 ? one : 0u;
        v.z = w == 2 ? one : 0u;
        v.w = w == 3 ? one : 0u;
      }
      *reinterpret_cast<uint4*>(eye + r * 128 + ((c ^ r) << 4)) = v;
    }
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc2(holder, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();  // the peer's mbarriers exist before anybody arrives on them
  const uint32_t tmem_base = *holder;
  // everything above touched only this kernel's own shared / tensor memory and the (constant) weights: it may overlap the
  // previous kernel's tail.  Activations are read, and buffers overwritten, only after the previous grid has completed.
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_a);
      tma_prefetch_desc(&tmap_b);
#ifdef FRS_GEMM_TRACE
      long long* gtr = nullptr;  // (role 2 = the peer CTA's epilogue warp in this build)
      int gtn = 0;
#endif
      uint32_t it = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int mt = tile_mt(tile), nt = tile_nt(tile);
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const uint32_t stage = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait_c(&empty[stage], ph ^ 1, 101u);
          FRS_GT(30);
          uint8_t* sa = ring + (size_t)stage * C::kStage;
          // both CTAs' boxes complete on the LEADER's barrier (its MMA thread consumes both halves)
          if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::kStage);
          const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
          tma_load_2d_cg2(sa, &tmap_a, lead_full, ks * 64, mt * kBM, kEvictNormal);
          if constexpr (kLN) {
            // columns [0,192) take weight rows 96 crank .. +96 of this CTA, columns [192,384) rows 192 + 96 crank ..
            tma_load_2d_cg2(sa + C::kStageA, &tmap_b, lead_full, ks * 64, (int)crank * (BN / 2), kEvictLast);
            tma_load_2d_cg2(sa + C::kStageA + (BN / 2) * 128, &tmap_b, lead_full, ks * 64, BN + (int)crank * (BN / 2),
                            kEvictLast);
          } else {
            tma_load_2d_cg2(sa + C::kStageA, &tmap_b, lead_full, ks * 64, nt * BN + (int)crank * (BN / 2), kEvictLast);
          }
        }
        if constexpr (kLN) {
          // the residual rides through the tensor core: six more K-steps whose A slab is the residual's own
          // [128 rows x 64 columns] box (tmap_out2 = the residual as an A operand)
          for (int r = 0; r < kResSteps; ++r, ++it) {
            const uint32_t stage = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1;
            mbar_wait_c(&empty[stage], ph ^ 1, 101u);
            if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageA);
            tma_load_2d_cg2(ring + (size_t)stage * C::kStage, &tmap_out2, mapa_u32(smem_u32(&full[stage]), 0), r * 64,
                            mt * kBM, kEvictNormal);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // The whole warp runs this loop converged; the elected lane executes the tcgen05 instructions
    // (elect.sync): descriptors stay in uniform registers and the MMAs issue back to back.
    if (crank == 0) {
      const uint32_t issue = elect_one_pred();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);  // provably warp-uniform: stays in a uniform register
      constexpr uint32_t idesc = make_idesc(1u, 2 * kBM, kNSub);
#ifdef FRS_GEMM_TRACE
      long long* gtr = (p.trace && blockIdx.x == 0 && issue) ? p.trace : nullptr;
      int gtn = 0;
#endif
      uint32_t it = 0, lt = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++lt) {
        const uint32_t acc = lt % C::kAcc;
        const uint32_t aph = (lt / C::kAcc) & 1;
        FRS_GT(1);
        {  // the peer CTA's epilogue warps arrive through the cluster: acquire at cluster scope
          uint32_t spins = 0;
          while (!mbar_try_wait_cluster(&tempty[acc], aph ^ 1)) {
            if (++spins > (1u << 22)) trap_with_code(102u, aph ^ 1);
          }
        }
        tc_fence_after();
        FRS_GT(2);
        FRS_GT(40 + acc);
        const uint32_t d_tmem = tmem_u + acc * BN;
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const uint32_t stage = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait_c(&full[stage], ph, 103u);
          tc_fence_after();
          FRS_GT(3);
          const uint32_t sa = smem_u32(ring + (size_t)stage * C::kStage);
          const uint64_t da = make_desc_sw128(sa);
          const uint64_t db = make_desc_sw128(sa + C::kStageA);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            tc_mma2_f16_pred(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (uint32_t)((ks | kk) != 0), issue);
            if constexpr (kLN) {  // second column half: the CTA's second 96-row box
              const uint64_t db1 = make_desc_sw128(sa + C::kStageA + (BN / 2) * 128);
              tc_mma2_f16_pred(d_tmem + BN, da + 2 * kk, db1 + 2 * kk, idesc, (uint32_t)((ks | kk) != 0), issue);
            }
          }
          tc_commit2_mc_pred(&empty[stage], (uint16_t)3, issue);  // frees the stage in both CTAs
        }
        if constexpr (kLN) {
          // acc[:, 64 r + 16 kk .. + 16) += residual columns [16 kk, +16) . I16^T  (bf16 x 1.0 accumulated in fp32: exact)
          constexpr uint32_t idesc_eye = make_idesc(1u, 2 * kBM, 16);
          const uint64_t deye = make_desc_sw128(smem_u32(sm + C::kParF + kEyeOff));
          for (int r = 0; r < kResSteps; ++r, ++it) {
            const uint32_t stage = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1;
            mbar_wait_c(&full[stage], ph, 103u);
            tc_fence_after();
            const uint64_t da = make_desc_sw128(smem_u32(ring + (size_t)stage * C::kStage));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma2_f16_pred(d_tmem + r * 64 + kk * 16, da + 2 * kk, deye, idesc_eye, 1u, issue);
            tc_commit2_mc_pred(&empty[stage], (uint16_t)3, issue);
          }
        }
        tc_commit2_mc_pred(&tfull[acc], (uint16_t)3, issue);  // both CTAs' epilogues
        FRS_GT(4);
      }
    }
  } else if (warp >= 2 + kGemmEpiWarps) {
    // ===================== TMA-store warps (one per column-half group) =====================
    // The epilogue warps only ARRIVE on the "staged" barrier of a box and go on with the next chunk; this warp
    // waits for it, issues the TMA store and, once the store of two chunks ago has read its box, arrives on that
    // box's "free" barrier, which an epilogue thread checks before it stages the chunk after this one.
    // (With the store issued by an epilogue thread, every chunk paid a 128-thread barrier plus ~280 clk of TMA
    // issue + read wait.)
    const uint32_t g = warp - (2 + kGemmEpiWarps);
    uint8_t* const sgroup = sout0 + g * (3 * kChunkBox);
    if (!(p.debug & 1)) {
      if (lane == 0) {
        tma_prefetch_desc(&tmap_out);
        if constexpr (EPI == kEpiQKV) tma_prefetch_desc(&tmap_out2);
      }
      constexpr uint32_t kChunks = C::kColsPerThread / 32;  // chunks of a group per tile
      int my_tiles = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) ++my_tiles;
      const uint32_t total = kChunks * (uint32_t)my_tiles;
      uint32_t n = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int mt = tile_mt(tile), nt = tile_nt(tile);
        const bool transposed = EPI == kEpiQKV && nt * BN >= 2 * kHid;
        for (uint32_t c = 0; c < kChunks; ++c, ++n) {
          const uint32_t b = n % 3;
          uint8_t* box = sgroup + b * kChunkBox;
          const int ct = (int)g * C::kColsPerThread + (int)c * 32;  // column within the tile
          named_bar_sync(kBarStaged + g * 6 + b, 160);
          if (lane == 0) {
            if (transposed) {
              tma_store_2d(&tmap_out2, box, mt * kBM, nt * BN - 2 * kHid + ct);
              tma_store_2d(&tmap_out2, box + 4096, mt * kBM + 64, nt * BN - 2 * kHid + ct);
            } else {
              tma_store_2d(&tmap_out, box, nt * BN + ct, mt * kBM);
            }
            tma_store_commit();
            // up to three stores in flight: a store takes ~700-900 clk until it has read its box (it queues behind
            // the ring's loads in the TMA unit), and waiting for each one made the store warp — and, three chunks
            // later, the epilogue — the bottleneck of the kernel (box-free waits of 750 clk per chunk)
            tma_store_wait_read<2>();
          }
          __syncwarp();
          // the store of chunk n - 2 has read its box: free it for chunk n + 1 (the last three have no taker)
          if (n >= 2 && n + 1 < total) named_bar_arrive(kBarFree + g * 6 + (n - 2) % 3, 160);
        }
      }
      if (lane == 0) tma_store_wait<0>();  // shared memory must outlive the last store
    }
  } else {
    // ===================== epilogue =====================
    const uint32_t quarter = warp & 3;           // TMEM lanes 32*quarter .. +32 are visible to this warp
    const uint32_t half = (warp - 2) >> 2;       // which half of the tile's columns
    const uint32_t row = quarter * 32 + lane;    // row within the tile
    uint8_t* const sgroup = sout0 + half * (3 * kChunkBox);
    uint32_t nchunk = 0;  // chunks staged so far by this group (ring position)
#ifdef FRS_GEMM_TRACE
    long long* gtr = (p.trace && blockIdx.x < 2 && threadIdx.x == 64) ? p.trace + (1 + blockIdx.x) * kGTraceCap : nullptr;
    int gtn = 0;
#endif
    // box of the next chunk, free again: the store warp has seen the store of three chunks ago read it
    auto next_box = [&]() -> uint8_t* {
      const uint32_t b = nchunk % 3;
      if (nchunk >= 3) named_bar_sync(kBarFree + half * 6 + b, 160);
      FRS_GT(16);
      return sgroup + b * kChunkBox;
    };
    // the chunk is in its box: hand it to the store warp and go on (no wait)
    auto chunk_staged = [&]() {
      FRS_GT(13);
      fence_proxy_async();
      named_bar_arrive(kBarStaged + half * 6 + (nchunk % 3), 160);
      FRS_GT(14);
      ++nchunk;
    };
    // stage 32 columns (packed bf16 pairs o[16]) of this thread's row as a box [128 rows x 32 cols]
    // (Staging the whole half tile and storing its three boxes behind one barrier measured 5 % SLOWER: the
    // chunk-by-chunk stores overlap the rest of the tile's epilogue.)
    auto stage_and_store = [&](const uint32_t (&o)[16]) {
      uint8_t* dst = next_box() + row * 64;  // SWIZZLE_64B: 16-byte chunk j of row r sits at j ^ ((r >> 1) & 3)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(dst + ((j ^ ((row >> 1) & 3)) << 4)) =
            make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      chunk_staged();
    };
    // the accumulator stage is drained: tell the MMA thread, which lives in the leader CTA.  The peer arrives
    // WITHOUT a cluster-scope release (~1 us per call): the payload is a drained TMEM accumulator, ordered by
    // tcgen05.wait::ld + tcgen05.fence.
    auto acc_release = [&](uint64_t* bar) {
      if (crank == 0) mbar_arrive(bar);
      else mbar_arrive_remote(mapa_u32(smem_u32(bar), 0));
    };
    uint32_t lt = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++lt) {
      const int nt = tile_nt(tile);
      const uint32_t acc = lt % C::kAcc;
      const uint32_t aph = (lt / C::kAcc) & 1;
      FRS_GT(10);
      mbar_wait_c(&tfull[acc], aph, 104u);
      tc_fence_after();
      FRS_GT(11);
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * BN + half * C::kColsPerThread;
      uint32_t v[32];
      if (p.debug & 1) {
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) acc_release(&tempty[acc]);
      } else if constexpr (EPI == kEpiQKV || EPI == kEpiGelu) {
        const bool transposed = EPI == kEpiQKV && nt * BN >= 2 * kHid;  // value projection
#pragma unroll 1
        for (int c = 0; c < C::kColsPerThread / 32; ++c) {
          const int ct = (int)half * C::kColsPerThread + c * 32;  // column within the tile
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          FRS_GT(12);
          if (c == C::kColsPerThread / 32 - 1) {  // accumulator drained: the next tile's MMAs may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) acc_release(&tempty[acc]);
            FRS_GT(50 + acc);
          }
          const float* bs = sbias + nt * BN + ct;
          if (transposed) {
            // value projection: stored transposed, vt[dim][token], so that it is the K-major B operand of
            // P.V in the attention kernel.  Two boxes [32 dims][64 tokens] (SWIZZLE_128B) per chunk.
            // Two tokens (rows r, r+1 = neighbouring lanes) share a 32-bit word of a vt row: the even lane writes
            // the even dims of the pair, the odd lane the odd dims, after one shuffle per dim pair.  (32 two-byte
            // stores per thread took ~1700 clk per chunk: ~50 clk per STS.U16 — a third of the QKV tiles.)
            const uint32_t odd = row & 1;
            uint8_t* dst = next_box() + (row >> 6) * 4096 + (row & 6) * 2 + odd * 128;
            const uint32_t tch = ((row & 63) >> 3) ^ odd;
            // (bias first, for all 32 dims: a shared-memory load behind a staging store cannot be hoisted above
            // it — the compiler has to assume they alias — and the loop ran fully serialised, 80 clk per dim pair)
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              float e = __uint_as_float(v[2 * jj]), o = __uint_as_float(v[2 * jj + 1]);
              const float2 bb = *reinterpret_cast<const float2*>(bs + 2 * jj);
              fadd2(e, o, bb.x, bb.y);
              v[2 * jj] = __float_as_uint(e);
              v[2 * jj + 1] = __float_as_uint(o);
            }
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const float e = __uint_as_float(v[2 * jj]), o = __uint_as_float(v[2 * jj + 1]);
              const float got = __shfl_xor_sync(0xffffffffu, odd ? e : o, 1);
              const uint32_t word = odd ? pack_bf16x2(got, o) : pack_bf16x2(e, got);  // low half = the lower token
              *reinterpret_cast<uint32_t*>(dst + (2 * jj) * 128 + ((tch ^ ((2 * jj) & 7)) << 4)) = word;
            }
            chunk_staged();
          } else {
            const float sc = (EPI == kEpiQKV && nt * BN < kHid) ? p.qscale : 1.0f;
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
              const float2 bb = *reinterpret_cast<const float2*>(bs + 2 * j);
              fadd2(a, b, bb.x, bb.y);
              if constexpr (EPI == kEpiGelu) {
#if defined(FRS_GELU_MIX) && !defined(FRS_EXACT_GELU)
                // odd column pairs on the FMA pipes (gelu_poly2), even ones on the MUFU.  Measured: FFN-up 0.94 ms
                // per pass against 0.84 ms with the tanh form alone (the degree-6 chain does not interleave well
                // across pairs), so this is off by default.
                if (j & 1) gelu_poly2(a, b);
                else gelu_erf2(a, b);
#else
                gelu_erf2(a, b);
#endif
              } else {
                fmul2(a, b, sc, sc);
              }
              o[j] = pack_bf16x2(a, b);
            }
            stage_and_store(o);
          }
        }
      } else {
        // LayerNorm over the 384-wide row: this thread owns 192 columns of its row, the thread of the other
        // column-half group (same lane quarter) the other 192.  Pass 1: bias (the residual is already in the
        // accumulator) and row statistics; pass 2 reads the accumulator again and repeats the one bias FADD2 per
        // pair (bit-identical), which is cheaper than writing the pre-LayerNorm value back to TMEM in between.
        // Statistics are SHIFTED: the thread accumulates sum(x - c) and sum((x - c)^2) with c = the mean of its own
        // first 32 values, so a row whose values share a large offset (|mean| >> spread) loses nothing to the
        // cancellation of E[x^2] - mean^2, and a single outlier dimension moves c by 1/32 of itself at most.  The
        // two halves' (sum, sumsq, c) are combined exactly: mean = (c0 + c1)/2 + (S0 + S1)/384,
        // sum (x - mean)^2 over half h = Q_h - 2 d_h S_h + 192 d_h^2 with d_h = mean - c_h.
        float sum = 0.f, sq = 0.f, sum1 = 0.f, sq1 = 0.f;  // even / odd columns (packed pairs)
        float shift = 0.f;
        // the CLS row of a sequence keeps an fp32 copy of the LAST layer's output (pooler / classifier / pooling read
        // it instead of the bf16-rounded row: that rounding was the largest single term of the logit error)
        int cls_slot = -1;
        if (p.cls_slot) {
          const int grow = tile_mt(tile) * kBM + (int)row;
          if (grow < p.M) cls_slot = __ldg(p.cls_slot + grow);
        }
#pragma unroll 1
        for (int c = 0; c < C::kColsPerThread / 32; ++c) {
          const int col = (int)half * C::kColsPerThread + c * 32;  // column of the 384-wide row
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          const float* bs = sbias + col;
          if (c == 0) {
            float p0 = 0.f, p1 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
              const float2 bb = *reinterpret_cast<const float2*>(bs + 2 * j);
              fadd2(a, b, bb.x, bb.y);
              fadd2(p0, p1, a, b);
            }
            shift = (p0 + p1) * (1.0f / 32.0f);
          }
          const float nshift = -shift;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
            const float2 bb = *reinterpret_cast<const float2*>(bs + 2 * j);
            fadd2(a, b, bb.x, bb.y);
            fadd2(a, b, nshift, nshift);
            fadd2(sum, sum1, a, b);
            ffma2_acc(sq, sq1, a, b, a, b);
          }
        }
        FRS_GT(17);
        sum += sum1;
        sq += sq1;
        // the two column halves of a row meet in shared memory (double-buffered by tile parity: a fast group may
        // be one tile ahead of the other group's read)
        float4* st = sstat + (lt & 1) * 256;
        st[half * 128 + row] = make_float4(sum, sq, shift, 0.f);
        named_bar_sync(1, kGemmEpiThreads);
        FRS_GT(18);
        const float4 s0 = st[row], s1 = st[128 + row];
        // the same operands in the same order in both groups: bit-identical mean / rstd for the two halves of a row
        const float mean = 0.5f * (s0.z + s1.z) + (s0.x + s1.x) * (1.0f / kHid);
        const float d0 = mean - s0.z, d1 = mean - s1.z;
        const float ss0 = fmaf(d0, fmaf(d0, (float)(kHid / 2), -2.0f * s0.x), s0.y);
        const float ss1 = fmaf(d1, fmaf(d1, (float)(kHid / 2), -2.0f * s1.x), s1.y);
        const float var = fmaxf((ss0 + ss1) * (1.0f / kHid), 0.f);
        const float rstd = rsqrtf(var + p.eps);
#pragma unroll 1
        for (int c = 0; c < C::kColsPerThread / 32; ++c) {
          const int col = (int)half * C::kColsPerThread + c * 32;  // column of the 384-wide row
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (c == C::kColsPerThread / 32 - 1) {  // accumulator drained
            tc_fence_before();
            __syncwarp();
            if (lane == 0) acc_release(&tempty[acc]);
          }
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // (((v + bias) - mean) * rstd) * gamma + beta on the packed pipes
            float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
            const float2 bb = *reinterpret_cast<const float2*>(sbias + col + 2 * j);
            const float2 gg = *reinterpret_cast<const float2*>(sgamma + col + 2 * j);
            const float2 be = *reinterpret_cast<const float2*>(sbeta + col + 2 * j);
            fadd2(a, b, bb.x, bb.y);
            fadd2(a, b, -mean, -mean);
            fmul2(a, b, rstd, rstd);
            ffma2(a, b, gg.x, gg.y, be.x, be.y);
            if (cls_slot >= 0) *reinterpret_cast<float2*>(p.cls_out + (size_t)cls_slot * kHid + col + 2 * j) = make_float2(a, b);
            o[j] = pack_bf16x2(a, b);
          }
          stage_and_store(o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's operands / arriving on its barriers
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// varlen attention over packed sequences, one (128-query block, head pair) per work item
// ---------------------------------------------------------------------------------------------
// Q and K of a head pair are 64 contiguous bf16 of a qk row = one 128-byte SWIZZLE_128B row; a head
// is selected by starting the MMA's K slices 64 bytes into the row.  V is read from its transposed
// copy vt[dim][token] so that keys are the K dimension of P.V.
// Warpgroup 0 = {TMA producer, MMA issuer, 2 idle warps}; warpgroups 1 and 2 = the softmax threads of
// head 0 / head 1 of the pair.  A softmax thread keeps a whole 128-key score row in registers, so the
// register file is re-divided at kernel start (setmaxnreg): 64 per thread for warpgroup 0, 216 for the
// softmax warpgroups.  P goes back to tensor memory (the A operand of P.V) and O stays there for the whole
// item; see the softmax branch for the lazy reference maximum and the turn-taking of the two heads.
// Two builds of the kernel (-DFRS_ATTN_KB=128 | 64):
//   128 keys per block, one role set per CTA (384 threads): a softmax thread holds a 128-score row (216 registers) —
//     the default;
//   64 keys per block, TWO independent role sets ("groups") in one 768-thread CTA: each group is the complete
//     384-thread machine — its own TMA producer, MMA issuer, eight softmax warps, its own shared-memory ring,
//     mbarriers and half of the CTA's tensor memory — working on its own items; with 64-key blocks a softmax thread
//     needs 104 registers, so both groups fit the register file.  Each SM sub-partition then hosts FOUR softmax warps
//     that are not synchronised with each other.  The idea was to keep the MUFU busy while a warp is in its load /
//     max / pack phases (one group per SM: 59-68 % MUFU utilisation).  MEASURED (B200, 128 x 512 tokens): 2.20 ms of
//     attention per 12-layer pass against 1.74 ms for the 128-key build — twice as many blocks means twice the
//     per-block fixed work (TMEM round trips, barrier hand-offs, P.V issue), the 104-register softmax threads and
//     the 32-register issuer spill, and nothing is left of the MUFU gain.  Parity-clean, kept as a build option.
//     Why not simply two CTAs per SM: a kernel that allocates tensor memory is limited to ONE resident CTA per SM
//     (scripts/micro/occ_probe.cu: cudaOccupancyMaxActiveBlocksPerMultiprocessor = 1 as soon as tcgen05.alloc appears,
//     whatever the registers and shared memory; measured: the 2-CTA build ran exactly as fast as one CTA).
#ifndef FRS_ATTN_KB
#define FRS_ATTN_KB 128
#endif
constexpr int kKB = FRS_ATTN_KB;      // keys per block
static_assert(kKB == 64 || kKB == 128, "key block is 64 or 128 keys");
constexpr int kAttnGroups = kKB == 64 ? 2 : 1;
constexpr int kAttnGroupThreads = 384;
constexpr int kAttnThreads = kAttnGroupThreads * kAttnGroups;
constexpr int kAttnRegsLow = kKB == 64 ? 32 : 64;   // per group 128 x low + 256 x high must fit what it was launched with (384 x 80 | 384 x 168)
constexpr int kAttnRegsHigh = kKB == 64 ? 104 : 216;
constexpr int kKVStages = 4;
constexpr int kQBytes = kBM * 128;            // 128 queries x 64 dims
constexpr int kKBytes = kKB * 128;            // kKB keys x 64 dims
constexpr int kVSlab = 64 * 128;              // 64 dims x 64 keys
constexpr int kVSlabs = kKB / 64;
constexpr int kVBytes = kVSlabs * kVSlab;     // kKB keys
constexpr int kKVBytes = kKBytes + kVBytes;
constexpr int kTmemCols = 4 * kKB;            // 512 | 256 (a power of two)
constexpr int kTmemS = 0;                     // TMEM columns: S of head h at [kKB h, +kKB)
constexpr int kTmemO = 2 * kKB;               //   O block of head h at [2 kKB + 32h, +32)
constexpr int kTmemP = 2 * kKB + 64;          //   P of head h (bf16 pairs) at [2 kKB + 64 + (kKB/2) h, +kKB/2)
static_assert(kTmemP + kKB <= kTmemCols, "TMEM budget");

struct AttnSmem {
  static constexpr int q = 0;                                  // [2]
  static constexpr int kv = q + 2 * kQBytes;                   // [kKVStages] K | Vt slab 0 | Vt slab 1
  static constexpr int bars = kv + kKVStages * kKVBytes;
  static constexpr int nbars = 2 + 2 + 2 * kKVStages + 8;
  static constexpr int holder = bars + nbars * 8;
  static constexpr int total = holder + 16;
  static constexpr int group = (total + 1023) / 1024 * 1024;   // a group's region (1024-aligned: SWIZZLE_128B tiles)
};

// -DFRS_ATTN_TRACE: event timeline (clock64 << 8 | event id) of CTA 0: role 0 = MMA issuer, role 1 / 2 = lane 0 of
// the first softmax warp of head 0 / head 1 (launch_attention dumps it to gpurun_out/attn_trace.txt)
#ifdef FRS_ATTN_TRACE
constexpr int kTraceCap = 4096;
#define FRS_TR(id) do { if (tr && trn < kTraceCap) tr[trn++] = (clock64() << 8) | (long long)(id); } while (0)
#else
#define FRS_TR(id) do { } while (0)
#endif
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the pointer in the shared address space: rounding the
  // pointer through uintptr_t made every staging access a generic LD.E / ST.E instead of LDS / STS)
  uint8_t* sm0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // role set ("group") of this thread: everything below is per group — shared-memory region, barriers, TMEM half, items
  const uint32_t grp = threadIdx.x / kAttnGroupThreads;
  const uint32_t gtid = threadIdx.x - grp * kAttnGroupThreads;
  uint8_t* sm = sm0 + grp * AttnSmem::group;
  const int item0 = (int)(blockIdx.x * kAttnGroups + grp), item_step = (int)(gridDim.x * kAttnGroups);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(sm + AttnSmem::bars);
  uint64_t* q_empty = q_full + 2;
  uint64_t* kv_full = q_empty + 2;
  uint64_t* kv_empty = kv_full + kKVStages;
  uint64_t* s_full = kv_empty + kKVStages;  // [2 heads]
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* o_full = p_full + 2;
  uint32_t* holder = reinterpret_cast<uint32_t*>(sm0 + AttnSmem::holder);  // (group 0's: one allocation for the CTA)

  const uint32_t warp = gtid >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int n_items = p.nqb * kHeadPairs;

  if (gtid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);  // one arrive per softmax warp of the head
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1 && grp == 0) {
    tmem_alloc(holder, kTmemCols * kAttnGroups);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder + grp * kTmemCols;
  pdl_launch_dependents();
  pdl_wait();  // (see gemm_kernel: the prologue above overlaps the previous kernel's tail)
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAttnRegsLow));
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qk);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_vt);
      uint32_t li = 0, g = 0;
      for (int item = item0; item < n_items; item += item_step, ++li) {
        const QBlock qb = p.qblk[item / kHeadPairs];
        const int hp = item % kHeadPairs;
        const uint32_t qbuf = li & 1;
        mbar_wait_bg(&q_empty[qbuf], ((li >> 1) & 1) ^ 1, 105u);
        mbar_arrive_expect_tx(&q_full[qbuf], kQBytes);
        tma_load_2d(sm + AttnSmem::q + qbuf * kQBytes, &tmap_qk, &q_full[qbuf], hp * 64, qb.q_tok0, kEvictNormal);
        const int nkb = (qb.seq_tok0 - qb.kv_tok0 + qb.seq_len + kKB - 1) / kKB;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t stage = g % kKVStages;
          mbar_wait_bg(&kv_empty[stage], ((g / kKVStages) & 1) ^ 1, 106u);
          mbar_arrive_expect_tx(&kv_full[stage], kKVBytes);
          uint8_t* dst = sm + AttnSmem::kv + (size_t)stage * kKVBytes;
          const int tok = qb.kv_tok0 + kb * kKB;
          tma_load_2d(dst, &tmap_k, &kv_full[stage], kHid + hp * 64, tok, kEvictNormal);  // box of kKB key rows
#pragma unroll
          for (int sl = 0; sl < kVSlabs; ++sl)
            tma_load_2d(dst + kKBytes + sl * kVSlab, &tmap_vt, &kv_full[stage], tok + 64 * sl, hp * 64, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop converged; only lane 0 executes the tcgen05 instructions (see
    // tc_mma_f16_pred): inside an `if (lane == 0)` region every MMA cost ~70 clk of R2UR/ELECT loops and
    // the issue of one block's P.V MMAs (555 clk) sat on the softmax warps' critical path.
    {
      const uint32_t issue = elect_one_pred();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);  // provably warp-uniform: stays in a uniform register
#ifdef FRS_ATTN_TRACE
      long long* tr = (p.timing && blockIdx.x == 0 && issue) ? p.timing + 64 : nullptr;
      int trn = 0;
#endif
      constexpr uint32_t idesc_s = make_idesc(1u, kBM, kKB);        // S = Q K^T : 128 x 128
      constexpr uint32_t idesc_o = make_idesc(1u, kBM, kHeadDim);   // O = P V   : 128 x 32
      // Two cursors over this CTA's (item, key block) sequence: the score MMAs run one block AHEAD of
      // the P.V MMAs, so that S(g+1) is computed while the softmax warps are still exponentiating S(g)
      // (they release S as soon as it is in their registers).
      struct Cursor {
        int item, kb, nkb;
        uint32_t li, g;
      };
      auto load_item = [&](Cursor& c) {
        if (c.item < n_items) {
          const QBlock qb = p.qblk[c.item / kHeadPairs];
          // (a loaded value is not provably warp-uniform; the shuffle makes the loop control uniform again)
          c.nkb = __shfl_sync(0xffffffffu, (qb.seq_tok0 - qb.kv_tok0 + qb.seq_len + kKB - 1) / kKB, 0);
        }
      };
      auto advance = [&](Cursor& c) {
        ++c.g;
        if (++c.kb == c.nkb) {
          c.kb = 0;
          c.item += item_step;
          ++c.li;
          load_item(c);
        }
      };
      auto issue_scores = [&](const Cursor& c) {
        const uint32_t qbuf = c.li & 1;
        if (c.kb == 0) mbar_wait_bg(&q_full[qbuf], (c.li >> 1) & 1, 107u);
        const uint32_t stage = c.g % kKVStages;
        mbar_wait_bg(&kv_full[stage], (c.g / kKVStages) & 1, 108u);
        const uint32_t q_addr = smem_u32(sm + AttnSmem::q + qbuf * kQBytes);
        const uint32_t k_addr = smem_u32(sm + AttnSmem::kv + (size_t)stage * kKVBytes);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait_bg(&s_free[h], (c.g & 1) ^ 1, 109u);  // softmax warps hold the previous S of this head in registers
          tc_fence_after();
          FRS_TR(1 + h);
          const uint64_t da = make_desc_sw128(q_addr) + 4 * h;  // +64 bytes: second head of the pair
          const uint64_t db = make_desc_sw128(k_addr) + 4 * h;
          tc_mma_f16_pred(tmem_u + kTmemS + kKB * h, da, db, idesc_s, 0u, issue);
          tc_mma_f16_pred(tmem_u + kTmemS + kKB * h, da + 2, db + 2, idesc_s, 1u, issue);
          tc_commit_pred(&s_full[h], issue);
          FRS_TR(3 + h);
        }
      };
      Cursor cs{item0, 0, 0, 0u, 0u}, cp{item0, 0, 0, 0u, 0u};
      load_item(cs);
      load_item(cp);
      if (cs.item < n_items) {
        issue_scores(cs);
        advance(cs);
      }
      while (cp.item < n_items) {
        if (cs.item < n_items) {
          issue_scores(cs);  // S(g+1): waits until S(g) has been read, not until P(g) is written
          advance(cs);
        }
        const uint32_t stage = cp.g % kKVStages;
        const uint32_t v_addr = smem_u32(sm + AttnSmem::kv + (size_t)stage * kKVBytes) + kKBytes;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait_bg(&p_full[h], cp.g & 1, 110u);  // P of this block is in tensor memory, previous O block was read
          tc_fence_after();
          FRS_TR(5 + h);
#pragma unroll
          for (int s = 0; s < kVSlabs; ++s) {
            const uint64_t db = make_desc_sw128(v_addr + s * kVSlab + h * (kHeadDim * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)  // 16 keys per MMA = 8 columns of packed bf16 pairs
              tc_mma_ts_pred(tmem_u + kTmemO + 32 * h, tmem_u + kTmemP + (kKB / 2) * h + (s * 4 + kk) * 8, db + 2 * kk,
                             idesc_o, (uint32_t)((cp.kb | s | kk) != 0), issue);
          }
          tc_commit_pred(&o_full[h], issue);
          FRS_TR(7 + h);
        }
        tc_commit_pred(&kv_empty[stage], issue);
        if (cp.kb == cp.nkb - 1) tc_commit_pred(&q_empty[cp.li & 1], issue);
        advance(cp);
      }
    }
  }
  } else {
    // ===================== softmax + output =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kAttnRegsHigh));
    const uint32_t quarter = warp & 3;
    const uint32_t h = (warp >> 2) - 1;
    const uint32_t row = quarter * 32 + lane;
    const uint32_t t_s = tmem_base + ((quarter * 32u) << 16) + kTmemS + kKB * h;
    const uint32_t t_o = tmem_base + ((quarter * 32u) << 16) + kTmemO + 32 * h;
    const uint32_t t_p = tmem_base + ((quarter * 32u) << 16) + kTmemP + (kKB / 2) * h;
    uint32_t g = 0;
#ifdef FRS_ATTN_TRACE
    long long* tr = (p.timing && blockIdx.x == 0 && quarter == 0 && lane == 0) ? p.timing + 64 + (1 + h) * kTraceCap : nullptr;
    int trn = 0;
#endif
    // -DFRS_ATTN_TIMING: per-phase clock sums of the softmax warps of CTA 0 (printed by launch_attention)
#ifdef FRS_ATTN_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define FRS_T(i) do { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; } while (0)
#else
#define FRS_T(i) do { } while (0)
#endif
    // O stays in TENSOR MEMORY for the whole item: the P.V MMAs of key blocks 1.. accumulate onto block 0's, and
    // a thread only touches O (a) at the end of the item and (b) when its running reference maximum has to move.
    // The reference moves lazily: softmax is invariant to the reference m as long as 2^(s - m) stays finite, so
    // a row keeps the reference of its first block until a later block exceeds it by more than 2^kLazyLog2
    // (P <= 256 is exact in bf16's range; its relative rounding does not depend on the scale).  Per block this
    // removes the O read, 64 FP32 ops and 32 registers per thread that the register-resident O cost.
    //
    // The output of an item (last O + normalisation + stores) is written while the FIRST block of the next item
    // is in flight: its P.V round trip and the QBlock fetch of the next item (prefetched one item ahead) used
    // to sit between two items with the MUFU idle.
#ifndef FRS_LAZY_LOG2
#define FRS_LAZY_LOG2 8.0f
#endif
    constexpr float kLazyLog2 = FRS_LAZY_LOG2;  // -DFRS_LAZY_LOG2=0.f: every increase of the maximum rescales (exercises the rare path)
    float m = -INFINITY, l = 0.f;
    QBlock qb_next = p.qblk[(item0 < n_items ? item0 : 0) / kHeadPairs];
    __nv_bfloat16* dst_prev = nullptr;  // context row of this thread in the item being accumulated (null: padding row)
    // writes the context row of the item that ended with block g - 1 (its accumulated O is in TMEM)
    auto flush_item = [&]() {
      mbar_wait_c(&o_full[h], (g - 1) & 1, 113u);
      tc_fence_after();
      FRS_TR(19);
      const float inv = 1.0f / l;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[16];
        tmem_ld_32x16(t_o + 16 * half, v);
        tmem_ld_wait();
        if (dst_prev != nullptr) {
          uint32_t ok[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ok[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
          uint4* dst = reinterpret_cast<uint4*>(dst_prev) + 2 * half;
          dst[0] = make_uint4(ok[0], ok[1], ok[2], ok[3]);
          dst[1] = make_uint4(ok[4], ok[5], ok[6], ok[7]);
        }
      }
      FRS_T(7);
      FRS_TR(20);
    };
    for (int item = item0; item < n_items; item += item_step) {
      const QBlock qb = qb_next;
      if (item + item_step < n_items) qb_next = p.qblk[(item + item_step) / kHeadPairs];
      const int hp = item % kHeadPairs;
      const int key_off = qb.seq_tok0 - qb.kv_tok0;  // keys of the first block before the sequence (< 8)
      const int nkb = (key_off + qb.seq_len + kKB - 1) / kKB;
      constexpr int kSC = kKB / 32;  // 32-score chunks of a row
      uint32_t sc[kSC][32];          // one score row of the block: a single pass over TMEM (all indices are static)
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        // keys [lo, hi) of this block belong to the sequence (hi - lo >= 1)
        const int lo = kb == 0 ? key_off : 0;
        const int hi = min(kKB, key_off + qb.seq_len - kb * kKB);
        const bool full = lo == 0 && hi == kKB;
        FRS_T(0);
        FRS_TR(10);
        mbar_wait_c(&s_full[h], g & 1, 111u);
        tc_fence_after();
        FRS_T(1);
        FRS_TR(11);
#pragma unroll
        for (int c = 0; c < kSC; ++c) tmem_ld_32x32(t_s + 32 * c, sc[c]);
        tmem_ld_wait();
        // S is in registers: the MMA warp may already compute the next block's scores into it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[h]);
        FRS_T(2);
        FRS_TR(12);
        if (!full) {
          auto mask = [&](uint32_t(&sv)[32], int c) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < lo || c * 32 + j >= hi) sv[j] = 0xff800000u;  // -inf
          };
#pragma unroll
          for (int c = 0; c < kSC; ++c) mask(sc[c], c);
        }
        // block maximum (scores are already in the log2 domain: q was scaled by log2e/sqrt(32))
        float mx[kSC];
#pragma unroll
        for (int c = 0; c < kSC; ++c) mx[c] = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {  // FMNMX3: two scores per instruction, kSC independent chains
#pragma unroll
          for (int c = 0; c < kSC; ++c) mx[c] = fmax3(mx[c], __uint_as_float(sc[c][j]), __uint_as_float(sc[c][j + 1]));
        }
        float bm = mx[0];
#pragma unroll
        for (int c = 1; c < kSC; ++c) bm = fmaxf(bm, mx[c]);
        FRS_T(3);
        FRS_TR(13);
        bool pv_done = g == 0;  // P.V of block g - 1 is known to have finished (its P buffer may be overwritten)
        if (kb == 0) {
          if (g > 0) {
            flush_item();  // previous item of this CTA
            pv_done = true;
          }
          m = bm;
          l = 0.f;
          const int qi = qb.q_tok0 - qb.seq_tok0 + (int)row;  // position of this query in its sequence
          dst_prev = qi < qb.seq_len ? p.ctx + (size_t)(qb.q_tok0 + row) * kHid + (hp * 2 + h) * kHeadDim : nullptr;
        } else {
          const bool move = bm > m + kLazyLog2;
          if (__any_sync(0xffffffffu, move)) {
            // rare: rescale this row's O (and l) to the new reference; rows that keep theirs multiply by 1
            const float m_new = move ? bm : m;
            const float alpha = ex2_approx(m - m_new);
            mbar_wait_c(&o_full[h], (g - 1) & 1, 112u);
            tc_fence_after();
            pv_done = true;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t v[16];
              tmem_ld_32x16(t_o + 16 * half, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
              tmem_st_32x16(t_o + 16 * half, v);
            }
            tmem_st_wait();
            l *= alpha;
            m = m_new;
          }
        }
        if (!pv_done) {
          // P.V of the previous block reads the P buffer this block overwrites; it finished while this warp
          // loaded S and took the maximum, and the check stays outside the exp pass (the other head waits for its end)
          mbar_wait_c(&o_full[h], (g - 1) & 1, 112u);
          tc_fence_after();
        }
        FRS_T(5);
        FRS_TR(15);
        // p = 2^(s - m), rounded to bf16 (the value the tensor core multiplies with V); masked keys give 0.
        // P goes straight back to tensor memory as the A operand of P.V: key k of this row = half (k & 1) of
        // column k / 2.  It never touches shared memory.
        float ps0 = 0.f, ps1 = 0.f;
        constexpr int kPairs = kKB / 2;  // score pairs = packed P words of a row
        uint32_t pw[kPairs / 16][16];    // stored 16 words at a time; a chunk is not reused while its asynchronous store may read it
#ifndef FRS_ATTN_NO_TURNS
        // The two warps of a sub-partition (same quarter, head 0 / head 1) share one MUFU unit.  Left alone
        // they run in lockstep (both are released by the same score MMAs): both exponentiate at half rate,
        // then both leave the unit idle.  They take turns instead, so that the exp pass of one head overlaps
        // the load / max phases of the other.  The float operand ties the barrier into the data flow: no exp
        // may be scheduled above it.
        // (Only with ONE group per CTA: with two, four unsynchronised warps share the unit and need no turns — and
        // the second group's barriers would not fit the 16 named barriers of a CTA.)
        if constexpr (kAttnGroups == 1) {
          if (h == 0) {
            if (g > 0) asm volatile("bar.sync %1, 64;" : "+f"(m) : "r"(5u + quarter) : "memory");
          } else {
            asm volatile("bar.sync %1, 64;" : "+f"(m) : "r"(1u + quarter) : "memory");
          }
        }
#endif
        FRS_TR(16);
        // One software-pipelined pass over the 64 score pairs: the row sums and the bf16 packing of pair
        // j - kExpLag are issued behind the exponentials of pair j.  (Written pair by pair, ptxas put each
        // FADD2 of the row sum right behind the two MUFUs it consumes: the warp stalled ~10 clk per pair on
        // the MUFU latency and an exclusive pass took 1600 clk instead of 128 x 8.)
        constexpr int kExpLag = 6;
#ifndef FRS_ATTN_FMA_OF4
#define FRS_ATTN_FMA_OF4 0
#endif
        constexpr int kExpFmaOf4 = FRS_ATTN_FMA_OF4;
        // The other head's warp is released when this pass has issued kArriveAt of its 64 score pairs, not at its
        // end: the tail of this pass then interleaves with the head of the other one on the shared MUFU (two warps
        // in the pass at once saturate the unit; one alone leaves it idle during its FADD2 / pack instructions).
#ifndef FRS_ATTN_ARRIVE_AT
#define FRS_ATTN_ARRIVE_AT kPairs
#endif
        constexpr int kArriveAt = FRS_ATTN_ARRIVE_AT;
#pragma unroll
        for (int j = 0; j < kPairs + kExpLag; ++j) {
#ifndef FRS_ATTN_NO_TURNS
          if (kAttnGroups == 1 && kArriveAt > 0 && kArriveAt < kPairs && j == kArriveAt) {
            // tied to the last exponential issued (a score register) and to m (every later FADD2 reads it)
            asm volatile("bar.arrive %2, 64;" : "+f"(m), "+r"(sc[(j - 1) >> 4][((j - 1) & 15) * 2 + 1]) : "r"((h == 0 ? 1u : 5u) + quarter) : "memory");
          }
#endif
          if (j < kPairs) {
            uint32_t(&sv)[32] = sc[j >> 4];
            const int i = (j & 15) * 2;
            float a = __uint_as_float(sv[i]), b = __uint_as_float(sv[i + 1]);
            fadd2(a, b, -m, -m);  // FADD2: one issue slot per score pair
            // kExpFmaOf4 of every four score pairs can take their exponentials on the FMA pipes (exp2_pair_fma)
            // instead of the MUFU.  Measured (-DFRS_ATTN_FMA_OF4=1..4): 1.78 / 1.84 / 2.19 / 2.57 ms per pass
            // against 1.78 with the MUFU alone — with ~200 registers live the cubic's dependent chain does not
            // interleave across pairs (36 clk per pair) — so the default is 0.
            if ((j & 3) < kExpFmaOf4) {
              exp2_pair_fma(a, b);
            } else {
              a = ex2_approx(a);
              b = ex2_approx(b);
            }
            sv[i] = __float_as_uint(a);
            sv[i + 1] = __float_as_uint(b);
          }
          if (j >= kExpLag) {
            const int jj = j - kExpLag;
            uint32_t(&sv)[32] = sc[jj >> 4];
            const int i = (jj & 15) * 2;
            const float a = __uint_as_float(sv[i]), b = __uint_as_float(sv[i + 1]);
            fadd2(ps0, ps1, a, b);
            pw[jj >> 4][jj & 15] = pack_bf16x2(a, b);
            if ((jj & 15) == 15) tmem_st_32x16(t_p + 16 * (jj >> 4), pw[jj >> 4]);
          }
        }
#ifndef FRS_ATTN_NO_TURNS
        if (kAttnGroups == 1 && (kArriveAt <= 0 || kArriveAt >= kPairs))
          asm volatile("bar.arrive %2, 64;" : "+f"(ps0), "+f"(ps1) : "r"((h == 0 ? 1u : 5u) + quarter) : "memory");
#endif
        FRS_TR(17);
        l += ps0 + ps1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[h]);
        FRS_T(6);
        FRS_TR(18);
      }
    }
    if (g > 0) flush_item();
#ifndef FRS_ATTN_NO_TURNS
    if (kAttnGroups == 1 && h == 0 && g > 0) named_bar_sync(5u + quarter, 64);  // consume the last turn of head 1
#endif
#ifdef FRS_ATTN_TIMING
    if (p.timing && lane == 0 && blockIdx.x == 0)
      for (int i = 0; i < 8; ++i) p.timing[(warp - 4) * 8 + i] = tacc[i];
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1 && grp == 0) tmem_dealloc(tmem_base, kTmemCols * kAttnGroups);
}

// ---------------------------------------------------------------------------------------------
// pooling + L2 normalisation (sentence-transformers Pooling + Normalize), cross-encoder head
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pool_normalize_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ cu,
                      const int32_t* __restrict__ row_start, int n_seqs, int pool_mode, float* __restrict__ out) {
  __shared__ float red[4];
  const int s = blockIdx.x;
  if (s >= n_seqs) return;
  const int t0 = row_start[s], t1 = t0 + (cu[s + 1] - cu[s]);
  float v[3] = {0.f, 0.f, 0.f};
  if (pool_mode == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = __bfloat162float(x[(size_t)t0 * kHid + threadIdx.x + 128 * i]);
  } else {
    for (int t = t0; t < t1; ++t) {
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] += __bfloat162float(x[(size_t)t * kHid + threadIdx.x + 128 * i]);
    }
    const float inv = 1.0f / (float)max(t1 - t0, 1);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] *= inv;
  }
  float sq = warp_sum(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  const float nrm = sqrtf((red[0] + red[1]) + (red[2] + red[3]));
  const float inv = 1.0f / fmaxf(nrm, 1e-12f);  // F.normalize(p=2, dim=1, eps=1e-12)
#pragma unroll
  for (int i = 0; i < 3; ++i) out[(size_t)s * kHid + threadIdx.x + 128 * i] = v[i] * inv;
}

__global__ void gather_rows_f32_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ row_of_tok,
                                       int n_tokens, float* __restrict__ out) {
  const int t = blockIdx.x;
  if (t >= n_tokens) return;
  const size_t r = (size_t)row_of_tok[t];
  for (int i = threadIdx.x; i < kHid; i += blockDim.x) out[(size_t)t * kHid + i] = __bfloat162float(x[r * kHid + i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, int64_t n, float* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, int64_t n, __nv_bfloat16* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
size_t gemm_smem_bytes(int epi) {
  (void)epi;
  return GemmCfg<192, false>::kTotal + 1024;  // (both configurations are within 1 KB of each other)
}
size_t attn_smem_bytes() { return (size_t)AttnSmem::group * kAttnGroups + 1024; }

cudaError_t launch_row_map(const int32_t* cu_seqlens, const int32_t* row_start, int n_seqs, int32_t* src_tok,
                           int32_t* pos_of_row, int32_t* row_of_tok, int32_t* cls_slot, cudaStream_t st) {
  if (n_seqs <= 0) return cudaSuccess;
  row_map_kernel<<<n_seqs, 128, 0, st>>>(cu_seqlens, row_start, n_seqs, src_tok, pos_of_row, row_of_tok, cls_slot);
  return cudaGetLastError();
}

cudaError_t launch_embed_ln(const int32_t* ids, const int32_t* type_ids, const int32_t* src_tok,
                            const int32_t* pos_of_row, int M, int vocab, const float* word, const float* pos,
                            const float* type, const float* gamma, const float* beta, float eps, __nv_bfloat16* x,
                            cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  embed_ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(ids, type_ids, src_tok, pos_of_row, M, vocab, word, pos, type, gamma,
                                               beta, eps, x);
  return cudaGetLastError();
}

// FRS_NO_PDL=1 launches every kernel fully serialised (A/B measurements)
static bool pdl_enabled() {
  static const bool on = getenv("FRS_NO_PDL") == nullptr;
  return on;
}

template <int BN, int EPI>
static cudaError_t launch_gemm_t(int sm_count, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                                 const CUtensorMap& to2, const GemmParams& p, cudaStream_t st) {
  static DeviceOnce once;  // (one table per template instance)
  const size_t smem = GemmCfg<BN, EPI == kEpiResLN>::kTotal + 1024;
  {
    cudaError_t e = once_per_device(once, [&] {
      return cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    });
    if (e != cudaSuccess) return e;
  }
  if (p.N % BN != 0 || p.K % 64 != 0 || p.N > 1536) return cudaErrorInvalidValue;
  GemmParams pd = p;
  static const int dbg = getenv("FRS_GEMM_DEBUG") ? atoi(getenv("FRS_GEMM_DEBUG")) : 0;
  pd.debug = dbg;
  pd.trace = nullptr;
#ifdef FRS_GEMM_TRACE
  static long long* gth = nullptr;
  static int gcalls = 0;
  const bool dump = ++gcalls == 20;  // one warm launch of this kernel class
  if (dump) {
    if (!gth) cudaHostAlloc(&gth, 3 * kGTraceCap * 8, cudaHostAllocMapped);
    memset(gth, 0, 3 * kGTraceCap * 8);
    cudaHostGetDevicePointer(&pd.trace, gth, 0);
  }
  struct Dump {
    bool on; int epi; long long* h; cudaStream_t st;
    ~Dump() {
      if (!on) return;
      cudaStreamSynchronize(st);
      char name[64];
      snprintf(name, sizeof name, "gpurun_out/gemm_trace_%d.txt", epi);
      FILE* f = fopen(name, "w");
      if (!f) return;
      for (int r = 0; r < 3; ++r)
        for (int i = 0; i < kGTraceCap && h[r * kGTraceCap + i]; ++i)
          fprintf(f, "%d %lld %lld\n", r, h[r * kGTraceCap + i] >> 8, h[r * kGTraceCap + i] & 255);
      fclose(f);
    }
  } dumper{dump, EPI, gth, st};
#endif
  // clusters of two CTAs, persistent over the cluster's tiles: ResLN = the column halves of one 128-row tile,
  // QKV / GELU = two consecutive row tiles of one 192-column tile under one 2-SM MMA (see gemm_kernel)
  if (EPI == kEpiResLN && p.N != 2 * BN) return cudaErrorInvalidValue;
  if (p.num_mtiles <= 0) return cudaSuccess;
  const int cluster_tiles = ((p.num_mtiles + 1) / 2) * (EPI == kEpiResLN ? 1 : p.N / BN);
  const int clusters = cluster_tiles < sm_count / 2 ? cluster_tiles : sm_count / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // prologue overlaps the previous kernel's tail (pdl_wait)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_kernel<BN, EPI>, ta, tb, to, to2, pd);
}

cudaError_t launch_gemm(int epi, int sm_count, const CUtensorMap& tmap_a, const CUtensorMap& tmap_b,
                        const CUtensorMap& tmap_out, const CUtensorMap& tmap_out2, const GemmParams& p,
                        cudaStream_t st) {
  switch (epi) {
    case kEpiQKV: return launch_gemm_t<192, kEpiQKV>(sm_count, tmap_a, tmap_b, tmap_out, tmap_out2, p, st);
    case kEpiGelu: return launch_gemm_t<192, kEpiGelu>(sm_count, tmap_a, tmap_b, tmap_out, tmap_out2, p, st);
    case kEpiResLN: return launch_gemm_t<192, kEpiResLN>(sm_count, tmap_a, tmap_b, tmap_out, tmap_out2, p, st);
    default: return cudaErrorInvalidValue;
  }
}

int attn_key_block() { return kKB; }

static cudaError_t launch_attention_kernel(int grid, size_t smem, cudaStream_t st, const CUtensorMap& tmap_qk,
                                           const CUtensorMap& tmap_k, const CUtensorMap& tmap_vt, const AttnParams& p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attention_kernel, tmap_qk, tmap_k, tmap_vt, p);
}

cudaError_t launch_attention(int sm_count, const CUtensorMap& tmap_qk, const CUtensorMap& tmap_k, const CUtensorMap& tmap_vt,
                             const AttnParams& p, cudaStream_t st) {
  static DeviceOnce once;
  const size_t smem = attn_smem_bytes();
  cudaError_t e = once_per_device(once, [&] {
    cudaError_t r = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r != cudaSuccess) return r;
    if (getenv("FRS_DEBUG_OCC")) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, attention_kernel, kAttnThreads, smem);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, attention_kernel);
      int nb0 = 0, nb1 = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb0, attention_kernel, kAttnThreads, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, attention_kernel, kAttnThreads, 64 * 1024);
      fprintf(stderr, "[attention] %d keys per block, %zu B dynamic + %zu B static shared memory per CTA, %d registers, %d resident "
              "CTA(s) per SM (%d role set(s) per CTA; %d with no dynamic smem, %d with 64 KB)\n", kKB, smem, fa.sharedSizeBytes, fa.numRegs, nb,
              kAttnGroups, nb0, nb1);
    }
    return cudaSuccess;
  });
  if (e != cudaSuccess) return e;
  const int items = p.nqb * kHeadPairs;
  if (items <= 0) return cudaSuccess;
  const int ctas = (items + kAttnGroups - 1) / kAttnGroups;  // a CTA's groups take consecutive items
  const int grid = ctas < sm_count ? ctas : sm_count;
#ifdef FRS_ATTN_TRACE
  {
    static long long* th = nullptr;
    static int calls = 0;
    const int n = 64 + 3 * kTraceCap;
    if (!th) cudaHostAlloc(&th, n * 8, cudaHostAllocMapped);
    memset(th, 0, n * 8);
    AttnParams pd = p;
    cudaHostGetDevicePointer(&pd.timing, th, 0);
    attention_kernel<<<grid, kAttnThreads, smem, st>>>(tmap_qk, tmap_k, tmap_vt, pd);
    cudaStreamSynchronize(st);
    if (++calls == 30) {
      FILE* f = fopen("gpurun_out/attn_trace.txt", "w");
      if (f) {
        for (int r = 0; r < 3; ++r)
          for (int i = 0; i < kTraceCap && th[64 + r * kTraceCap + i]; ++i)
            fprintf(f, "%d %lld %lld\n", r, th[64 + r * kTraceCap + i] >> 8, th[64 + r * kTraceCap + i] & 255);
        fclose(f);
      }
    }
    return cudaGetLastError();
  }
#endif
#ifdef FRS_ATTN_TIMING
  static long long* th = nullptr;
  static int calls = 0;
  if (!th) cudaHostAlloc(&th, 64 * 8, cudaHostAllocMapped);
  AttnParams pd = p;
  cudaHostGetDevicePointer(&pd.timing, th, 0);
  attention_kernel<<<grid, kAttnThreads, smem, st>>>(tmap_qk, tmap_k, tmap_vt, pd);
  if (++calls % 12 == 0) {
    cudaStreamSynchronize(st);
    static const char* names[8] = {"loop", "wait_s", "ld_s", "max", "wait_o", "o_upd", "exp+store", "item_tail"};
    for (int w = 0; w < 8; w += 4) {
      fprintf(stderr, "[attn timing] CTA0 warp %d:", w + 4);
      for (int i = 0; i < 8; ++i) fprintf(stderr, " %s=%lld", names[i], th[w * 8 + i]);
      fprintf(stderr, "\n");
    }
  }
  return cudaGetLastError();
#endif
  return launch_attention_kernel(grid, smem, st, tmap_qk, tmap_k, tmap_vt, p);
}

cudaError_t launch_pool_normalize(const __nv_bfloat16* x, const int32_t* cu_seqlens, const int32_t* row_start,
                                  int n_seqs, int pool_mode, float* out, cudaStream_t st) {
  if (n_seqs <= 0) return cudaSuccess;
  pool_normalize_kernel<<<n_seqs, 128, 0, st>>>(x, cu_seqlens, row_start, n_seqs, pool_mode, out);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows_f32(const __nv_bfloat16* x, const int32_t* row_of_tok, int n_tokens, float* out,
                                   cudaStream_t st) {
  if (n_tokens <= 0) return cudaSuccess;
  gather_rows_f32_kernel<<<n_tokens, 128, 0, st>>>(x, row_of_tok, n_tokens, out);
  return cudaGetLastError();
}
cudaError_t launch_bf16_to_f32(const __nv_bfloat16* src, int64_t n, float* dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  bf16_to_f32_kernel<<<grid, 256, 0, st>>>(src, n, dst);
  return cudaGetLastError();
}
cudaError_t launch_f32_to_bf16(const float* src, int64_t n, __nv_bfloat16* dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  f32_to_bf16_kernel<<<grid, 256, 0, st>>>(src, n, dst);
  return cudaGetLastError();
}

}  // namespace frs

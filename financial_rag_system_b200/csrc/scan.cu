// scan.cu — exact cosine top-k over the chunk store, sm_100a.
//
// Replaces Qdrant's `query_points(query=vec, limit=15, query_filter=Filter(must=[ticker==T,...]))`
// (reference main.py:215-239, main2.py:160-163; collection schema ingest.py:86-96) with an exact
// brute-force scan.  Two phases make it both HBM-rate and exact:
//
//   phase 1 (this file, scan_kernel): every row's score against all 32 queries is computed by the
//     tensor cores (tcgen05.mma, bf16 or tf32, fp32 accumulate in TMEM) from 128-row tiles that TMA
//     streams once from HBM.  Those scores carry a bounded error eps, so they are only used as a
//     PRE-FILTER: a row can be dropped as soon as k rows with pre-filter score >= its own + 2*eps
//     are known.  Survivors go to small per-CTA candidate lists in shared memory.
//   phase 2 (merge_kernel): the lists of all CTAs are merged, the candidates inside the 2*eps band
//     of the k-th best are re-scored exactly (fp64 dot of the stored row with the prepared query)
//     and ordered by (score desc, row id asc).
//
// The result therefore equals the exact top-k of the fp64 scores whatever the grid size, the CTA
// schedule or the GPU count — the property the parity tests check against oracle/search_oracle.py.
//
// Warp roles in scan_kernel (352 threads, one CTA per SM, persistent over tiles b, b+G, ...):
//   warp 0    TMA producer: one 128x128B slab (SWIZZLE_128B) per mbarrier stage, ring of slabs
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer (M=128 rows, N=32 queries, K=16|8)
//   warp 10   bound refresher: raises the pass thresholds from the cross-CTA table in the background
//   warp 2-9  epilogue, 2 per SM sub-partition so their latencies overlap: warp w reads TMEM lane
//             group w%4 (32 rows) x 16 query columns with tcgen05.ld, applies the payload filter
//             and the per-query thresholds (fast path: 16 compares per row), and appends survivors
//             to the CTA's per-query candidate lists (rare path).  No stack anywhere: the CTA
//             leaves no L1, so local memory would cost an L2 round trip per access.
#include "common.cuh"
#include "scan.cuh"

namespace frs {

template <bool F32>
struct ScanCfg {
  static constexpr int kElemBytes = F32 ? 4 : 2;
  static constexpr int kSlabK = 128 / kElemBytes;       // elements per 128 B slab row (32 | 64)
  static constexpr int kSlabs = kDim / kSlabK;          // slabs per tile (12 | 6)
  static constexpr int kRing = F32 ? 8 : 10;            // slabs in flight
  static constexpr int kQSlabBytes = kNQ * 128;         // 4 KiB
  static constexpr int kQBytes = kSlabs * kQSlabBytes;  // 48 KiB | 24 KiB
  static constexpr int kMmasPerSlab = 4;                // each advances 32 B along K
};

// ---- shared-memory carve-up (offsets from a 1024-aligned base) ------------------------------
template <bool F32>
struct ScanSmem {
  using C = ScanCfg<F32>;
  static constexpr size_t ring = 0;
  static constexpr size_t qop = ring + (size_t)C::kRing * kSlabBytes;
  static constexpr size_t keys = qop + C::kQBytes;                       // u64 [32][kListCap]
  static constexpr size_t scratch = keys + (size_t)kNQ * kListCap * 8;   // f64 [kEpiWarps][kListCap]
  static constexpr size_t cnt = scratch + (size_t)kEpiWarps * kListCap * 8;  // u32 [32]
  static constexpr size_t taua = cnt + kNQ * 4;                          // f32 [32]
  static constexpr size_t qcode = taua + kNQ * 4;
  static constexpr size_t qmask = qcode + kNQ * 4;
  static constexpr size_t lmax = qmask + kNQ * 4;                        // u32 [32] ordered best score
  static constexpr size_t stage = lmax + kNQ * 4;                        // f32 [kEpiWarps][32][kQW+1]
  static constexpr size_t bars = stage + (size_t)kEpiWarps * 32 * (kQW + 1) * 4;  // full[R] empty[R] tfull[A] tempty[A] qbar
  static constexpr size_t nbars = 2 * C::kRing + 2 * kAccStages + 1;
  static constexpr size_t holder = bars + nbars * 8;  // u32 TMEM base, u32 epilogue-done flag
  static constexpr size_t tiles = holder + 16;        // u32 [kMaxTileSlots]: this CTA's tiles of a restricted scan
  static constexpr size_t total = tiles + (size_t)kMaxTileSlots * 4;
};

size_t scan_smem_bytes(bool f32) {
  return (f32 ? ScanSmem<true>::total : ScanSmem<false>::total) + 1024;  // + alignment slack
}

// ---- keys -----------------------------------------------------------------------------------
// (pre-filter score, row) packed so that a larger key is a better candidate and ties on the score
// go to the lower row id.
__device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
  return ((uint64_t)f32_ordered(s) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__device__ __forceinline__ float key_score(uint64_t k) { return f32_from_ordered((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// Monotone map double -> uint64 (same construction as f32_ordered).
__device__ __forceinline__ uint64_t f64_ordered(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_ordered(uint64_t k) {
  uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// ---- exact score ----------------------------------------------------------------------------
// fp64 dot product of stored row `row` with the prepared query, by one warp.  Every product of two
// fp32-representable values is exact in fp64; the summation order is fixed (lane-strided partial
// sums, xor butterfly), so the value depends only on the data.  All lanes return the same bits.
template <bool F32>
__device__ __forceinline__ double exact_dot(const void* __restrict__ rows, uint32_t row,
                                            const float* __restrict__ q) {
  const uint32_t lane = lane_id();
  double acc = 0.0;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  if constexpr (F32) {
    const float4* a4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rows) + (size_t)row * kDim);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 a = __ldg(a4 + lane + 32 * c);
      const float4 b = __ldg(q4 + lane + 32 * c);
      acc = fma((double)a.x, (double)b.x, acc);
      acc = fma((double)a.y, (double)b.y, acc);
      acc = fma((double)a.z, (double)b.z, acc);
      acc = fma((double)a.w, (double)b.w, acc);
    }
  } else {
    const uint2* a2 = reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(rows) + (size_t)row * kDim);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint2 a = __ldg(a2 + lane + 32 * c);
      const float4 b = __ldg(q4 + lane + 32 * c);
      acc = fma((double)__uint_as_float(a.x << 16), (double)b.x, acc);
      acc = fma((double)__uint_as_float(a.x & 0xFFFF0000u), (double)b.y, acc);
      acc = fma((double)__uint_as_float(a.y << 16), (double)b.z, acc);
      acc = fma((double)__uint_as_float(a.y & 0xFFFF0000u), (double)b.w, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// ---- warp selection ---------------------------------------------------------------------------
// Bitonic sort of one value per lane, descending: lane i ends up with the i-th largest.  15
// shuffle steps; small code on purpose (see slow_path).
__device__ __forceinline__ float warp_sort_desc(float x) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (uint32_t k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (uint32_t j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      const float y = __shfl_xor_sync(0xffffffffu, x, j2);
      const bool up = (lane & k2) == 0;
      const bool lower = (lane & j2) == 0;
      x = (lower == up) ? fmaxf(x, y) : fminf(x, y);
    }
  }
  return x;
}
__device__ __forceinline__ float warp_kth_largest(float x, int k) {
  return __shfl_sync(0xffffffffu, warp_sort_desc(x), k - 1);
}

// Running top-kMaxK of a stream of values, kept sorted descending in registers.
__device__ __forceinline__ void topk_insert(float (&t)[kMaxK], float v) {
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) {
    const float hi = fmaxf(t[i], v);
    v = fminf(t[i], v);
    t[i] = hi;
  }
}
// The array is pre-filled with (kMaxK - k) copies of +inf, so that the k-th largest of the inserted
// values always ends up in the LAST slot: a static register index (a runtime index would push the
// array into local memory).
__device__ __forceinline__ void topk_init(float (&t)[kMaxK], int k) {
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) t[i] = (i < kMaxK - k) ? INFINITY : -INFINITY;
}
__device__ __forceinline__ float topk_kth(const float (&t)[kMaxK]) { return t[kMaxK - 1]; }

// Cross-CTA bound.  Each lane holds the best appended pre-filter scores of up to kGmaxPerLane CTAs
// for one query.  The lane maxima belong to 32 disjoint CTA groups, i.e. to distinct rows, so the
// k-th largest of them is a lower bound of the global k-th best pre-filter score (-inf when fewer
// than k groups have a finite value).
__device__ __forceinline__ float warp_kth_of_lane_max(const float (&v)[kGmaxPerLane], int k) {
  float m = v[0];
#pragma unroll
  for (int i = 1; i < kGmaxPerLane; ++i) m = fmaxf(m, v[i]);
  return warp_kth_largest(m, k);
}

// max on a float in shared memory with integer atomics (works for any sign, +-inf included)
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ---- list compaction (one warp, one query) ----------------------------------------------------
// Keeps every entry that can still be in the exact top-k of the rows this CTA has seen:
// all entries with pre-filter score >= A_k - 2*eps (A_k = k-th best pre-filter score in the list).
// If more than kKeep entries sit in that band the band is resolved exactly (fp64) and only the
// exact top-k stay.  Raises the pass threshold taua[q] accordingly.
template <bool F32>
__device__ __forceinline__ void compact_list(uint64_t* __restrict__ L, uint32_t* cnt_q, float* taua_q,
                                          double* __restrict__ scratch, const ScanParams& p, int q,
                                          unsigned long long& n_resolutions) {
  const uint32_t lane = lane_id();
  const uint32_t c = min(*cnt_q, (uint32_t)kListCap);
  const int k = p.k;
  __syncwarp();
  const uint64_t e0 = lane < c ? L[lane] : 0ull;
  const uint64_t e1 = lane + 32 < c ? L[lane + 32] : 0ull;
  uint32_t r0 = 0, r1 = 0;
  for (uint32_t j0 = 0; j0 < c; j0 += 8) {
    uint64_t kj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) kj[u] = (j0 + u < c) ? L[j0 + u] : 0ull;  // 0 never outranks an entry
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      r0 += kj[u] > e0;
      r1 += kj[u] > e1;
    }
  }
  float cutoff = -INFINITY;
  if (c >= (uint32_t)k) {
    uint64_t ak = 0;
    if (lane < c && r0 == (uint32_t)(k - 1)) ak = e0;
    if (lane + 32 < c && r1 == (uint32_t)(k - 1)) ak = e1;
    ak = warp_max_u64(ak);
    cutoff = __fsub_rd(key_score(ak), 2.0f * p.eps);
  }
  const bool k0 = lane < c && key_score(e0) >= cutoff;
  const bool k1 = lane + 32 < c && key_score(e1) >= cutoff;
  const uint32_t nkeep = __popc(__ballot_sync(0xffffffffu, k0)) + __popc(__ballot_sync(0xffffffffu, k1));
  __syncwarp();
  // kept entries are a prefix in key order, so rank == destination slot
  if (k0) L[r0] = e0;
  if (k1) L[r1] = e1;
  __syncwarp();
  if (nkeep <= (uint32_t)kKeep) {
    if (lane == 0) {
      *cnt_q = nkeep;
      atomic_max_f32(taua_q, cutoff);
    }
    __syncwarp();
    return;
  }
  // ---- dense band: exact resolution -----------------------------------------------------------
  n_resolutions++;
  const float* qv = p.qrec + (size_t)q * kDim;
  for (uint32_t j = 0; j < nkeep; ++j) {
    const double ex = exact_dot<F32>(p.rows, key_row(L[j]), qv);
    if (lane == 0) scratch[j] = ex;
  }
  __syncwarp();
  const uint64_t g0 = lane < nkeep ? L[lane] : 0ull;
  const uint64_t g1 = lane + 32 < nkeep ? L[lane + 32] : 0ull;
  const double x0 = lane < nkeep ? scratch[lane] : 0.0;
  const double x1 = lane + 32 < nkeep ? scratch[lane + 32] : 0.0;
  const uint32_t row0 = key_row(g0), row1 = key_row(g1);
  uint32_t xr0 = 0, xr1 = 0;
  for (uint32_t j = 0; j < nkeep; ++j) {
    const double xj = scratch[j];
    const uint32_t rj = key_row(L[j]);
    xr0 += (xj > x0) || (xj == x0 && rj < row0);
    xr1 += (xj > x1) || (xj == x1 && rj < row1);
  }
  // exact k-th best
  uint64_t ek = 0;
  if (lane < nkeep && xr0 == (uint32_t)(k - 1)) ek = f64_ordered(x0);
  if (lane + 32 < nkeep && xr1 == (uint32_t)(k - 1)) ek = f64_ordered(x1);
  ek = warp_max_u64(ek);
  const double exk = f64_from_ordered(ek);
  __syncwarp();
  if (lane < nkeep && xr0 < (uint32_t)k) L[xr0] = g0;
  if (lane + 32 < nkeep && xr1 < (uint32_t)k) L[xr1] = g1;
  if (lane == 0) {
    *cnt_q = (uint32_t)k;
    float t = __fsub_rd(__double2float_rd(exk), p.eps);
    if (cutoff > t) t = cutoff;
    atomic_max_f32(taua_q, t);
  }
  __syncwarp();
}

// ---- epilogue state shared by the rare-path functions -------------------------------------------
struct Epi {
  uint64_t* keys;     // smem [32][kListCap]
  double* scratch;    // smem [kEpiWarps][kListCap]
  uint32_t* cnt;      // smem [32]
  float* taua;        // smem [32] pass thresholds (pre-filter domain)
  uint32_t* lmax;     // smem [32] ordered best appended score
  uint32_t ew;        // epilogue warp 0..kEpiWarps-1: owns queries ew, ew+kEpiWarps, ...
  uint32_t q0;        // first query column this warp scores (kQW columns)
  float* stage;       // smem [32][kQW+1]: this warp's scores of the current tile (rare path only)
  unsigned long long n_app, n_comp, n_res;
};

// Background refresher (its own warp): keeps raising the pass thresholds from the cross-CTA table
// gmax[cta][query] while the epilogue warps work.  Lane q handles query q: the k-th largest of the
// CTAs' best appended scores belongs to k distinct rows (CTAs scan disjoint rows), so it bounds the
// global k-th best pre-filter score from below; rows under it minus 2*eps can be dropped by every
// CTA.  Thresholds only ever rise (atomic max), so when an update lands does not matter for
// correctness.  Small code on purpose (see slow_path).
__device__ __forceinline__ void refresher_loop(float* taua, const uint32_t* lmax, const volatile uint32_t* done,
                                               const ScanParams& p) {
  const uint32_t lane = lane_id();
  while (*done == 0u) {
    float t[kMaxK];
    topk_init(t, p.k);
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < gridDim.x; c0 += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u)  // 16 rows of the table in flight: one L2 round trip per group
        v[u] = (c0 + u < gridDim.x) ? __ldcg(p.gmax + (size_t)(c0 + u) * kNQ + lane) : -INFINITY;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (c0 + u == blockIdx.x) v[u] = f32_from_ordered(lmax[lane]);
        topk_insert(t, v[u]);
      }
    }
    const float kth = topk_kth(t);
    if ((int)lane < p.nq && kth > -INFINITY) atomic_max_f32(&taua[lane], __fsub_rd(kth, 2.0f * p.eps));
    __nanosleep(1000);
  }
}

// The rare path of the epilogue: append the rows that passed, compact full lists, retry.
// Compact code on purpose (dynamic query index, scores in local memory): it runs for few tiles and
// must not bloat the instruction footprint of the per-tile fast path.
template <bool F32>
__device__ __forceinline__ void slow_path(Epi& e, const ScanParams& p, const uint32_t (&v)[kQW], uint32_t pend,
                                          uint32_t row) {
  // Two rules shape this code.  (1) No local memory: the CTA leaves almost no L1 (the 228 KB are
  // carved out as shared memory), so a stack access is an L2 round trip.  (2) Small code: this
  // path runs for a handful of tiles, its instructions are fetched cold while L2 and HBM are
  // saturated by the scan's own traffic, and a straight-line unrolled version was measured at
  // ~30 us for one pass.  Hence: scores staged in shared memory, real loops, dynamic query index.
  const uint32_t lane = lane_id();
  const float eps2 = 2.0f * p.eps;
  float* st = e.stage + lane * (kQW + 1);
#pragma unroll
  for (int j = 0; j < kQW; ++j) st[j] = __uint_as_float(v[j]);
  __syncwarp();
  do {
    // query columns (bit j <-> query q0 + j) with a pending row anywhere in this warp
    uint32_t wq = pend;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wq |= __shfl_xor_sync(0xffffffffu, wq, o);
#pragma unroll 1
    while (wq) {
      const int j = __ffs(wq) - 1;
      wq &= wq - 1;
      const int q = e.q0 + j;
      const float sq = st[j];
      bool mine = (pend >> j) & 1u;
      uint32_t b = __ballot_sync(0xffffffffu, mine);
      if (__popc(b) > p.k) {
        // flood: more than k rows of this warp pass.  They are distinct rows, so the k-th best of
        // them bounds the k-th best overall from below; rows under it minus 2*eps are out for good.
        const float thr = __fsub_rd(warp_kth_largest(mine ? sq : -INFINITY, p.k), eps2);
        if (mine && sq < thr) {
          mine = false;
          pend &= ~(1u << j);
        }
        if (lane == 0) atomic_max_f32(&e.taua[q], thr);
        b = __ballot_sync(0xffffffffu, mine);
      }
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&e.cnt[q], (uint32_t)__popc(b));  // one atomic per warp
      base = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t slot = base + __popc(b & ((1u << lane) - 1u));
      const bool put = mine && slot < (uint32_t)kListCap;
      if (put) {
        e.keys[q * kListCap + slot] = make_key(sq, row);
        pend &= ~(1u << j);
        e.n_app++;
      }
      float m = put ? sq : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0 && m > -INFINITY) atomicMax(&e.lmax[q], f32_ordered(m));
    }
    named_bar_sync(1, kEpiThreads);
#pragma unroll 1
    for (int q = e.ew; q < kNQ; q += kEpiWarps) {
      if (e.cnt[q] >= (uint32_t)kListCap) {
        compact_list<F32>(e.keys + q * kListCap, &e.cnt[q], &e.taua[q], e.scratch + e.ew * kListCap, p, q, e.n_res);
        e.n_comp++;
      }
    }
    named_bar_sync(1, kEpiThreads);
    // rows that could not be appended (list was full): drop those the new threshold rules out
    uint32_t rest = pend;
#pragma unroll 1
    while (rest) {
      const int j = __ffs(rest) - 1;
      rest &= rest - 1;
      if (!(st[j] >= e.taua[e.q0 + j])) pend &= ~(1u << j);
    }
  } while (named_bar_or(1, kEpiThreads, pend != 0));
  // publish this CTA's best scores for the cross-CTA bound
  if (lane < kNQ / kEpiWarps) {
    const int q = e.ew + kEpiWarps * lane;
    if (q < p.nq) __stcg(p.gmax + (size_t)blockIdx.x * kNQ + q, f32_from_ordered(e.lmax[q]));
  }
}

// ---------------------------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------------------------
template <bool F32, bool DUMP>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
            const ScanParams p) {
  using C = ScanCfg<F32>;
  using S = ScanSmem<F32>;
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the pointer in the shared address space: rounding the
  // pointer through uintptr_t made every staging access a generic LD.E / ST.E instead of LDS / STS)
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = sm + S::ring;
  uint8_t* qop = sm + S::qop;
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm + S::keys);
  double* scratch = reinterpret_cast<double*>(sm + S::scratch);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(sm + S::cnt);
  float* taua = reinterpret_cast<float*>(sm + S::taua);
  uint32_t* qcode = reinterpret_cast<uint32_t*>(sm + S::qcode);
  uint32_t* qmask = reinterpret_cast<uint32_t*>(sm + S::qmask);
  uint32_t* lmax = reinterpret_cast<uint32_t*>(sm + S::lmax);
  float* stage_all = reinterpret_cast<float*>(sm + S::stage);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::bars);
  uint64_t* empty = full + C::kRing;
  uint64_t* tfull = empty + C::kRing;
  uint64_t* tempty = tfull + kAccStages;
  uint64_t* qbar = tempty + kAccStages;
  uint32_t* holder = reinterpret_cast<uint32_t*>(sm + S::holder);
  uint32_t* stile = reinterpret_cast<uint32_t*>(sm + S::tiles);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kRing; ++i) {
      mbar_init(&full[i], 1);   // producer's arrive.expect_tx (+ TMA complete_tx bytes)
      mbar_init(&empty[i], 1);  // tcgen05.commit from the MMA thread
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull[i], 1);   // tcgen05.commit
      mbar_init(&tempty[i], kEpiWarps);  // one arrive per epilogue warp
    }
    mbar_init(qbar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < kNQ) {
    const int t = threadIdx.x;
    cnt[t] = 0;
    // padded queries never pass; live ones start from the prep kernel's bootstrap threshold (-inf when it found none)
    taua[t] = t < p.nq ? (DUMP ? -INFINITY : __ldcg(p.tau0 + t)) : INFINITY;
    qcode[t] = t < p.nq ? p.qcode[t] : 0u;
    qmask[t] = t < p.nq ? p.qmask[t] : 0u;
    lmax[t] = f32_ordered(-INFINITY);
  }
  if (threadIdx.x == 0) holder[1] = 0u;  // epilogue-done flag for the refresher warp
  // Restricted scan (ticker-segmented search): slot t of the tile sequence is tile p.tile_ids[t]; this
  // CTA's slots blockIdx.x, +G, +2G ... are staged in shared memory once.  Full scan: slot == tile.
  if (p.tile_ids)
    for (uint32_t i = threadIdx.x; blockIdx.x + i * gridDim.x < p.num_tiles; i += blockDim.x)
      stile[i] = __ldg(p.tile_ids + blockIdx.x + i * gridDim.x);
#define FRS_TILE(lt) (p.tile_ids ? stile[(lt)] : blockIdx.x + (lt) * gridDim.x)
  if (warp == 1) {
    tmem_alloc(holder, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  // optional timeline (diagnostics): per CTA {start, first slab landed, last MMA issued,
  // first tile consumed, last tile consumed, exit} in globaltimer ns
#define FRS_TL(slot) do { if (p.timeline) p.timeline[(size_t)blockIdx.x * 16 + (slot)] = globaltimer_ns(); } while (0)
  if (threadIdx.x == 0) FRS_TL(0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_rows);
      tma_prefetch_desc(&tmap_q);
      mbar_arrive_expect_tx(qbar, C::kQBytes);
      for (int s = 0; s < C::kSlabs; ++s)
        tma_load_2d(qop + s * C::kQSlabBytes, &tmap_q, qbar, s * C::kSlabK, 0, kEvictLast);
      uint32_t it = 0;
      for (uint32_t slot = blockIdx.x, plt = 0; slot < p.num_tiles; slot += gridDim.x, ++plt) {
        const uint32_t tile = FRS_TILE(plt);
        for (int s = 0; s < C::kSlabs; ++s, ++it) {
          const uint32_t stage = it % C::kRing;
          const uint32_t ph = (it / C::kRing) & 1;
          mbar_wait(&empty[stage], ph ^ 1);
          mbar_arrive_expect_tx(&full[stage], kSlabBytes);
          tma_load_2d(ring + (size_t)stage * kSlabBytes, &tmap_rows, &full[stage], s * C::kSlabK,
                      (int32_t)(tile * kTileM), kEvictFirst);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(F32 ? 2u : 1u, kTileM, kNQ);
      mbar_wait(qbar, 0);
      tc_fence_after();
      uint32_t it = 0, lt = 0;
      for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t acc = lt % kAccStages;
        const uint32_t aph = (lt / kAccStages) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kNQ;
        for (int s = 0; s < C::kSlabs; ++s, ++it) {
          const uint32_t stage = it % C::kRing;
          const uint32_t ph = (it / C::kRing) & 1;
          mbar_wait(&full[stage], ph);
          tc_fence_after();
          if (it == 0) FRS_TL(1);
          const uint64_t da = make_desc_sw128(smem_u32(ring + (size_t)stage * kSlabBytes));
          const uint64_t db = make_desc_sw128(smem_u32(qop + s * C::kQSlabBytes));
#pragma unroll
          for (int kk = 0; kk < C::kMmasPerSlab; ++kk)
            tc_mma<F32>(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (uint32_t)((s | kk) != 0));
          tc_commit(&empty[stage]);  // slab free once these MMAs have read it
        }
        tc_commit(&tfull[acc]);  // accumulator complete
      }
      FRS_TL(2);
    }
  } else if (warp == 2 + kEpiWarps) {
    // ===================== bound refresher =====================
    if constexpr (!DUMP) {
      if (gridDim.x >= (uint32_t)p.k) refresher_loop(taua, lmax, reinterpret_cast<volatile uint32_t*>(holder + 1), p);
    }
  } else {
    // ===================== epilogue =====================
    Epi e;
    e.keys = keys; e.scratch = scratch; e.cnt = cnt; e.taua = taua; e.lmax = lmax;
    e.ew = warp - 2;                // which queries this warp owns (compaction, refresh, hand-over)
    e.q0 = (e.ew >> 2) * kQW;       // query columns this warp scores
    e.stage = stage_all + (size_t)e.ew * 32 * (kQW + 1);
    e.n_app = e.n_comp = e.n_res = 0;
    const uint32_t lg = warp & 3;   // TMEM lane group (rows) this warp may read
    uint32_t lt = 0;
    unsigned long long t_wait = 0, t_slow = 0, n_slow = 0;  // diagnostics (timeline mode)
    // Payload codes are prefetched three tiles ahead.  Under the scan's own traffic (~24 MB of TMA
    // requests in flight chip-wide) a DRAM access takes a few microseconds, more than one tile
    // period, and must not sit on the per-tile critical path of the epilogue.
    auto load_code = [&](uint32_t l) -> uint32_t {  // l = local tile counter of this CTA
      if (blockIdx.x + l * gridDim.x >= p.num_tiles) return 0xFFFFFFFFu;
      const uint32_t r = FRS_TILE(l) * kTileM + lg * 32 + lane;
      return r < p.n ? __ldg(p.codes + r) : 0xFFFFFFFFu;
    };
    uint32_t code_a = load_code(0);
    uint32_t code_b = load_code(1);
    uint32_t code_c = load_code(2);

    for (uint32_t slot = blockIdx.x; slot < p.num_tiles; slot += gridDim.x, ++lt) {
      const uint32_t tile = FRS_TILE(lt);
      const uint32_t acc = lt % kAccStages;
      const uint32_t aph = (lt / kAccStages) & 1;
      const uint32_t row = tile * kTileM + lg * 32 + lane;
      const bool live = row < p.n;
      const uint32_t code = code_a;
      code_a = code_b;
      code_b = code_c;
      code_c = load_code(lt + 3);
      const unsigned long long tw0 = p.timeline ? globaltimer_ns() : 0;
      mbar_wait(&tfull[acc], aph);
      if (p.timeline) t_wait += globaltimer_ns() - tw0;
      tc_fence_after();
      uint32_t v[kQW];
      tmem_ld_32xN<kQW>(tmem_base + ((lg * 32u) << 16) + acc * kNQ + e.q0, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);

      if constexpr (DUMP) {
        if (live) {
#pragma unroll
          for (int j = 0; j < kQW; ++j) p.dbg_scores[(size_t)(e.q0 + j) * p.n + row] = __uint_as_float(v[j]);
        }
        continue;
      }

      // fast path: compares against the per-query thresholds + payload predicate
      uint32_t pend = 0;
#pragma unroll
      for (int j = 0; j < kQW; ++j) {
        const int q = e.q0 + j;
        const bool pass = (__uint_as_float(v[j]) >= taua[q]) && (((code ^ qcode[q]) & qmask[q]) == 0u);
        pend |= (uint32_t)pass << j;
      }
      if (!live) pend = 0;

      // rare path: some row of this tile passed some query's threshold
      if (named_bar_or(1, kEpiThreads, pend != 0)) {
        const unsigned long long ts0 = p.timeline ? globaltimer_ns() : 0;
        slow_path<F32>(e, p, v, pend, row);
        if (p.timeline) { t_slow += globaltimer_ns() - ts0; n_slow++; }
      }

      if (e.ew == 0 && lane == 0) {
        if (lt == 0) FRS_TL(3);
        FRS_TL(4);
      }
    }

    if (e.ew == 0 && lane == 0) *reinterpret_cast<volatile uint32_t*>(holder + 1) = 1u;  // stop the refresher
    if (p.timeline && e.ew == 0 && lane == 0) {
      unsigned long long* ts = p.timeline + (size_t)blockIdx.x * 16;
      ts[12] = n_slow; ts[13] = t_wait; ts[14] = t_slow; ts[15] = 0;
    }
    if constexpr (!DUMP) {
      named_bar_sync(1, kEpiThreads);
      // Hand the surviving candidates to the merge kernel: every list entry that is not already
      // ruled out by this CTA's final threshold, unsorted; plus the CTA's best score per query.
      for (int q = e.ew; q < kNQ; q += kEpiWarps) {
        const uint32_t c = min(cnt[q], (uint32_t)kListCap);
        const float t = taua[q];
        const uint64_t e0 = lane < c ? keys[q * kListCap + lane] : 0ull;
        const uint64_t e1 = lane + 32 < c ? keys[q * kListCap + lane + 32] : 0ull;
        const bool k0 = lane < c && key_score(e0) >= t;
        const bool k1 = lane + 32 < c && key_score(e1) >= t;
        const uint32_t b0 = __ballot_sync(0xffffffffu, k0), b1 = __ballot_sync(0xffffffffu, k1);
        uint64_t* dst = p.part_keys + ((size_t)blockIdx.x * kNQ + q) * kListCap;
        const uint32_t below = (1u << lane) - 1u;
        if (k0) dst[__popc(b0 & below)] = e0;
        if (k1) dst[__popc(b0) + __popc(b1 & below)] = e1;
        if (lane == 0) {
          p.part_cnt[blockIdx.x * kNQ + q] = __popc(b0) + __popc(b1);
          __stcg(p.gmax + (size_t)blockIdx.x * kNQ + q, f32_from_ordered(lmax[q]));
        }
      }
      unsigned long long n_app = e.n_app;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n_app += __shfl_xor_sync(0xffffffffu, n_app, o);
      if (lane == 0) {
        atomicAdd(p.stats + kStatAppended, n_app);
        atomicAdd(p.stats + kStatCompactions, e.n_comp);
        atomicAdd(p.stats + kStatResolutions, e.n_res);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
  if (threadIdx.x == 0) FRS_TL(5);
#undef FRS_TL
}

// ---------------------------------------------------------------------------------------------
// phase 2: merge the per-CTA lists of one query, rescore the band exactly, emit top-k
// ---------------------------------------------------------------------------------------------
struct Best {
  uint64_t hi;  // ordered score bits
  uint32_t lo;  // ~row : larger = lower row id
};
__device__ __forceinline__ bool best_less(const Best& a, const Best& b) {
  return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}
__device__ __forceinline__ Best block_max_best(Best v, Best* red /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best w;
    w.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
    w.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
    if (best_less(v, w)) v = w;
  }
  __syncthreads();  // red[] free
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  Best m = red[0];
#pragma unroll
  for (int i = 1; i < kMergeThreads / 32; ++i)
    if (best_less(m, red[i])) m = red[i];
  return m;
}

// The merge kernel keeps survivors / band scores / band rows (20 B per entry) in shared memory when the CTAs'
// lists hold at most kMergeSmemEnt entries for the query in total — the normal case is a few dozen — so that
// the kernel's 32 CTAs fit next to a running scan on the SMs it leaves free (pipelined search).  Fuller lists
// (adversarial: millions of near-duplicates) use the per-query spill area in global memory instead.
constexpr int kMergeSmemEnt = 1536;
constexpr size_t kMergeEntBytes = 8 /*survivor keys*/ + 8 /*band exact*/ + 4 /*band rows*/;
size_t merge_smem_bytes(int) { return (size_t)kMergeSmemEnt * kMergeEntBytes + 64; }
size_t merge_spill_bytes() { return (size_t)kNQ * kGmaxPad * kListCap * kMergeEntBytes; }

// up to 4 exact scores at once (same arithmetic and summation order as exact_dot; the loads of all
// rows are issued before the first use so the DRAM latency is paid once per group)
template <bool F32>
__device__ __forceinline__ void exact_dot4(const void* __restrict__ rows, const uint32_t (&r)[4], int n,
                                           const float* __restrict__ q, double (&out)[4]) {
  const uint32_t lane = lane_id();
  const float4* q4 = reinterpret_cast<const float4*>(q);
  float4 b[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) b[c] = __ldg(q4 + lane + 32 * c);
  if constexpr (F32) {
    float4 a[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4* a4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rows) + (size_t)r[i] * kDim);
#pragma unroll
      for (int c = 0; c < 3; ++c) a[i][c] = i < n ? __ldg(a4 + lane + 32 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc = fma((double)a[i][c].x, (double)b[c].x, acc);
        acc = fma((double)a[i][c].y, (double)b[c].y, acc);
        acc = fma((double)a[i][c].z, (double)b[c].z, acc);
        acc = fma((double)a[i][c].w, (double)b[c].w, acc);
      }
      out[i] = acc;
    }
  } else {
    uint2 a[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint2* a2 = reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(rows) + (size_t)r[i] * kDim);
#pragma unroll
      for (int c = 0; c < 3; ++c) a[i][c] = i < n ? __ldg(a2 + lane + 32 * c) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc = fma((double)__uint_as_float(a[i][c].x << 16), (double)b[c].x, acc);
        acc = fma((double)__uint_as_float(a[i][c].x & 0xFFFF0000u), (double)b[c].y, acc);
        acc = fma((double)__uint_as_float(a[i][c].y << 16), (double)b[c].z, acc);
        acc = fma((double)__uint_as_float(a[i][c].y & 0xFFFF0000u), (double)b[c].w, acc);
      }
      out[i] = acc;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) out[i] += __shfl_xor_sync(0xffffffffu, out[i], o);
  }
}

// local row -> global id (MergeParams): contiguous shard or block-cyclic shard
__device__ __forceinline__ int64_t global_id(const MergeParams& p, uint32_t row) {
  if (p.id_block == 0u) return p.base + (int64_t)row;
  const uint32_t blk = row / p.id_block;
  return ((int64_t)blk * p.id_shards + p.id_shard) * (int64_t)p.id_block + (int64_t)(row - blk * p.id_block);
}

constexpr uint32_t kRankCountMax = 512;  // above this many entries selection falls back to k rounds

// One CTA per query.
//  1. coarse cut: k-th largest of the CTAs' best scores (disjoint CTA groups => distinct rows) minus
//     2*eps; every list entry under it is out.
//  2. exact cut on the survivors: A_k = k-th best pre-filter key, band = entries >= A_k - 2*eps.
//  3. band entries are re-scored in fp64 and ordered by (score desc, row asc).
template <bool F32>
__global__ void __launch_bounds__(kMergeThreads) merge_kernel(const MergeParams p) {
  extern __shared__ __align__(16) uint8_t msm[];
  const int q = blockIdx.x;
  __shared__ Best red[kMergeThreads / 32];
  __shared__ uint32_t cnt_s[kGmaxPad];
  __shared__ uint32_t n_surv, n_band, n_total;
  __shared__ uint64_t ak_s;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const int k = p.k;
  if (tid == 0) { n_surv = 0; n_band = 0; ak_s = 0ull; n_total = 0; }
  __syncthreads();
  if (tid < p.nparts) {
    const uint32_t c = min(p.part_cnt[tid * kNQ + q], (uint32_t)kListCap);
    cnt_s[tid] = c;
    atomicAdd(&n_total, c);
  }
  __syncthreads();
  const bool spill = n_total > (uint32_t)kMergeSmemEnt;
  const int maxent = spill ? kGmaxPad * kListCap : kMergeSmemEnt;
  uint8_t* area = spill ? p.spill + (size_t)q * kGmaxPad * kListCap * kMergeEntBytes : msm;
  uint64_t* surv = reinterpret_cast<uint64_t*>(area);
  double* band_x = reinterpret_cast<double*>(surv + maxent);
  uint32_t* band_row = reinterpret_cast<uint32_t*>(band_x + maxent);

  // 1. coarse cut (computed redundantly by every warp: no block-level exchange needed)
  float thr0 = -INFINITY;
  if (p.nparts >= k) {
    float g[kGmaxPerLane];
#pragma unroll
    for (int i = 0; i < kGmaxPerLane; ++i) {
      const int c = lane + 32 * i;
      g[i] = c < p.nparts ? __ldcg(p.gmax + (size_t)c * kNQ + q) : -INFINITY;
    }
    const float kth = warp_kth_of_lane_max(g, k);
    if (kth > -INFINITY) thr0 = __fsub_rd(kth, 2.0f * p.eps);
  }
  __syncthreads();

  const int total_slots = p.nparts * kListCap;
  for (int i = tid; i < total_slots; i += kMergeThreads) {
    const int part = i / kListCap, slot = i % kListCap;
    if ((uint32_t)slot < cnt_s[part]) {
      const uint64_t key = __ldg(p.part_keys + ((size_t)part * kNQ + q) * kListCap + slot);
      if (key_score(key) >= thr0) surv[atomicAdd(&n_surv, 1u)] = key;
    }
  }
  __syncthreads();
  const uint32_t S = n_surv;

  // 2. k-th best pre-filter key among the survivors
  if (S >= (uint32_t)k) {
    if (S <= kRankCountMax) {
      for (uint32_t i = tid; i < S; i += kMergeThreads) {
        const uint64_t mine = surv[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < S; ++j) rank += surv[j] > mine;
        if (rank == (uint32_t)(k - 1)) ak_s = mine;
      }
    } else {
      uint64_t prev = ~0ull;
      for (int r = 0; r < k; ++r) {
        Best loc{0ull, 0u};
        for (uint32_t i = tid; i < S; i += kMergeThreads) {
          const uint64_t c = surv[i];
          if (c < prev && loc.hi < c) loc.hi = c;
        }
        prev = block_max_best(loc, red).hi;
      }
      if (tid == 0) ak_s = prev;
    }
  }
  __syncthreads();
  const float thr1 = S >= (uint32_t)k ? __fsub_rd(key_score(ak_s), 2.0f * p.eps) : -INFINITY;
  for (uint32_t i = tid; i < S; i += kMergeThreads) {
    const uint64_t key = surv[i];
    if (key_score(key) >= thr1) band_row[atomicAdd(&n_band, 1u)] = key_row(key);
  }
  __syncthreads();
  const uint32_t NB = n_band;

  // 3. exact fp64 scores of the band: each warp takes groups of 4 rows
  const float* qv = p.qrec + (size_t)q * kDim;
  for (uint32_t j0 = warp * 4; j0 < NB; j0 += (kMergeThreads / 32) * 4) {
    uint32_t r[4];
    const int n = (int)min(4u, NB - j0);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = band_row[min(j0 + i, NB - 1)];
    double x[4];
    exact_dot4<F32>(p.rows, r, n, qv, x);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < n) band_x[j0 + i] = x[i];
    }
  }
  __syncthreads();
  if (tid == 0 && p.stats) atomicAdd(p.stats + kStatRescored, (unsigned long long)NB);

  // exact top-k, ordered (score desc, row asc)
  const uint32_t nout = NB < (uint32_t)k ? NB : (uint32_t)k;
  if (NB <= kRankCountMax) {
    for (uint32_t i = tid; i < NB; i += kMergeThreads) {
      const double xi = band_x[i];
      const uint32_t ri = band_row[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < NB; ++j) {
        const double xj = band_x[j];
        rank += (xj > xi) || (xj == xi && band_row[j] < ri);
      }
      if (rank < (uint32_t)k) {
        const size_t o = (size_t)q * k + rank;
        if (p.out_s64) p.out_s64[o] = xi;
        if (p.out_s32) p.out_s32[o] = (float)xi;
        p.out_ids[o] = global_id(p, ri);
      }
    }
  } else {
    Best pb{~0ull, ~0u};
    for (uint32_t r = 0; r < nout; ++r) {
      Best loc{0ull, 0u};
      for (uint32_t i = tid; i < NB; i += kMergeThreads) {
        const Best c{f64_ordered(band_x[i]), ~band_row[i]};
        if (best_less(c, pb) && best_less(loc, c)) loc = c;
      }
      const Best m = block_max_best(loc, red);
      if (tid == 0) {
        const size_t o = (size_t)q * k + r;
        const double sc = f64_from_ordered(m.hi);
        if (p.out_s64) p.out_s64[o] = sc;
        if (p.out_s32) p.out_s32[o] = (float)sc;
        p.out_ids[o] = global_id(p, ~m.lo);
      }
      pb = m;
    }
  }
  for (uint32_t r = nout + tid; r < (uint32_t)k; r += kMergeThreads) {
    const size_t o = (size_t)q * k + r;
    if (p.out_s64) p.out_s64[o] = -INFINITY;
    if (p.out_s32) p.out_s32[o] = -INFINITY;
    p.out_ids[o] = -1;
  }
  if (p.push.peer_gather) {
    // Exchange step fused into this kernel: this query's k (fp64 score, id) words go straight into slot `rank`
    // of every peer's gather buffer over NVLink peer memory (plain stores through IPC-mapped pointers).
    __syncthreads();  // this CTA's results are in p.out_s64 / p.out_ids
    const PushTarget& t = p.push;
    const size_t slot = ((size_t)(t.seq % kExchangeSlots) * t.world + t.rank) * t.block_words;
    const uint64_t* s64 = reinterpret_cast<const uint64_t*>(p.out_s64);
    const uint64_t* ids = reinterpret_cast<const uint64_t*>(p.out_ids);
    for (uint32_t i = tid; i < (uint32_t)(t.n_targets * 2 * k); i += kMergeThreads) {
      const uint32_t peer = i / (2 * k), w = i % (2 * k), pl = w / k, r = w % k;
      const size_t o = (size_t)q * k + r;
      t.peer_gather[peer][slot + pl * t.plane_words + o] = pl ? ids[o] : s64[o];
    }
    __syncthreads();
    __shared__ unsigned int s_last;
    if (tid == 0) {
      __threadfence_system();  // cumulative: the CTA's stores (ordered before it by the barrier) are visible
                               // system-wide before this CTA is counted
      const bool last = atomicAdd(t.counter, 1u) == gridDim.x - 1;
      if (last) __threadfence_system();  // acquire side: the other CTAs' fenced stores precede the flag stores below
      s_last = last;
    }
    __syncthreads();
    if (s_last) {  // every other CTA has fenced its stores and been counted: publish the sequence number
      if (tid == 0) *t.counter = 0;  // next launch (stream-ordered)
      if (t.peer_flags)
        for (uint32_t peer = tid; peer < (uint32_t)t.n_targets; peer += kMergeThreads)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.peer_flags[peer] + t.rank), "r"(t.seq) : "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cross-shard merge: [n_shards, nq, k] exact (fp64 score, global id) -> [nq, k]; one CTA of kShardMergeThreads per
// query.  The n_shards * k candidates of the query are staged in shared memory with ONE round of independent L2 loads
// (ld.global.cg: the words were written by peers over NVLink and sit in this GPU's L2 / HBM), then every thread ranks
// its candidates against the staged list (rank = number of strictly better candidates; ties break by global id).  The
// first version ranked straight from global memory, one warp per query: 4 x 120 dependent-latency loads per lane made
// the kernel take ~70 us at 8 shards — and it sits on the critical chain of the pipelined exchange.
// ---------------------------------------------------------------------------------------------
constexpr int kShardMergeThreads = 128;
constexpr int kShardMergeStage = 1024;  // candidates staged in shared memory (16 KB); more shards fall back to global reads

__device__ __forceinline__ void merge_shards_body(const double* __restrict__ s64, const int64_t* __restrict__ ids,
                                                  int n_shards, int k, size_t shard_stride, float* __restrict__ out_s32,
                                                  int64_t* __restrict__ out_ids, bool poisoned) {
  __shared__ double sh_s[kShardMergeStage];
  __shared__ int64_t sh_id[kShardMergeStage];
  __shared__ int sh_valid;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int total = n_shards * k;
  if (poisoned) {  // the exchange timed out: never hand out a partially gathered result
    for (int r = tid; r < k; r += kShardMergeThreads) {
      out_s32[(size_t)q * k + r] = -INFINITY;
      out_ids[(size_t)q * k + r] = -1;
    }
    return;
  }
  const bool staged = total <= kShardMergeStage;
  if (tid == 0) sh_valid = 0;
  __syncthreads();
  int valid = 0;
  for (int c = tid; c < total; c += kShardMergeThreads) {
    const int sh = c / k, j = c - sh * k;
    const size_t off = (size_t)sh * shard_stride + (size_t)q * k + j;
    const int64_t id = __ldcg(ids + off);
    valid += id >= 0;
    if (staged) {
      sh_id[c] = id;
      sh_s[c] = __ldcg(s64 + off);
    }
  }
  if (valid) atomicAdd(&sh_valid, valid);
  __syncthreads();
  for (int c = tid; c < total; c += kShardMergeThreads) {
    int64_t id;
    double s;
    if (staged) {
      id = sh_id[c];
      s = sh_s[c];
    } else {
      const int sh = c / k, j = c - sh * k;
      const size_t off = (size_t)sh * shard_stride + (size_t)q * k + j;
      id = __ldcg(ids + off);
      s = __ldcg(s64 + off);
    }
    if (id < 0) continue;
    int rank = 0;
    if (staged) {
      for (int d = 0; d < total; ++d) {
        const int64_t id2 = sh_id[d];
        const double s2 = sh_s[d];
        rank += (id2 >= 0) && ((s2 > s) || (s2 == s && id2 < id));
      }
    } else {
      for (int d = 0; d < total; ++d) {
        const int sh2 = d / k, j2 = d - sh2 * k;
        const size_t off2 = (size_t)sh2 * shard_stride + (size_t)q * k + j2;
        const int64_t id2 = __ldcg(ids + off2);
        if (id2 < 0) continue;
        const double s2 = __ldcg(s64 + off2);
        rank += (s2 > s) || (s2 == s && id2 < id);
      }
    }
    if (rank < k) {
      out_s32[(size_t)q * k + rank] = (float)s;
      out_ids[(size_t)q * k + rank] = id;
    }
  }
  // slots beyond the number of valid candidates
  for (int r = sh_valid + tid; r < k; r += kShardMergeThreads) {
    out_s32[(size_t)q * k + r] = -INFINITY;
    out_ids[(size_t)q * k + r] = -1;
  }
}

__global__ void __launch_bounds__(kShardMergeThreads)
merge_shards_kernel(const double* __restrict__ s64, const int64_t* __restrict__ ids, int n_shards, int nq, int k,
                    size_t shard_stride, float* __restrict__ out_s32, int64_t* __restrict__ out_ids,
                    const uint32_t* __restrict__ poison) {
  merge_shards_body(s64, ids, n_shards, k, shard_stride, out_s32, out_ids, poison && __ldcg(poison) != 0u);
}

// The exchange's wait + cross-shard merge in ONE launch (csrc/exchange.cu): in every CTA (one per query) the first warp
// waits until all `world` ranks have published sequence number `seq` in this rank's flag words (lane r polls rank
// r's flag, ld.acquire.sys), then the CTA merges.  The wait is bounded by wall-clock time; on time-out the exchange is
// poisoned (sticky), the sequence number is reported to the host through `status` (pinned, host-mapped) and the query
// comes back empty — nothing traps, the CUDA context and the resident shard survive.
__global__ void __launch_bounds__(kShardMergeThreads)
wait_merge_shards_kernel(const uint32_t* __restrict__ flags, int world, uint32_t seq, unsigned long long timeout_ns,
                         uint32_t* poison, uint32_t* status, const double* __restrict__ s64,
                         const int64_t* __restrict__ ids, int nq, int k, size_t shard_stride,
                         float* __restrict__ out_s32, int64_t* __restrict__ out_ids) {
  bool bad = false;
  if (threadIdx.x < 32) {
    bad = *reinterpret_cast<volatile uint32_t*>(poison) != 0u;
    if (!bad) {
      const unsigned long long t0 = globaltimer_ns();
      for (int r = threadIdx.x; r < world; r += 32) {
        for (;;) {
          uint32_t v;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
          if ((int32_t)(v - seq) >= 0) break;
          __nanosleep(200);
          if (globaltimer_ns() - t0 > timeout_ns || *reinterpret_cast<volatile uint32_t*>(poison) != 0u) {
            *reinterpret_cast<volatile uint32_t*>(poison) = 1u;
            *reinterpret_cast<volatile uint32_t*>(status) = seq;
            __threadfence_system();
            bad = true;
            break;
          }
        }
      }
    }
  }
  // (bar.sync orders the other warps' loads behind the acquiring loads of warp 0 at CTA scope; the data words were
  // released at system scope by the pushing rank before its flag)
  bad = __syncthreads_or(bad) != 0;
  merge_shards_body(s64, ids, world, k, shard_stride, out_s32, out_ids, bad);
}

// ---------------------------------------------------------------------------------------------
// query preparation and row storage (cosine collection: L2-normalise on insert and on query)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Query preparation + bootstrap sample.  kSampleBlocks blocks of 32 warps.  Every block normalises the 32 query
// slots (one warp per slot; slots >= nq are zero) into shared memory; block 0 also writes the MMA operand, the
// fp32 record copy, the predicate copies and resets the per-search tables.  Each block then scores kSampleRows
// sampled rows (a stride over the whole store) against all queries in fp32 and publishes, per group of 16 rows, the
// best MATCHING score per query; the last block to finish turns the 64 group maxima into the scan's starting thresholds:
//   the k-th largest X of the block maxima belongs to k distinct rows with exact score >= X - eps_s, i.e. with
//   pre-filter score >= X - eps_s - eps, so rows under X - eps_s - 3 eps can never reach the top-k.
// The scan kernel starts from tau0 instead of accepting everything, and reads 32 floats instead of ranking the
// sample itself on its ramp.  (A 4096-row sample halves the scan's list appends but does not shorten the launch —
// its CTAs are HBM-bound from the first tiles on — while the longer prep kernel, squeezed onto the few SMs a running
// scan leaves free, became the pipeline's bottleneck at 1.25M rows per GPU: measured, reverted.)
constexpr int kQsStride = kDim + 1;  // +1: lanes read different queries at the same element
constexpr size_t kPrepSmem = ((size_t)kNQ * kQsStride + (size_t)kSampleRows * kDim + (size_t)kSampleRows * 2 * 32) * sizeof(float);
template <bool F32>
__global__ void __launch_bounds__(32 * kNQ) prep_queries_kernel(
    const float* __restrict__ q, const uint32_t* __restrict__ code, const uint32_t* __restrict__ mask, int nq,
    void* __restrict__ qop, float* __restrict__ qrec, uint32_t* __restrict__ qcode, uint32_t* __restrict__ qmask,
    unsigned long long* stats, float* __restrict__ gmax, float* __restrict__ gsample,
    const void* __restrict__ rows, const uint32_t* __restrict__ codes, uint32_t n, int k, float eps) {
  static_assert(kSampleRows == 64 && kNQ == 32, "scoring maps 16 groups of 4 rows x 2 halves to the 32 warps");
  static_assert(kSampleBlocks == 64, "the final selection holds two group maxima per lane");
  static_assert(kSampleGroupRows == 16 && kSampleRows / kSampleGroupRows == 4, "a CTA scores 4 groups of 16 rows");
  constexpr int kGroups = kSampleRows / kSampleGroupRows;
  extern __shared__ __align__(16) float psm[];
  float* rs = psm;                              // [kSampleRows][kDim] sampled rows widened to fp32
  float* qs = rs + kSampleRows * kDim;          // [32][kQsStride] prepared queries
  float* part = qs + kNQ * kQsStride;           // [kSampleRows][2 halves][32 queries] partial dot products
  __shared__ uint32_t smax[kSampleRows / kSampleGroupRows][kNQ], s_code[kNQ], s_mask[kNQ], s_rcode[kSampleRows];
  __shared__ bool s_last;
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool first = blockIdx.x == 0;

  // 1. sampled rows: a stride over the whole store (or the first rows of a small one); 16-byte loads, all issued
  //    before anything depends on them (one DRAM latency per block, overlapped with the normalisation)
  const uint32_t total = kSampleGrid * kSampleRows;
  const uint32_t stride = n >= total ? n / total : 1u;
  constexpr int kVecPerRow = kDim * (F32 ? 4 : 2) / 16;                 // 16-byte vectors per row (96 | 48)
  constexpr int kVecPerThread = kSampleRows * kVecPerRow / (32 * kNQ);  // 6 | 3
  static_assert(kSampleRows * kVecPerRow % (32 * kNQ) == 0, "row vectors must split evenly over the block");
  uint4 rv[kVecPerThread];
#pragma unroll
  for (int e = 0; e < kVecPerThread; ++e) {
    const uint32_t idx = threadIdx.x + 32 * kNQ * e;
    const uint32_t j = idx / kVecPerRow, c = idx - j * kVecPerRow;
    const uint32_t r = (blockIdx.x * kSampleRows + j) * stride;
    rv[e] = r < n ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(rows) + (size_t)r * (kVecPerRow * 16)) + c)
                  : make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x < kSampleRows) {
    const uint32_t r = (blockIdx.x * kSampleRows + threadIdx.x) * stride;
    s_rcode[threadIdx.x] = r < n ? __ldg(codes + r) : 0xFFFFFFFFu;
  }
  if (first) {
    if (threadIdx.x < kStatSlots && stats) stats[threadIdx.x] = 0ull;
    for (int i = threadIdx.x; i < kNQ * kGmaxPad; i += blockDim.x) gmax[i] = -INFINITY;
  }

  // 2. normalise the queries
  float x[12];
  float ss = 0.f;
  if (slot < nq) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      x[i] = q[(size_t)slot * kDim + lane + 32 * i];
      ss = fmaf(x[i], x[i], ss);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = 0.f;
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float y = nrm > 0.f ? x[i] / nrm : 0.f;
    const size_t o = (size_t)slot * kDim + lane + 32 * i;
    float rec;
    if constexpr (F32) {
      rec = y;
      if (first) reinterpret_cast<float*>(qop)[o] = round_tf32(y);
    } else {
      const __nv_bfloat16 b = __float2bfloat16_rn(y);
      rec = __bfloat162float(b);
      if (first) reinterpret_cast<__nv_bfloat16*>(qop)[o] = b;
    }
    if (first) qrec[o] = rec;
    qs[slot * kQsStride + lane + 32 * i] = rec;
  }
  if (lane == 0) {
    const uint32_t c = slot < nq ? code[slot] : 0u, m = slot < nq ? mask[slot] : 0u;
    s_code[slot] = c;
    s_mask[slot] = m;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) smax[g][slot] = f32_ordered(-INFINITY);
    if (first) {
      qcode[slot] = c;
      qmask[slot] = m;
    }
  }
#pragma unroll
  for (int e = 0; e < kVecPerThread; ++e) {
    const uint32_t idx = threadIdx.x + 32 * kNQ * e;
    if constexpr (F32) {
      reinterpret_cast<uint4*>(rs)[idx] = rv[e];
    } else {  // 8 bf16 -> 8 fp32
      float4 lo, hi;
      lo.x = __uint_as_float(rv[e].x << 16); lo.y = __uint_as_float(rv[e].x & 0xFFFF0000u);
      lo.z = __uint_as_float(rv[e].y << 16); lo.w = __uint_as_float(rv[e].y & 0xFFFF0000u);
      hi.x = __uint_as_float(rv[e].z << 16); hi.y = __uint_as_float(rv[e].z & 0xFFFF0000u);
      hi.z = __uint_as_float(rv[e].w << 16); hi.w = __uint_as_float(rv[e].w & 0xFFFF0000u);
      reinterpret_cast<float4*>(rs)[2 * idx] = lo;
      reinterpret_cast<float4*>(rs)[2 * idx + 1] = hi;
    }
  }
  __syncthreads();

  // 3. score: warp -> (4 consecutive rows, one half of the 384 elements), lane -> query.  The phase is bound by
  //    shared-memory wavefronts, so a lane's query elements are read once per FOUR rows (8 wavefronts per 16 FMA;
  //    one (row, half) pair per warp pass took 5 per 4).  Summation order per (row, half) and the order of the final
  //    add are those of the one-pair form: the thresholds are bit-identical.
  {
    const int h = slot & 1, j0 = (slot >> 1) * 4;
    const float* qv = qs + lane * kQsStride + h * (kDim / 2);
    const float4* r4 = reinterpret_cast<const float4*>(rs + j0 * kDim) + h * (kDim / 8);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int i = 0; i < kDim / 8; ++i) {
      const float q0 = qv[4 * i], q1 = qv[4 * i + 1], q2 = qv[4 * i + 2], q3 = qv[4 * i + 3];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 a = r4[r * (kDim / 4) + i];  // same address in every lane: broadcast
        acc[r] = fmaf(a.x, q0, acc[r]);
        acc[r] = fmaf(a.y, q1, acc[r]);
        acc[r] = fmaf(a.z, q2, acc[r]);
        acc[r] = fmaf(a.w, q3, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) part[((j0 + r) * 2 + h) * 32 + lane] = acc[r];
  }
  __syncthreads();
#pragma unroll
  for (int jj = slot; jj < kSampleRows; jj += 32) {  // row jj, lane = query
    const float sc = part[(jj * 2) * 32 + lane] + part[(jj * 2 + 1) * 32 + lane];
    const uint32_t rc = s_rcode[jj];
    if (lane < nq && rc != 0xFFFFFFFFu && ((rc ^ s_code[lane]) & s_mask[lane]) == 0u)
      atomicMax(&smax[jj / kSampleGroupRows][lane], f32_ordered(sc));
  }
  __syncthreads();
  if (threadIdx.x < kGroups * kNQ)
    gsample[(blockIdx.x * kGroups + (threadIdx.x >> 5)) * kNQ + lane] = f32_from_ordered(smax[threadIdx.x >> 5][lane]);

  // 4. the last block turns the block maxima into starting thresholds
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* counter = reinterpret_cast<unsigned int*>(gsample + kSampleCounter);
    const bool last = atomicAdd(counter, 1u) == gridDim.x - 1;
    if (last) {
      *counter = 0u;  // next launch (stream-ordered per workspace)
      __threadfence();
    }
    s_last = last;
  }
  __syncthreads();
  if (!s_last) return;
  {
    // warp `slot` = query `slot`; lane holds block maxima lane and lane + 32; rank by (value desc, block asc)
    const float v0 = __ldcg(gsample + lane * kNQ + slot), v1 = __ldcg(gsample + (lane + 32) * kNQ + slot);
    uint32_t r0 = 0, r1 = 0;
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
      const float a0 = __shfl_sync(0xffffffffu, v0, b), a1 = __shfl_sync(0xffffffffu, v1, b);
      r0 += (a0 > v0) || (a0 == v0 && b < lane);
      r0 += (a1 > v0);
      r1 += (a0 >= v1);
      r1 += (a1 > v1) || (a1 == v1 && b < lane);
    }
    float kth = -INFINITY;
    if (r0 == (uint32_t)(k - 1)) kth = v0;
    if (r1 == (uint32_t)(k - 1)) kth = v1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, o));
    if (lane == 0)
      gsample[kSampleTau0 + slot] = (slot < nq && kth > -INFINITY) ? __fsub_rd(__fsub_rd(kth, kEpsSample), 3.0f * eps) : -INFINITY;
  }
}

// one warp per row
template <bool F32>
__global__ void store_rows_kernel(const float* __restrict__ vecs, const uint32_t* __restrict__ codes,
                                  int64_t n, void* __restrict__ rows_dst, uint32_t* __restrict__ codes_dst) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float x[12];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    x[i] = vecs[row * kDim + lane + 32 * i];
    ss = fmaf(x[i], x[i], ss);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float y = nrm > 0.f ? x[i] / nrm : 0.f;
    const size_t o = (size_t)row * kDim + lane + 32 * i;
    if constexpr (F32) reinterpret_cast<float*>(rows_dst)[o] = y;
    else reinterpret_cast<__nv_bfloat16*>(rows_dst)[o] = __float2bfloat16_rn(y);
  }
  if (lane == 0 && codes_dst) codes_dst[row] = codes ? codes[row] : 0u;
}

template <bool F32>
__global__ void read_rows_kernel(const void* __restrict__ rows, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kDim) return;
  if constexpr (F32) out[i] = reinterpret_cast<const float*>(rows)[i];
  else out[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[i]);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <bool F32, bool DUMP>
static cudaError_t launch_scan_t(int grid, const CUtensorMap& tr, const CUtensorMap& tq,
                                 const ScanParams& p, cudaStream_t st) {
  const size_t smem = scan_smem_bytes(F32);
  static DeviceOnce once;
  cudaError_t e = once_per_device(once, [&] {
    return cudaFuncSetAttribute(scan_kernel<F32, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (e != cudaSuccess) return e;
  scan_kernel<F32, DUMP><<<grid, kScanThreads, smem, st>>>(tr, tq, p);
  return cudaGetLastError();
}

cudaError_t launch_scan(bool f32, bool dump, int grid, const CUtensorMap& tr, const CUtensorMap& tq,
                        const ScanParams& p, cudaStream_t st) {
  if (f32) return dump ? launch_scan_t<true, true>(grid, tr, tq, p, st) : launch_scan_t<true, false>(grid, tr, tq, p, st);
  return dump ? launch_scan_t<false, true>(grid, tr, tq, p, st) : launch_scan_t<false, false>(grid, tr, tq, p, st);
}

cudaError_t launch_merge(bool f32, const MergeParams& p, cudaStream_t st) {
  const size_t smem = merge_smem_bytes(p.nparts);
  // the opt-in limit is set once per device to what the largest grid (kGmaxPad parts) needs
  static DeviceOnce once;
  cudaError_t e = once_per_device(once, [&] {
    const int lim = (int)merge_smem_bytes(kGmaxPad);
    cudaError_t r = cudaFuncSetAttribute(merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    return r != cudaSuccess ? r : cudaFuncSetAttribute(merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
  });
  if (e != cudaSuccess) return e;
  if (f32) merge_kernel<true><<<p.nq, kMergeThreads, smem, st>>>(p);
  else merge_kernel<false><<<p.nq, kMergeThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_merge_shards(const double* s64, const int64_t* ids, int n_shards, int nq, int k,
                                size_t shard_stride, float* out_s32, int64_t* out_ids, cudaStream_t st,
                                const uint32_t* poison) {
  merge_shards_kernel<<<nq, kShardMergeThreads, 0, st>>>(s64, ids, n_shards, nq, k, shard_stride, out_s32, out_ids, poison);
  return cudaGetLastError();
}

cudaError_t launch_wait_merge_shards(const uint32_t* flags, int world, uint32_t seq, unsigned long long timeout_ns,
                                     uint32_t* poison, uint32_t* status, const double* s64, const int64_t* ids, int nq, int k,
                                     size_t shard_stride, float* out_s32, int64_t* out_ids, cudaStream_t st) {
  wait_merge_shards_kernel<<<nq, kShardMergeThreads, 0, st>>>(flags, world, seq, timeout_ns, poison, status, s64, ids, nq, k, shard_stride,
                                              out_s32, out_ids);
  return cudaGetLastError();
}

cudaError_t launch_prep_queries(bool f32, const float* q, const uint32_t* code, const uint32_t* mask,
                                int nq, void* qop, float* qrec, uint32_t* qcode, uint32_t* qmask,
                                unsigned long long* stats, float* gmax, float* gsample, const void* rows,
                                const uint32_t* codes, uint32_t n, int k, float eps, cudaStream_t st) {
  const size_t smem = kPrepSmem;
  static DeviceOnce once;
  cudaError_t e = once_per_device(once, [&] {
    cudaError_t r = cudaFuncSetAttribute(prep_queries_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return r != cudaSuccess ? r : cudaFuncSetAttribute(prep_queries_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (e != cudaSuccess) return e;
  if (f32)
    prep_queries_kernel<true><<<kSampleGrid, 32 * kNQ, smem, st>>>(q, code, mask, nq, qop, qrec, qcode, qmask, stats,
                                                                     gmax, gsample, rows, codes, n, k, eps);
  else
    prep_queries_kernel<false><<<kSampleGrid, 32 * kNQ, smem, st>>>(q, code, mask, nq, qop, qrec, qcode, qmask, stats,
                                                                      gmax, gsample, rows, codes, n, k, eps);
  return cudaGetLastError();
}

// CUDA loads a kernel's code lazily, at its first launch, and that load synchronises with the device.  A first
// launch that happens while an exchange-wait kernel is spinning for a peer (whose own work is queued behind the
// load) would stall until the wait times out — so every kernel of the search path is loaded when an index is created.
cudaError_t preload_search_kernels() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)scan_kernel<false, false>, (const void*)scan_kernel<true, false>,
                       (const void*)scan_kernel<false, true>,  (const void*)scan_kernel<true, true>,
                       (const void*)merge_kernel<false>,       (const void*)merge_kernel<true>,
                       (const void*)merge_shards_kernel,       (const void*)wait_merge_shards_kernel,
                       (const void*)prep_queries_kernel<false>,
                       (const void*)prep_queries_kernel<true>, (const void*)store_rows_kernel<false>,
                       (const void*)store_rows_kernel<true>,   (const void*)read_rows_kernel<false>,
                       (const void*)read_rows_kernel<true>};
  for (const void* f : fns) {
    cudaError_t e = cudaFuncGetAttributes(&a, f);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_store_rows(bool f32, const float* vecs, const uint32_t* codes, int64_t n,
                              void* rows_dst, uint32_t* codes_dst, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int wpb = 8;
  const unsigned grid = (unsigned)((n + wpb - 1) / wpb);
  if (f32) store_rows_kernel<true><<<grid, wpb * 32, 0, st>>>(vecs, codes, n, rows_dst, codes_dst);
  else store_rows_kernel<false><<<grid, wpb * 32, 0, st>>>(vecs, codes, n, rows_dst, codes_dst);
  return cudaGetLastError();
}

cudaError_t launch_read_rows(bool f32, const void* rows_src, int64_t n, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t total = n * kDim;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (f32) read_rows_kernel<true><<<grid, 256, 0, st>>>(rows_src, n, out);
  else read_rows_kernel<false><<<grid, 256, 0, st>>>(rows_src, n, out);
  return cudaGetLastError();
}

}  // namespace frs

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
// Warp roles in scan_kernel (192 threads, one CTA per SM, persistent over tiles b, b+G, ...):
//   warp 0   TMA producer: one 128x128B slab (SWIZZLE_128B) per mbarrier stage, ring of slabs
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (M=128 rows, N=32 queries, K=16|8)
//   warp 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns, payload filter, threshold, list insert
#include "common.cuh"
#include "scan.cuh"

namespace frs {

template <bool F32>
struct ScanCfg {
  static constexpr int kElemBytes = F32 ? 4 : 2;
  static constexpr int kSlabK = 128 / kElemBytes;       // elements per 128 B slab row (32 | 64)
  static constexpr int kSlabs = kDim / kSlabK;          // slabs per tile (12 | 6)
  static constexpr int kRing = F32 ? 9 : 10;            // slabs in flight
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
  static constexpr size_t scratch = keys + (size_t)kNQ * kListCap * 8;   // f64 [4][kListCap]
  static constexpr size_t cnt = scratch + 4 * kListCap * 8;              // u32 [32]
  static constexpr size_t taua = cnt + kNQ * 4;                          // f32 [32]
  static constexpr size_t qcode = taua + kNQ * 4;
  static constexpr size_t qmask = qcode + kNQ * 4;
  static constexpr size_t bars = qmask + kNQ * 4;  // full[R] empty[R] tfull[A] tempty[A] qbar
  static constexpr size_t nbars = 2 * C::kRing + 2 * kAccStages + 1;
  static constexpr size_t holder = bars + nbars * 8;
  static constexpr size_t total = holder + 16;
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

// ---- list compaction (one warp, one query) ----------------------------------------------------
// Keeps every entry that can still be in the exact top-k of the rows this CTA has seen:
// all entries with pre-filter score >= A_k - 2*eps (A_k = k-th best pre-filter score in the list).
// If more than kKeep entries sit in that band the band is resolved exactly (fp64) and only the
// exact top-k stay.  Raises the pass threshold taua[q] accordingly.
template <bool F32>
__device__ __noinline__ void compact_list(uint64_t* __restrict__ L, uint32_t* cnt_q, float* taua_q,
                                          double* __restrict__ scratch, const ScanParams& p, int q,
                                          unsigned long long& n_resolutions) {
  const uint32_t lane = lane_id();
  const uint32_t c = min(*cnt_q, (uint32_t)kListCap);
  const int k = p.k;
  __syncwarp();
  const uint64_t e0 = lane < c ? L[lane] : 0ull;
  const uint64_t e1 = lane + 32 < c ? L[lane + 32] : 0ull;
  uint32_t r0 = 0, r1 = 0;
  for (uint32_t j = 0; j < c; ++j) {
    const uint64_t kj = L[j];
    r0 += kj > e0;
    r1 += kj > e1;
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
      if (cutoff > *taua_q) *taua_q = cutoff;
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
    if (t > *taua_q) *taua_q = t;
  }
  __syncwarp();
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
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = sm + S::ring;
  uint8_t* qop = sm + S::qop;
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm + S::keys);
  double* scratch = reinterpret_cast<double*>(sm + S::scratch);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(sm + S::cnt);
  float* taua = reinterpret_cast<float*>(sm + S::taua);
  uint32_t* qcode = reinterpret_cast<uint32_t*>(sm + S::qcode);
  uint32_t* qmask = reinterpret_cast<uint32_t*>(sm + S::qmask);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::bars);
  uint64_t* empty = full + C::kRing;
  uint64_t* tfull = empty + C::kRing;
  uint64_t* tempty = tfull + kAccStages;
  uint64_t* qbar = tempty + kAccStages;
  uint32_t* holder = reinterpret_cast<uint32_t*>(sm + S::holder);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kRing; ++i) {
      mbar_init(&full[i], 1);   // producer's arrive.expect_tx (+ TMA complete_tx bytes)
      mbar_init(&empty[i], 1);  // tcgen05.commit from the MMA thread
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull[i], 1);   // tcgen05.commit
      mbar_init(&tempty[i], 4);  // one arrive per epilogue warp
    }
    mbar_init(qbar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < kNQ) {
    const int t = threadIdx.x;
    cnt[t] = 0;
    taua[t] = t < p.nq ? -INFINITY : INFINITY;  // padded queries never pass
    qcode[t] = t < p.nq ? p.qcode[t] : 0u;
    qmask[t] = t < p.nq ? p.qmask[t] : 0u;
  }
  if (warp == 1) {
    tmem_alloc(holder, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_rows);
      tma_prefetch_desc(&tmap_q);
      mbar_arrive_expect_tx(qbar, C::kQBytes);
      for (int s = 0; s < C::kSlabs; ++s)
        tma_load_2d(qop + s * C::kQSlabBytes, &tmap_q, qbar, s * C::kSlabK, 0, kEvictLast);
      uint32_t it = 0;
      for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
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
          const uint64_t da = make_desc_sw128(smem_u32(ring + (size_t)stage * kSlabBytes));
          const uint64_t db = make_desc_sw128(smem_u32(qop + s * C::kQSlabBytes));
#pragma unroll
          for (int kk = 0; kk < C::kMmasPerSlab; ++kk)
            tc_mma<F32>(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (uint32_t)((s | kk) != 0));
          tc_commit(&empty[stage]);  // slab free once these MMAs have read it
        }
        tc_commit(&tfull[acc]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue =====================
    const uint32_t ew = warp - 2;   // 0..3: which queries this warp compacts
    const uint32_t lg = warp & 3;   // TMEM lane group this warp may read
    unsigned long long n_app = 0, n_comp = 0, n_res = 0;
    uint32_t lt = 0;
    for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      const uint32_t acc = lt % kAccStages;
      const uint32_t aph = (lt / kAccStages) & 1;
      const uint32_t row = tile * kTileM + lg * 32 + lane;
      const bool live = row < p.n;
      const uint32_t code = live ? __ldg(p.codes + row) : 0xFFFFFFFFu;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((lg * 32u) << 16) + acc * kNQ, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);

      if constexpr (DUMP) {
        if (live) {
#pragma unroll
          for (int q = 0; q < kNQ; ++q) p.dbg_scores[(size_t)q * p.n + row] = __uint_as_float(v[q]);
        }
        continue;
      }

      uint32_t pend = 0;
#pragma unroll
      for (int q = 0; q < kNQ; ++q) {
        const bool pass = (__uint_as_float(v[q]) >= taua[q]) && (((code ^ qcode[q]) & qmask[q]) == 0u);
        pend |= (uint32_t)pass << q;
      }
      if (!live) pend = 0;

      // Rare path: some row of this tile passed some query's threshold.
      while (named_bar_or(1, 128, pend != 0)) {
#pragma unroll
        for (int q = 0; q < kNQ; ++q) {
          if (pend & (1u << q)) {
            const uint32_t slot = atomicAdd(&cnt[q], 1u);
            if (slot < (uint32_t)kListCap) {
              keys[q * kListCap + slot] = make_key(__uint_as_float(v[q]), row);
              pend &= ~(1u << q);
              n_app++;
            }
          }
        }
        named_bar_sync(1, 128);
        for (int q = ew; q < kNQ; q += 4) {
          if (cnt[q] >= (uint32_t)kListCap) {
            compact_list<F32>(keys + q * kListCap, &cnt[q], &taua[q], scratch + ew * kListCap, p, q, n_res);
            n_comp++;
          }
        }
        named_bar_sync(1, 128);
        if (pend) {
#pragma unroll
          for (int q = 0; q < kNQ; ++q)
            if ((pend & (1u << q)) && !(__uint_as_float(v[q]) >= taua[q])) pend &= ~(1u << q);
        }
      }
    }

    if constexpr (!DUMP) {
      named_bar_sync(1, 128);
      for (int q = ew; q < kNQ; q += 4) {
        compact_list<F32>(keys + q * kListCap, &cnt[q], &taua[q], scratch + ew * kListCap, p, q, n_res);
        const uint32_t c = cnt[q];
        uint64_t* dst = p.part_keys + ((size_t)blockIdx.x * kNQ + q) * kKeep;
        if (lane < c) dst[lane] = keys[q * kListCap + lane];
        if (lane == 0) p.part_cnt[blockIdx.x * kNQ + q] = c;
      }
      // per-warp counters -> global (lane sums for n_app, lane 0 for the warp-uniform ones)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n_app += __shfl_xor_sync(0xffffffffu, n_app, o);
      if (lane == 0) {
        atomicAdd(p.stats + kStatAppended, n_app);
        atomicAdd(p.stats + kStatCompactions, n_comp);
        atomicAdd(p.stats + kStatResolutions, n_res);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
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

size_t merge_smem_bytes(int nparts) {
  const size_t ent = (size_t)nparts * kKeep;
  return ent * 8 /*keys*/ + ent * 8 /*band exact*/ + ent * 4 /*band rows*/ + 256;
}

template <bool F32>
__global__ void __launch_bounds__(kMergeThreads) merge_kernel(const MergeParams p) {
  extern __shared__ __align__(16) uint8_t msm[];
  const int q = blockIdx.x;
  const int maxent = p.nparts * kKeep;
  uint64_t* ent = reinterpret_cast<uint64_t*>(msm);
  double* band_x = reinterpret_cast<double*>(ent + maxent);
  uint32_t* band_row = reinterpret_cast<uint32_t*>(band_x + maxent);
  __shared__ Best red[kMergeThreads / 32];
  __shared__ uint32_t n_ent, n_band;
  const int tid = threadIdx.x;
  const int k = p.k;
  if (tid == 0) { n_ent = 0; n_band = 0; }
  __syncthreads();

  // gather the surviving entries of every CTA for this query
  for (int i = tid; i < maxent; i += kMergeThreads) {
    const int part = i / kKeep, j = i - part * kKeep;
    if ((uint32_t)j < p.part_cnt[part * kNQ + q]) {
      const uint64_t key = p.part_keys[((size_t)part * kNQ + q) * kKeep + j];
      ent[atomicAdd(&n_ent, 1u)] = key;
    }
  }
  __syncthreads();
  const uint32_t T = n_ent;

  // k-th best pre-filter key: k rounds of "largest key below the previous winner"
  Best prev{~0ull, ~0u};
  uint64_t ak = 0;
  for (int r = 0; r < k; ++r) {
    Best loc{0ull, 0u};
    for (uint32_t i = tid; i < T; i += kMergeThreads) {
      const Best c{ent[i], 0u};
      if (c.hi < prev.hi && loc.hi < c.hi) loc = c;
    }
    const Best m = block_max_best(loc, red);
    if (m.hi == 0ull) break;
    prev = m;
    ak = m.hi;
  }
  const float cutoff = (T >= (uint32_t)k) ? __fsub_rd(key_score(ak), 2.0f * p.eps) : -INFINITY;

  // the band: everything that may still be in the exact top-k
  for (uint32_t i = tid; i < T; i += kMergeThreads) {
    const uint64_t key = ent[i];
    if (key_score(key) >= cutoff) band_row[atomicAdd(&n_band, 1u)] = key_row(key);
  }
  __syncthreads();
  const uint32_t NB = n_band;

  // exact fp64 scores of the band, one warp per entry
  const float* qv = p.qrec + (size_t)q * kDim;
  for (uint32_t j = tid >> 5; j < NB; j += kMergeThreads / 32) {
    const double ex = exact_dot<F32>(p.rows, band_row[j], qv);
    if ((tid & 31) == 0) band_x[j] = ex;
  }
  __syncthreads();
  if (tid == 0 && p.stats) atomicAdd(p.stats + kStatRescored, (unsigned long long)NB);

  // exact top-k, ordered (score desc, row asc)
  Best pb{~0ull, ~0u};
  bool exhausted = false;
  for (int r = 0; r < k; ++r) {
    Best loc{0ull, 0u};
    if (!exhausted) {
      for (uint32_t i = tid; i < NB; i += kMergeThreads) {
        const Best c{f64_ordered(band_x[i]), ~band_row[i]};
        if (best_less(c, pb) && best_less(loc, c)) loc = c;
      }
    }
    const Best m = block_max_best(loc, red);
    if (m.hi == 0ull && m.lo == 0u) exhausted = true;
    if (tid == 0) {
      const size_t o = (size_t)q * k + r;
      if (!exhausted) {
        const double s = f64_from_ordered(m.hi);
        if (p.out_s64) p.out_s64[o] = s;
        if (p.out_s32) p.out_s32[o] = (float)s;
        p.out_ids[o] = p.base + (int64_t)(~m.lo);
      } else {
        if (p.out_s64) p.out_s64[o] = -INFINITY;
        if (p.out_s32) p.out_s32[o] = -INFINITY;
        p.out_ids[o] = -1;
      }
    }
    pb = m;
  }
}

// ---------------------------------------------------------------------------------------------
// cross-shard merge: [n_shards, nq, k] exact (fp64 score, global id) -> [nq, k]; one warp per query
// ---------------------------------------------------------------------------------------------
__global__ void merge_shards_kernel(const double* __restrict__ s64, const int64_t* __restrict__ ids,
                                    int n_shards, int nq, int k, float* __restrict__ out_s32,
                                    int64_t* __restrict__ out_ids) {
  const int q = blockIdx.x;
  const int lane = threadIdx.x;
  const int total = n_shards * k;
  // each lane owns candidates lane, lane+32, ... ; rank = number of strictly better candidates
  for (int c = lane; c < total; c += 32) {
    const int sh = c / k, j = c - sh * k;
    const size_t off = ((size_t)sh * nq + q) * k + j;
    const int64_t id = ids[off];
    if (id < 0) continue;
    const double s = s64[off];
    int rank = 0;
    for (int d = 0; d < total; ++d) {
      const int sh2 = d / k, j2 = d - sh2 * k;
      const size_t off2 = ((size_t)sh2 * nq + q) * k + j2;
      const int64_t id2 = ids[off2];
      if (id2 < 0) continue;
      const double s2 = s64[off2];
      rank += (s2 > s) || (s2 == s && id2 < id);
    }
    if (rank < k) {
      out_s32[(size_t)q * k + rank] = (float)s;
      out_ids[(size_t)q * k + rank] = id;
    }
  }
  // slots beyond the number of valid candidates
  int valid = 0;
  for (int c = lane; c < total; c += 32) {
    const int sh = c / k, j = c - sh * k;
    valid += ids[((size_t)sh * nq + q) * k + j] >= 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
  for (int r = valid + lane; r < k; r += 32) {
    out_s32[(size_t)q * k + r] = -INFINITY;
    out_ids[(size_t)q * k + r] = -1;
  }
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

// one warp per query slot (32 slots); slots >= nq are zero filled
template <bool F32>
__global__ void __launch_bounds__(32 * kNQ) prep_queries_kernel(const float* __restrict__ q, const uint32_t* __restrict__ code,
                                    const uint32_t* __restrict__ mask, int nq, void* __restrict__ qop,
                                    float* __restrict__ qrec, uint32_t* __restrict__ qcode,
                                    uint32_t* __restrict__ qmask, unsigned long long* stats) {
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kStatSlots && stats) stats[threadIdx.x] = 0ull;
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
    float y = nrm > 0.f ? x[i] / nrm : 0.f;
    const size_t o = (size_t)slot * kDim + lane + 32 * i;
    if constexpr (F32) {
      qrec[o] = y;
      reinterpret_cast<float*>(qop)[o] = round_tf32(y);
    } else {
      const __nv_bfloat16 b = __float2bfloat16_rn(y);
      reinterpret_cast<__nv_bfloat16*>(qop)[o] = b;
      qrec[o] = __bfloat162float(b);
    }
  }
  if (lane == 0) {
    qcode[slot] = slot < nq ? code[slot] : 0u;
    qmask[slot] = slot < nq ? mask[slot] : 0u;
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
  if (lane == 0) codes_dst[row] = codes ? codes[row] : 0u;
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
  static bool configured[64] = {};  // per device: the attribute lives in the device's context
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(scan_kernel<F32, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
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
  cudaError_t e;
  if (f32) {
    e = cudaFuncSetAttribute(merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    merge_kernel<true><<<p.nq, kMergeThreads, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    merge_kernel<false><<<p.nq, kMergeThreads, smem, st>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_merge_shards(const double* s64, const int64_t* ids, int n_shards, int nq, int k,
                                float* out_s32, int64_t* out_ids, cudaStream_t st) {
  merge_shards_kernel<<<nq, 32, 0, st>>>(s64, ids, n_shards, nq, k, out_s32, out_ids);
  return cudaGetLastError();
}

cudaError_t launch_prep_queries(bool f32, const float* q, const uint32_t* code, const uint32_t* mask,
                                int nq, void* qop, float* qrec, uint32_t* qcode, uint32_t* qmask,
                                unsigned long long* stats, cudaStream_t st) {
  if (f32) prep_queries_kernel<true><<<1, 32 * kNQ, 0, st>>>(q, code, mask, nq, qop, qrec, qcode, qmask, stats);
  else prep_queries_kernel<false><<<1, 32 * kNQ, 0, st>>>(q, code, mask, nq, qop, qrec, qcode, qmask, stats);
  return cudaGetLastError();
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

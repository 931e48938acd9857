// scan.cuh — launch interface of the exact cosine top-k scan (kernels in scan.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace frs {

constexpr int kDim = 384;       // VectorParams(size=384)            reference ingest.py:89-95
constexpr int kNQ = 32;         // queries per pass == MMA N         reference main2.py:51
constexpr int kTileM = 128;     // corpus rows per tile == MMA M
constexpr int kSlabBytes = kTileM * 128;  // 128 rows x 128 B: one SWIZZLE_128B K-slab (16 KiB)
constexpr int kListCap = 64;    // per-query candidate list capacity inside a CTA
constexpr int kKeep = 32;       // most entries a list keeps after a compaction
constexpr int kMaxK = 32;       // largest `limit` supported (the reference asks for 15, main.py:215; 5, evaluate.py:86)
constexpr int kAccStages = 16;  // TMEM accumulator ring (16 x 32 columns = all 512)
constexpr int kTmemCols = kAccStages * kNQ;
constexpr int kEpiWarps = 8;      // epilogue warps: 4 TMEM lane groups x 2 query-column groups
constexpr int kQW = kNQ / (kEpiWarps / 4);  // query columns per epilogue warp (16)
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kScanThreads = 64 + kEpiThreads + 32;  // warp0 TMA, warp1 MMA, epilogue warps, bound refresher
constexpr int kMergeThreads = 256;
constexpr int kGmaxPerLane = 5;                 // cross-CTA bound table: up to 160 CTAs
constexpr int kGmaxPad = 32 * kGmaxPerLane;     // floats per query
constexpr int kMaxTileSlots = 944;              // tiles per CTA of a restricted scan (what shared memory leaves)

// bootstrap sample scored by the prep kernel: kSampleBlocks blocks x kSampleRows rows
constexpr int kSampleBlocks = 64;  // groups of sampled rows; each publishes its best matching score per query
constexpr int kSampleGrid = 16;    // CTAs of the prep kernel: few and fat, so that it is short even when it is
                                   // squeezed onto the handful of SMs a running scan leaves free
constexpr int kSampleRows = 64;    // rows per CTA = 4 groups of 16 (1024 rows in all)
constexpr int kSampleGroupRows = kSampleRows * kSampleGrid / kSampleBlocks;
// gsample buffer: [kSampleBlocks][32] block maxima | tau0 [32] starting thresholds | block counter (u32)
constexpr int kSampleTau0 = kSampleBlocks * kNQ;
constexpr int kSampleCounter = kSampleTau0 + kNQ;
constexpr int kSampleFloats = kSampleCounter + 1;
constexpr float kEpsSample = 3.0e-5f;  // |fp32 FMA-chain score - fp64 score| bound (384 terms)

// bound on |tensor-core pre-filter score - fp64 score| for unit-norm rows and queries
constexpr float kEpsBF16 = 3.0e-5f;  // exact products, fp32 accumulation of 384 terms
constexpr float kEpsTF32 = 2.0e-3f;  // 2^-9 operand truncation (Cauchy-Schwarz) + accumulation

enum StatSlot { kStatAppended = 0, kStatCompactions = 1, kStatResolutions = 2, kStatRescored = 3, kStatSlots = 8 };

struct ScanParams {
  const void* rows;        // [capacity, 384] storage dtype
  const uint32_t* codes;   // [capacity] payload codes
  uint32_t n;              // rows in use
  uint32_t num_tiles;      // tiles to scan: ceil(n / 128), or the length of tile_ids
  const uint32_t* tile_ids;  // null = every tile; else ascending 128-row tile numbers (restricted scan)
  const float* qrec;       // [32, 384] prepared queries widened to fp32 (exact rescoring operand)
  const uint32_t* qcode;   // [32]
  const uint32_t* qmask;   // [32]
  int nq;                  // live queries (<= 32)
  int k;                   // top-k (<= kMaxK)
  float eps;               // pre-filter error bound
  uint64_t* part_keys;     // [grid, 32, kListCap] surviving (approx score, row) keys per CTA, unsorted
  uint32_t* part_cnt;      // [grid, 32]
  float* gmax;             // [kGmaxPad, 32] best appended pre-filter score per (CTA, query)
  const float* tau0;       // [32] starting pass thresholds from the prep kernel's bootstrap sample (-inf = none)
  float* dbg_scores;       // DUMP mode only: [32, n]
  unsigned long long* stats;  // [kStatSlots]
  unsigned long long* timeline;  // diagnostics: [grid, 8] globaltimer stamps, or null
};

// Peer-memory push fused into the tail of the merge kernel (csrc/exchange.cu): every query's CTA copies its k
// (score, id) words into slot `rank` of every peer's gather buffer; the last CTA to finish publishes `seq` in
// every peer's flag word.
// Gather buffers are a ring of kExchangeSlots batches.  Rule for the callers: a rank issues the push of batch t + d
// only after its own final merge of batch t.  A peer's push of s + D lands on slot s; it comes after the peer's merge
// of s + D - d, which needed this rank's push of s + D - d, issued after this rank's merge of s + D - 2d: D >= 2d keeps
// a slot from being overwritten while it is read.  d = 2 (one batch of overlap between a local pass and the previous
// batch's wait + merge on a side stream) needs D = 4.
constexpr uint32_t kExchangeSlots = 4;

struct PushTarget {
  uint64_t* const* peer_gather;  // device array [n_targets] of gather buffers, null = no push
  uint32_t* const* peer_flags;   // device array [n_targets] of flag arrays, or null: no flags are published (one process
                                 // driving several GPUs orders the consumer with CUDA events instead, csrc/sharded.cu)
  unsigned int* counter;         // CTAs done (local, self-resetting)
  int n_targets;                 // how many buffers receive this rank's block (world, or 1 = only the collecting GPU)
  int world, rank;               // the block lands in slot `rank` of `world`
  uint32_t seq;
  size_t block_words;            // 2 * nq_max * k_max
  size_t plane_words;            // nq_max * k_max: offset of the id plane inside a block
};

struct MergeParams {
  const uint64_t* part_keys;
  const uint32_t* part_cnt;
  const float* gmax;   // [kGmaxPad, 32] final best pre-filter score per (CTA, query)
  int nparts;
  const void* rows;
  const float* qrec;
  int nq;
  int k;
  float eps;
  int64_t base;        // global id of local row 0 (contiguous shards)
  // block-cyclic shards (csrc/sharded.cu): local row r is global row ((r / id_block) * id_shards + id_shard) * id_block
  // + r % id_block; id_block == 0 selects base + r.  Monotone in r, so local ties (row asc) are global ties (id asc).
  uint32_t id_block;
  uint32_t id_shards, id_shard;
  double* out_s64;     // [nq, k] or null
  float* out_s32;      // [nq, k] or null
  int64_t* out_ids;    // [nq, k]
  unsigned long long* stats;
  uint8_t* spill;      // [merge_spill_bytes()] global scratch for queries whose lists exceed the shared-memory area
  PushTarget push;
};

cudaError_t preload_search_kernels();
size_t scan_smem_bytes(bool f32);
size_t merge_smem_bytes(int nparts);
size_t merge_spill_bytes();

// all launchers return the cudaError_t of configuration + launch
cudaError_t launch_scan(bool f32, bool dump, int grid, const CUtensorMap& tmap_rows,
                        const CUtensorMap& tmap_q, const ScanParams& p, cudaStream_t st);
cudaError_t launch_merge(bool f32, const MergeParams& p, cudaStream_t st);
// shard_stride: elements between consecutive shards in both arrays (nq*k when they are dense)
// poison: optional device word; non-zero = the exchange that filled the buffer timed out -> every slot is emitted empty
cudaError_t launch_merge_shards(const double* s64, const int64_t* ids, int n_shards, int nq, int k,
                                size_t shard_stride, float* out_s32, int64_t* out_ids, cudaStream_t st,
                                const uint32_t* poison = nullptr);
// the exchange's bounded wait for all ranks' sequence flags fused with the cross-shard merge (one launch per batch)
cudaError_t launch_wait_merge_shards(const uint32_t* flags, int world, uint32_t seq, unsigned long long timeout_ns,
                                     uint32_t* poison, uint32_t* status, const double* s64, const int64_t* ids, int nq, int k,
                                     size_t shard_stride, float* out_s32, int64_t* out_ids, cudaStream_t st);
// queries [nq,384] fp32 -> qop (MMA operand, bf16 or tf32-rounded fp32, zero padded to 32 rows),
// qrec (fp32 record copy), qcode/qmask copies padded to 32
// also scores a strided sample of the stored rows against the prepared queries (bootstrap bound)
cudaError_t launch_prep_queries(bool f32, const float* q, const uint32_t* code, const uint32_t* mask,
                                int nq, void* qop, float* qrec, uint32_t* qcode, uint32_t* qmask,
                                unsigned long long* stats, float* gmax, float* gsample, const void* rows,
                                const uint32_t* codes, uint32_t n, int k, float eps, cudaStream_t st);
// rows [n,384] fp32 -> L2-normalised storage rows (+ codes) at dst row offset
cudaError_t launch_store_rows(bool f32, const float* vecs, const uint32_t* codes, int64_t n,
                              void* rows_dst, uint32_t* codes_dst, cudaStream_t st);
cudaError_t launch_read_rows(bool f32, const void* rows_src, int64_t n, float* out, cudaStream_t st);

}  // namespace frs

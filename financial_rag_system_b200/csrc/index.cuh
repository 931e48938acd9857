// index.cuh — the chunk-store object behind the frs_index_* ABI, shared by index.cu (one shard), exchange.cu
// (multi-process exchange) and sharded.cu (one process driving several GPUs).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <condition_variable>
#include <mutex>

#include "../../include/frs_b200.h"
#include "exchange.cuh"
#include "scan.cuh"

namespace frs {

// Everything one search owns between its prep kernel and its merge kernel.  An index keeps a small ring of them
// so that the prep of batch i+1 and the merge of batch i-1 can run while batch i is being scanned.
struct SearchWs {
  void* qop = nullptr;        // [32,384] MMA B operand (bf16 | tf32-rounded fp32)
  float* qrec = nullptr;      // [32,384] prepared queries widened to fp32 (exact rescoring operand)
  uint32_t* qcode = nullptr;  // [32]
  uint32_t* qmask = nullptr;  // [32]
  uint64_t* part_keys = nullptr;
  uint32_t* part_cnt = nullptr;
  unsigned long long* stats = nullptr;
  float* gmax = nullptr;
  float* gsample = nullptr;
  uint8_t* spill = nullptr;   // merge kernel scratch for over-full lists
  CUtensorMap tmap_q;
  cudaEvent_t free = nullptr;  // recorded after the merge kernel that last used this workspace
};

// Staging of one *_host call: one pinned + one device buffer each way, so a call is ONE H2D and ONE D2H copy.
//   in : queries [32*384] f32 | code [32] u32 | mask [32] u32
//   out: ids [32*16] i64 | scores [32*16] f32
struct HostSlot {
  uint8_t* h_in = nullptr;
  uint8_t* d_in = nullptr;
  uint8_t* h_out = nullptr;
  uint8_t* d_out = nullptr;
  cudaEvent_t done = nullptr;
  bool busy = false;
  int nq = 0, k = 0;
};
constexpr size_t kHostInBytes = (size_t)kNQ * kDim * 4 + 2 * kNQ * 4;
constexpr size_t kHostOutBytes = (size_t)kNQ * kMaxK * (8 + 4);
constexpr int kWsRing = 4;      // searches in flight per index (prep | scan | scan | merge)
constexpr int kHostSlots = 4;   // host calls in flight per index
constexpr int kJobRing = 16;    // event sets of the pipelined form
constexpr int kProfEvents = 8;  // per recorded search: pre/post prep, pre/post scan, pre/post merge, post exchange

struct SearchArgs {
  const float* q = nullptr;
  const uint32_t* code = nullptr;
  const uint32_t* mask = nullptr;
  int nq = 0, k = 0;
  float* out_s32 = nullptr;
  double* out_s64 = nullptr;
  int64_t* out_ids = nullptr;
  const uint32_t* tile_ids = nullptr;
  int64_t n_tile_ids = 0;
  const PushTarget* push = nullptr;  // exchange step fused into the merge kernel
};

// Which streams a search runs on.  in-stream: every kernel on `st`.  pipelined: prep / scan / merge on the index's
// three internal streams; `st` (may be null) only supplies the "inputs are ready" ordering.
struct SearchLaunch {
  cudaStream_t prep = nullptr, scan = nullptr, merge = nullptr;
  cudaEvent_t ev_prep = nullptr, ev_scan = nullptr;  // hand-over events between the streams (pipelined form)
  cudaEvent_t* prof = nullptr;  // kProfEvents events, or null
  int reserve_sms = 0;          // SMs the persistent scan leaves to the other streams' kernels (pipelined form)
};

int abi_set_err(int code, const char* fmt, ...);
int abi_make_tmap_bf16(CUtensorMap* m, void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows);

}  // namespace frs

struct frs_index {
  int device = 0;
  int dtype = FRS_DTYPE_BF16;
  int64_t capacity = 0;
  int64_t size = 0;
  int64_t base = 0;
  uint32_t id_block = 0, id_shards = 1, id_shard = 0;  // block-cyclic id mapping (sharded.cu), 0 = base + row
  int sm_count = 0;
  int grid_override = 0;
  int pipe_reserve = -1;  // see frs_index_set_pipeline_reserve (-1 = chosen from the shard size)
  void* rows = nullptr;
  uint32_t* codes = nullptr;
  CUtensorMap tmap_rows;
  frs::SearchWs ws[frs::kWsRing];
  int ws_next = 0;
  int ws_last = 0;
  int max_parts = 0;
  // internal streams of the pipelined form + the write path
  // Consecutive scans alternate between TWO streams: they are independent kernels (separate workspaces), so the
  // CTAs of scan i+1 start on SMs as the CTAs of scan i leave them — the ramp of one launch fills the tail of the
  // previous one instead of both being bubbles (at 1.25M rows per GPU they were ~15 % of a launch).
  // s_xchg: the cross-shard wait + merge of a sharded search.  Not on s_merge: the wait for the slowest peer's push of
  // batch i would hold up this rank's LOCAL merge of batch i+1 behind it, the workspace ring would fill and the scans
  // stall (8 GPUs, 1.25M rows each: the in-order chain merge -> wait -> merge took 100-300 us per batch, the scan 150).
  cudaStream_t s_prep = nullptr, s_scan[2] = {nullptr, nullptr}, s_merge = nullptr, s_xchg = nullptr;
  int scan_streams = 2;
  cudaStream_t stream = nullptr;  // host-call write path (add_host, read_rows_host)
  cudaEvent_t job_in[frs::kJobRing], job_prep[frs::kJobRing], job_scan[frs::kJobRing], job_done[frs::kJobRing];
  cudaEvent_t job_merge[frs::kJobRing];  // local merge (+ push) enqueued: hand-over s_merge -> s_xchg
  cudaEvent_t xchg_last = nullptr;       // recorded on s_xchg after the newest sharded job
  bool xchg_used = false;
  uint64_t jobs = 0;
  cudaEvent_t rows_ready = nullptr;  // recorded after the last write-path kernel; every search waits on it
  bool writes_pending = false;       // a write was issued and rows_ready has not been seen complete yet
  // profiling (off by default)
  static constexpr int kProfRing = 256;
  int prof_mode = 0;   // 0 off, 1 events, 2 events + in-kernel timeline
  int prof_calls = 0;  // searches recorded since the last read
  cudaEvent_t* prof_ev = nullptr;  // [kProfRing][kProfEvents]
  unsigned long long* timeline = nullptr;
  // mode 3 (bracket): ONE event before the first scan and one after the latest scan, on the scan's stream — the
  // average scan launch duration (+ the gaps between launches) without events between the kernels of a step
  cudaEvent_t br_first = nullptr, br_last = nullptr;
  int br_count = 0;
  // host-call staging
  frs::HostSlot hs[frs::kHostSlots];
  std::mutex hs_mu;
  std::condition_variable hs_cv;
  std::mutex mu;  // serialises enqueueing (stream order == call order) and the size / ring bookkeeping
  int last_grid = 0;
  int last_launches = 0;
  bool f32() const { return dtype == FRS_DTYPE_F32; }
  size_t row_bytes() const { return (size_t)frs::kDim * (f32() ? 4 : 2); }
};

namespace frs {
// Enqueues prep -> scan -> merge of one search on the streams of `L` (caller holds ix->mu and has set the device).
// Every stream of L is ordered behind the index's pending writes and behind the workspace's previous user.
int search_enqueue(frs_index* ix, const SearchArgs& a, const SearchLaunch& L);
// Pipelined form: picks the next job slot, orders the internal prep stream behind `in_stream` (if any) and returns
// the job slot; job_done[slot] must be recorded on ix->s_merge by the caller once everything of the job is enqueued.
int pipelined_begin(frs_index* ix, bool has_in, cudaStream_t in_stream, int* slot, SearchLaunch* L);
int host_slot_acquire(frs_index* ix, int* slot);
void host_slot_release(frs_index* ix, int slot);
}  // namespace frs

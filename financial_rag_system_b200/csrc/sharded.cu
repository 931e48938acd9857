// sharded.cu — ONE process driving several GPUs: the chunk store sharded by rows over the GPUs of a box behind a
// single handle, which is how the reference's server would use it: get_qdrant() returns ONE client object inside
// ONE FastAPI process (main.py:92-95, main2.py:104-108) and retrieve_from_qdrant calls it from a thread pool
// (main.py:215-239, main2.py:160-163).  The multi-process form (one rank per GPU, csrc/exchange.cu) is what
// `torchrun` deployments and bench.py use; both share every kernel.
//
// Placement is block-cyclic: global row g lives in block g / B (B = kShardBlock rows), block b on shard b % n at
// local block b / n.  Appends therefore fill all shards evenly from the first row on (a contiguous split would
// leave all but one GPU idle until the store is nearly full), and local row order equals global id order, so the
// (score desc, id asc) tie rule of a shard is the global one.
//
// A search fans out from one host thread: per shard one H2D copy of the packed batch, then prep -> scan -> merge
// on that index's internal streams (index.cu, pipelined form); each shard's merge kernel writes its exact top-k
// STRAIGHT INTO THE COLLECTING GPU's gather buffer over NVLink peer memory (PushTarget without flags).  Inside
// one process CUDA events order GPUs, so the collecting GPU's stream simply waits for the shards' merge events —
// no flag spinning, no collective — runs the cross-shard merge and copies the [nq, k] result out.  Several host
// threads (or submit / collect) keep up to kHostSlots batches in flight.
#include <cuda_runtime.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <vector>

#include "index.cuh"

using namespace frs;

#define SH_TRY(expr)                                                                                              \
  do {                                                                                                            \
    cudaError_t _e = (expr);                                                                                      \
    if (_e != cudaSuccess)                                                                                        \
      return abi_set_err(FRS_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace {
constexpr int64_t kShardBlock = 4096;  // rows per placement block (a multiple of the 128-row scan tile)
constexpr size_t kPlaneWords = (size_t)kNQ * kMaxK;
constexpr size_t kBlockWords = 2 * kPlaneWords;

struct ShardSlot {                     // one batch in flight
  uint8_t* h_in = nullptr;             // pinned, shared by all shards' H2D copies
  uint8_t* h_out = nullptr;
  std::vector<uint8_t*> d_in;          // per shard
  std::vector<uint64_t*> d_local;      // per shard: [2][32][16] words, the shard's own exact top-k
  std::vector<uint64_t**> d_target;    // per shard: device array {gather} (the one push target)
  std::vector<cudaEvent_t> shard_done; // per shard, recorded on that shard's merge stream
  uint64_t* gather = nullptr;          // collecting GPU: [n][kBlockWords]
  uint8_t* d_out = nullptr;            // collecting GPU
  cudaEvent_t done = nullptr;
  bool busy = false;
  int nq = 0, k = 0;
};
}  // namespace

struct frs_sharded {
  int n = 0;
  int dtype = FRS_DTYPE_BF16;
  int64_t capacity = 0;  // global rows
  int64_t size = 0;
  std::vector<int> devices;
  std::vector<frs_index*> shards;
  std::vector<unsigned int*> counters;  // per shard: merge CTAs done (self-resetting)
  cudaStream_t s_final = nullptr;       // collecting GPU (shard 0's device)
  ShardSlot slots[kHostSlots];
  std::mutex slot_mu;
  std::condition_variable slot_cv;
  std::mutex mu;  // enqueue order / size
};

namespace {

void free_sharded(frs_sharded* sh) {
  if (!sh) return;
  for (ShardSlot& sl : sh->slots) {
    for (size_t s = 0; s < sl.d_in.size(); ++s) {
      cudaSetDevice(sh->devices[s]);
      cudaFree(sl.d_in[s]);
      if (s < sl.d_local.size()) cudaFree(sl.d_local[s]);
      if (s < sl.d_target.size()) cudaFree(sl.d_target[s]);
      if (s < sl.shard_done.size() && sl.shard_done[s]) cudaEventDestroy(sl.shard_done[s]);
    }
    if (!sh->devices.empty()) cudaSetDevice(sh->devices[0]);
    cudaFree(sl.gather);
    cudaFree(sl.d_out);
    cudaFreeHost(sl.h_in);
    cudaFreeHost(sl.h_out);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  for (size_t s = 0; s < sh->counters.size(); ++s) {
    cudaSetDevice(sh->devices[s]);
    cudaFree(sh->counters[s]);
  }
  if (sh->s_final) {
    cudaSetDevice(sh->devices[0]);
    cudaStreamDestroy(sh->s_final);
  }
  for (frs_index* ix : sh->shards) frs_index_destroy(ix);
  delete sh;
}

// global rows [row0, row0 + n) as (shard, local row0, offset into the caller's array, rows) segments
template <typename F>
int for_segments(const frs_sharded* sh, int64_t row0, int64_t n, F&& f) {
  for (int64_t o = 0; o < n;) {
    const int64_t g = row0 + o;
    const int64_t blk = g / kShardBlock, in = g % kShardBlock;
    const int64_t m = std::min<int64_t>(n - o, kShardBlock - in);
    const int s = (int)(blk % sh->n);
    const int64_t local = (blk / sh->n) * kShardBlock + in;
    int rc = f(s, local, o, m);
    if (rc) return rc;
    o += m;
  }
  return FRS_OK;
}

}  // namespace

extern "C" int frs_sharded_create(int n_devices, const int* devices, int dim, int64_t capacity_total, int dtype,
                                  frs_sharded** out) {
  if (!out) return abi_set_err(FRS_E_INVALID, "out is null");
  *out = nullptr;
  int visible = 0;
  SH_TRY(cudaGetDeviceCount(&visible));
  if (n_devices < 1 || n_devices > 64 || (!devices && n_devices > visible))
    return abi_set_err(FRS_E_INVALID, "n_devices must be in [1, %d] (got %d)", visible < 64 ? visible : 64, n_devices);
  if (capacity_total <= 0) return abi_set_err(FRS_E_INVALID, "capacity out of range");
  frs_sharded* sh = new (std::nothrow) frs_sharded();
  if (!sh) return abi_set_err(FRS_E_INVALID, "out of host memory");
  sh->n = n_devices;
  sh->dtype = dtype;
  sh->capacity = capacity_total;
  for (int s = 0; s < n_devices; ++s) {
    // a device may be listed more than once (several shards on one GPU): same code path, used by the 1-GPU tests
    const int d = devices ? devices[s] : s;
    if (d < 0 || d >= visible) {
      free_sharded(sh);
      return abi_set_err(FRS_E_INVALID, "device %d is not visible (%d devices)", d, visible);
    }
    sh->devices.push_back(d);
  }
  const int64_t blocks = (capacity_total + kShardBlock - 1) / kShardBlock;
  const int64_t cap_shard = (blocks + n_devices - 1) / n_devices * kShardBlock;
  int rc = FRS_OK;
  for (int s = 0; s < n_devices && rc == FRS_OK; ++s) {
    frs_index* ix = nullptr;
    rc = frs_index_create(sh->devices[s], dim, cap_shard, dtype, &ix);
    if (rc) break;
    ix->id_block = (uint32_t)kShardBlock;
    ix->id_shards = (uint32_t)n_devices;
    ix->id_shard = (uint32_t)s;
    sh->shards.push_back(ix);
  }
  auto cuda_fail = [&](cudaError_t e, const char* what) {
    rc = abi_set_err(FRS_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  };
  // peer access: every shard's GPU writes into the collecting GPU's gather buffers
  for (int s = 1; s < n_devices && rc == FRS_OK; ++s) {
    if (sh->devices[s] == sh->devices[0]) continue;
    int can = 0;
    cudaError_t e = cudaDeviceCanAccessPeer(&can, sh->devices[s], sh->devices[0]);
    if (e != cudaSuccess) cuda_fail(e, "cudaDeviceCanAccessPeer");
    else if (!can) rc = abi_set_err(FRS_E_CUDA, "device %d cannot access device %d's memory (no peer access)", sh->devices[s], sh->devices[0]);
    else {
      e = cudaSetDevice(sh->devices[s]);
      if (e == cudaSuccess) e = cudaDeviceEnablePeerAccess(sh->devices[0], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        e = cudaSuccess;
      }
      if (e != cudaSuccess) cuda_fail(e, "cudaDeviceEnablePeerAccess");
    }
  }
  for (int s = 0; s < n_devices && rc == FRS_OK; ++s) {
    unsigned int* c = nullptr;
    cudaError_t e = cudaSetDevice(sh->devices[s]);
    if (e == cudaSuccess) e = cudaMalloc(&c, 4);
    if (e == cudaSuccess) e = cudaMemset(c, 0, 4);
    if (e != cudaSuccess) cuda_fail(e, "counter allocation");
    sh->counters.push_back(c);
  }
  for (ShardSlot& sl : sh->slots) {
    if (rc) break;
    cudaError_t e = cudaSetDevice(sh->devices[0]);
    if (e == cudaSuccess) e = cudaMalloc(&sl.gather, (size_t)n_devices * kBlockWords * 8);
    if (e == cudaSuccess) e = cudaMemset(sl.gather, 0, (size_t)n_devices * kBlockWords * 8);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_out, kHostOutBytes);
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_in, kHostInBytes);
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_out, kHostOutBytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
    for (int s = 0; s < n_devices && e == cudaSuccess; ++s) {
      uint8_t* din = nullptr;
      uint64_t* loc = nullptr;
      uint64_t** tgt = nullptr;
      cudaEvent_t ev = nullptr;
      e = cudaSetDevice(sh->devices[s]);
      if (e == cudaSuccess) e = cudaMalloc(&din, kHostInBytes);
      if (e == cudaSuccess) e = cudaMalloc(&loc, kBlockWords * 8);
      if (e == cudaSuccess) e = cudaMalloc(&tgt, sizeof(void*));
      if (e == cudaSuccess) e = cudaMemcpy(tgt, &sl.gather, sizeof(void*), cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      sl.d_in.push_back(din);
      sl.d_local.push_back(loc);
      sl.d_target.push_back(tgt);
      sl.shard_done.push_back(ev);
    }
    if (e != cudaSuccess) cuda_fail(e, "slot allocation");
  }
  if (rc == FRS_OK) {
    cudaError_t e = cudaSetDevice(sh->devices[0]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sh->s_final, cudaStreamNonBlocking);
    for (int s = 0; s < n_devices && e == cudaSuccess; ++s) {
      e = cudaSetDevice(sh->devices[s]);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) cuda_fail(e, "stream set-up");
  }
  if (rc) {
    free_sharded(sh);
    return rc;
  }
  *out = sh;
  return FRS_OK;
}

extern "C" int frs_sharded_destroy(frs_sharded* sh) {
  if (!sh) return FRS_OK;
  for (int d : sh->devices) {
    cudaSetDevice(d);
    cudaDeviceSynchronize();
  }
  free_sharded(sh);
  return FRS_OK;
}

extern "C" int frs_sharded_n_shards(const frs_sharded* sh) { return sh ? sh->n : FRS_E_INVALID; }
extern "C" frs_index* frs_sharded_shard(frs_sharded* sh, int s) {
  return (sh && s >= 0 && s < sh->n) ? sh->shards[s] : nullptr;
}
extern "C" int64_t frs_sharded_size(const frs_sharded* sh) { return sh ? sh->size : 0; }
extern "C" int64_t frs_sharded_capacity(const frs_sharded* sh) { return sh ? sh->capacity : 0; }
extern "C" int64_t frs_sharded_block_rows(const frs_sharded* sh) { return sh ? kShardBlock : 0; }

// Rows filled in place on the shards (frs_sharded_shard + frs_index_add in block-cyclic order): publish the count.
extern "C" int frs_sharded_set_size(frs_sharded* sh, int64_t n) {
  if (!sh || n < 0 || n > sh->capacity) return abi_set_err(FRS_E_INVALID, "size out of range");
  std::lock_guard<std::mutex> lk(sh->mu);
  int64_t want = 0;
  for (int s = 0; s < sh->n; ++s) want += frs_index_size(sh->shards[s]);
  if (want != n) return abi_set_err(FRS_E_STATE, "the shards hold %lld rows, not %lld", (long long)want, (long long)n);
  sh->size = n;
  return FRS_OK;
}

extern "C" int frs_sharded_add_host(frs_sharded* sh, const float* host_vecs, const uint32_t* host_codes, int64_t n) {
  if (!sh || n < 0 || (n > 0 && !host_vecs)) return abi_set_err(FRS_E_INVALID, "bad argument");
  std::lock_guard<std::mutex> lk(sh->mu);
  if (sh->size + n > sh->capacity)
    return abi_set_err(FRS_E_CAPACITY, "index full: size %lld + %lld > capacity %lld", (long long)sh->size, (long long)n,
                       (long long)sh->capacity);
  int64_t done = 0;
  int rc = for_segments(sh, sh->size, n, [&](int s, int64_t local, int64_t o, int64_t m) {
    if (frs_index_size(sh->shards[s]) != local)
      return abi_set_err(FRS_E_STATE, "shard %d holds %lld rows where %lld were expected (filled out of order?)", s,
                         (long long)frs_index_size(sh->shards[s]), (long long)local);
    int r = frs_index_add_host(sh->shards[s], host_vecs + o * kDim, host_codes ? host_codes + o : nullptr, m);
    if (r == FRS_OK) done = o + m;
    return r;
  });
  sh->size += done;
  return rc;
}

extern "C" int frs_sharded_import_raw(frs_sharded* sh, const void* host_rows, const uint32_t* host_codes, int64_t n) {
  if (!sh || n < 0 || (n > 0 && (!host_rows || !host_codes))) return abi_set_err(FRS_E_INVALID, "bad argument");
  std::lock_guard<std::mutex> lk(sh->mu);
  if (sh->size + n > sh->capacity) return abi_set_err(FRS_E_CAPACITY, "index full");
  const size_t rb = sh->shards[0]->row_bytes();
  int64_t done = 0;
  int rc = for_segments(sh, sh->size, n, [&](int s, int64_t local, int64_t o, int64_t m) {
    if (frs_index_size(sh->shards[s]) != local) return abi_set_err(FRS_E_STATE, "shard %d filled out of order", s);
    int r = frs_index_import_raw(sh->shards[s], (const char*)host_rows + (size_t)o * rb, host_codes + o, m);
    if (r == FRS_OK) done = o + m;
    return r;
  });
  sh->size += done;
  return rc;
}

extern "C" int frs_sharded_export_raw(frs_sharded* sh, int64_t row0, int64_t n, void* host_rows, uint32_t* host_codes) {
  if (!sh || n < 0 || row0 < 0 || row0 + n > sh->size || (n > 0 && (!host_rows || !host_codes)))
    return abi_set_err(FRS_E_INVALID, "bad argument");
  const size_t rb = sh->shards[0]->row_bytes();
  return for_segments(sh, row0, n, [&](int s, int64_t local, int64_t o, int64_t m) {
    return frs_index_export_raw(sh->shards[s], local, m, (char*)host_rows + (size_t)o * rb, host_codes + o);
  });
}

extern "C" int frs_sharded_read_rows_host(frs_sharded* sh, int64_t row0, int64_t n, float* host_out) {
  if (!sh || n < 0 || row0 < 0 || row0 + n > sh->size || (n > 0 && !host_out)) return abi_set_err(FRS_E_INVALID, "bad argument");
  return for_segments(sh, row0, n, [&](int s, int64_t local, int64_t o, int64_t m) {
    return frs_index_read_rows_host(sh->shards[s], local, m, host_out + o * kDim);
  });
}

// in-place overwrite (idempotent upsert on an existing id) / payload-code update (tombstones)
extern "C" int frs_sharded_set_rows_host(frs_sharded* sh, int64_t row0, const float* host_vecs,
                                         const uint32_t* host_codes, int64_t n) {
  if (!sh || n < 0 || row0 < 0 || row0 + n > sh->size || (n > 0 && !host_vecs && !host_codes))
    return abi_set_err(FRS_E_INVALID, "bad argument");
  return for_segments(sh, row0, n, [&](int s, int64_t local, int64_t o, int64_t m) {
    frs_index* ix = sh->shards[s];
    SH_TRY(cudaSetDevice(ix->device));
    float* d_v = nullptr;
    uint32_t* d_c = nullptr;
    int rc = FRS_OK;
    cudaError_t e = cudaSuccess;
    if (host_vecs) e = cudaMalloc(&d_v, (size_t)m * kDim * 4);
    if (e == cudaSuccess && host_codes) e = cudaMalloc(&d_c, (size_t)m * 4);
    if (e == cudaSuccess && host_vecs)
      e = cudaMemcpyAsync(d_v, host_vecs + o * kDim, (size_t)m * kDim * 4, cudaMemcpyHostToDevice, ix->stream);
    if (e == cudaSuccess && host_codes)
      e = cudaMemcpyAsync(d_c, host_codes + o, (size_t)m * 4, cudaMemcpyHostToDevice, ix->stream);
    if (e != cudaSuccess) rc = abi_set_err(FRS_E_CUDA, "staging: %s", cudaGetErrorString(e));
    if (rc == FRS_OK)
      rc = host_vecs ? frs_index_set_rows(ix, local, d_v, d_c, m, ix->stream) : frs_index_set_codes(ix, local, d_c, m, ix->stream);
    if (rc == FRS_OK && (e = cudaStreamSynchronize(ix->stream)) != cudaSuccess)
      rc = abi_set_err(FRS_E_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    cudaFree(d_v);
    cudaFree(d_c);
    return rc;
  });
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
static int slot_acquire(frs_sharded* sh, int* slot) {
  std::unique_lock<std::mutex> lk(sh->slot_mu);
  int found = -1;
  const bool ok = sh->slot_cv.wait_for(lk, std::chrono::seconds(30), [&] {
    for (int i = 0; i < kHostSlots; ++i)
      if (!sh->slots[i].busy) {
        found = i;
        return true;
      }
    return false;
  });
  if (!ok) return abi_set_err(FRS_E_STATE, "all %d batch slots stayed busy for 30 s (submits without collects?)", kHostSlots);
  sh->slots[found].busy = true;
  *slot = found;
  return FRS_OK;
}
static void slot_release(frs_sharded* sh, int slot) {
  {
    std::lock_guard<std::mutex> lk(sh->slot_mu);
    sh->slots[slot].busy = false;
  }
  sh->slot_cv.notify_one();
}

extern "C" int frs_sharded_search_host_submit(frs_sharded* sh, const float* host_queries, const uint32_t* host_q_code,
                                              const uint32_t* host_q_mask, int nq, int k, int* ticket) {
  if (!sh) return abi_set_err(FRS_E_INVALID, "handle is null");
  if (nq < 1 || nq > kNQ) return abi_set_err(FRS_E_INVALID, "nq must be in [1,%d] (got %d)", kNQ, nq);
  if (k < 1 || k > kMaxK) return abi_set_err(FRS_E_INVALID, "k must be in [1,%d] (got %d)", kMaxK, k);
  if (!host_queries || !host_q_code || !host_q_mask || !ticket) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  int si = 0;
  int rc = slot_acquire(sh, &si);
  if (rc) return rc;
  ShardSlot& sl = sh->slots[si];
  const size_t qb = (size_t)nq * kDim * 4;
  memcpy(sl.h_in, host_queries, qb);
  memcpy(sl.h_in + qb, host_q_code, (size_t)nq * 4);
  memcpy(sl.h_in + qb + (size_t)nq * 4, host_q_mask, (size_t)nq * 4);
  sl.nq = nq;
  sl.k = k;
  auto fail = [&](int code) {
    slot_release(sh, si);
    return code;
  };
  std::lock_guard<std::mutex> lk(sh->mu);
  for (int s = 0; s < sh->n; ++s) {
    frs_index* ix = sh->shards[s];
    cudaError_t e = cudaSetDevice(ix->device);
    if (e != cudaSuccess) return fail(abi_set_err(FRS_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e)));
    std::lock_guard<std::mutex> lki(ix->mu);
    SearchLaunch L{};
    int job = 0;
    rc = pipelined_begin(ix, false, nullptr, &job, &L);
    if (rc) return fail(rc);
    e = cudaMemcpyAsync(sl.d_in[s], sl.h_in, qb + (size_t)nq * 8, cudaMemcpyHostToDevice, ix->s_prep);
    if (e != cudaSuccess) return fail(abi_set_err(FRS_E_CUDA, "cudaMemcpyAsync (queries): %s", cudaGetErrorString(e)));
    PushTarget t{};
    t.peer_gather = sl.d_target[s];
    t.peer_flags = nullptr;  // ordered by CUDA events below
    t.counter = sh->counters[s];
    t.n_targets = 1;
    t.world = sh->n;
    t.rank = s;
    t.seq = 0;
    t.block_words = kBlockWords;
    t.plane_words = kPlaneWords;
    SearchArgs a;
    a.q = reinterpret_cast<const float*>(sl.d_in[s]);
    a.code = reinterpret_cast<const uint32_t*>(sl.d_in[s] + qb);
    a.mask = a.code + nq;
    a.nq = nq;
    a.k = k;
    a.out_s64 = reinterpret_cast<double*>(sl.d_local[s]);
    a.out_ids = reinterpret_cast<int64_t*>(sl.d_local[s]) + kPlaneWords;
    a.push = &t;
    rc = search_enqueue(ix, a, L);
    if (rc) return fail(rc);
    e = cudaEventRecord(sl.shard_done[s], ix->s_merge);
    if (e == cudaSuccess) e = cudaEventRecord(ix->job_done[job], ix->s_merge);
    if (e != cudaSuccess) return fail(abi_set_err(FRS_E_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e)));
  }
  // collecting GPU: wait for every shard's merge (+ peer writes), cross-shard merge, one copy out
  cudaError_t e = cudaSetDevice(sh->devices[0]);
  for (int s = 0; s < sh->n && e == cudaSuccess; ++s) e = cudaStreamWaitEvent(sh->s_final, sl.shard_done[s], 0);
  if (e != cudaSuccess) return fail(abi_set_err(FRS_E_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e)));
  int64_t* d_ids = reinterpret_cast<int64_t*>(sl.d_out);
  float* d_scores = reinterpret_cast<float*>(sl.d_out + (size_t)nq * k * 8);
  e = launch_merge_shards(reinterpret_cast<const double*>(sl.gather), reinterpret_cast<const int64_t*>(sl.gather) + kPlaneWords,
                          sh->n, nq, k, kBlockWords, d_scores, d_ids, sh->s_final);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sl.h_out, sl.d_out, (size_t)nq * k * 12, cudaMemcpyDeviceToHost, sh->s_final);
  if (e == cudaSuccess) e = cudaEventRecord(sl.done, sh->s_final);
  if (e != cudaSuccess) return fail(abi_set_err(FRS_E_CUDA, "final merge: %s", cudaGetErrorString(e)));
  *ticket = si;
  return FRS_OK;
}

extern "C" int frs_sharded_search_host_collect(frs_sharded* sh, int ticket, float* host_out_scores, int64_t* host_out_ids) {
  if (!sh || ticket < 0 || ticket >= kHostSlots || !host_out_scores || !host_out_ids)
    return abi_set_err(FRS_E_INVALID, "bad argument");
  ShardSlot& sl = sh->slots[ticket];
  if (!sl.busy) return abi_set_err(FRS_E_STATE, "ticket %d is not in flight", ticket);
  int rc = FRS_OK;
  cudaError_t e = cudaEventSynchronize(sl.done);
  if (e != cudaSuccess) rc = abi_set_err(FRS_E_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
  if (rc == FRS_OK) {
    memcpy(host_out_ids, sl.h_out, (size_t)sl.nq * sl.k * 8);
    memcpy(host_out_scores, sl.h_out + (size_t)sl.nq * sl.k * 8, (size_t)sl.nq * sl.k * 4);
  }
  slot_release(sh, ticket);
  return rc;
}

extern "C" int frs_sharded_search_host(frs_sharded* sh, const float* host_queries, const uint32_t* host_q_code,
                                       const uint32_t* host_q_mask, int nq, int k, float* host_out_scores,
                                       int64_t* host_out_ids) {
  if (!host_out_scores || !host_out_ids) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  int ticket = -1;
  int rc = frs_sharded_search_host_submit(sh, host_queries, host_q_code, host_q_mask, nq, k, &ticket);
  if (rc) return rc;
  return frs_sharded_search_host_collect(sh, ticket, host_out_scores, host_out_ids);
}

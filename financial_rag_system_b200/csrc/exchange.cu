// exchange.cu — the cross-shard exchange of the sharded search over NVLink peer memory (no collective library).
//
// One process per GPU.  After its local pass a rank holds its shard's exact top-k as [2][nq][k] 64-bit words
// (plane 0 = fp64 score bits, plane 1 = int64 global ids; 7.7 KB at 32 x 15).  Instead of an all-gather,
//
//   push   every rank WRITES that block into slot `rank` of EVERY peer's gather buffer with plain stores through
//          peer-mapped pointers (CUDA IPC handles opened once), fences at system scope and then publishes its
//          sequence number in the peer's flag word.  frs_index_search_push (index.cu) does this in the TAIL OF
//          THE LOCAL MERGE KERNEL (scan.cu merge_kernel: one CTA per query pushes its k results, the last CTA
//          publishes the flags); frs_exchange_push is the stand-alone form for a block that already exists;
//   wait   the HEAD of the cross-shard merge kernel (scan.cu wait_merge_shards_kernel): one warp polls the `world`
//          flags (ld.acquire.sys) until all have reached the current sequence number, then the block merges.  In the
//          pipelined engine that kernel runs on the index's own exchange stream (index.cu exchange_tail), so a late
//          peer holds up nothing but this batch's final result.  The wait is
//          bounded by a wall-clock time-out (default 30 s, frs_exchange_set_timeout_ms): a rank that is merely late
//          (host stall, lazy module load, a save in progress) is waited for; a lost one poisons the batch — the
//          merge emits an empty result and the host entry points return FRS_E_TIMEOUT — instead of trapping,
//          which would destroy this rank's CUDA context and its resident shard;
//   merge  the existing cross-shard merge kernel runs over the local gather buffer.
//
// The data crosses NVSwitch exactly once per (source, destination) pair, no kernel of a collective library has to be
// resident next to the persistent scan kernel, and the latency is one store + one flag per peer.
//
// Buffer reuse: gather buffers are a ring of kExchangeSlots = 4 batches (scan.cuh): a rank may issue the push of
// batch t + 2 only after its own final merge of batch t (index.cu exchange_order_push makes the merge stream wait for
// the event `merged[(t) % 4]`; the synchronous form pushes t + 1 after merging t), which keeps a slot from being overwritten while a peer still
// reads it.  Flags are monotonic (no reset, no ABA).
//
// The reference has no counterpart (one Qdrant server, main.py:215-239); this is north-star item (3), the exchange
// step of `ShardedIndex` (sharded.py).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "index.cuh"

using frs::abi_set_err;
using frs::kExchangeSlots;

#define EX_TRY(expr)                                                                                              \
  do {                                                                                                            \
    cudaError_t _e = (expr);                                                                                      \
    if (_e != cudaSuccess)                                                                                        \
      return abi_set_err(FRS_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace {

// grid = world CTAs: CTA p copies the local block into peer p's gather slot, then publishes the sequence number
__global__ void __launch_bounds__(256)
exchange_push_kernel(const uint64_t* __restrict__ local, uint64_t* const* __restrict__ peer_gather,
                     uint32_t* const* __restrict__ peer_flags, int world, int rank, uint32_t words,
                     size_t block_words, uint32_t seq) {
  const int peer = blockIdx.x;
  uint64_t* dst = peer_gather[peer] + ((size_t)(seq % kExchangeSlots) * world + rank) * block_words;
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = local[i];
  __threadfence_system();  // the payload is visible system-wide before the flag
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[peer] + rank), "r"(seq) : "memory");
  }
}

void free_exchange(frs_exchange* ex) {
  for (void* p : ex->opened) cudaIpcCloseMemHandle(p);
  cudaFree(ex->gather);
  cudaFree(ex->flags);
  cudaFree(ex->d_peer_gather);
  cudaFree(ex->d_peer_flags);
  cudaFree(ex->local);
  cudaFree(ex->counter);
  cudaFree(ex->poison);
  for (void* ev : ex->merged)
    if (ev) cudaEventDestroy((cudaEvent_t)ev);
  if (ex->h_status) cudaFreeHost(const_cast<uint32_t*>(ex->h_status));
  delete ex;
}

}  // namespace

extern "C" int frs_exchange_create(int device, int world, int rank, int nq_max, int k_max, frs_exchange** out) {
  if (!out) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  *out = nullptr;
  if (world < 1 || world > 64 || rank < 0 || rank >= world || nq_max < 1 || nq_max > FRS_MAX_BATCH || k_max < 1 ||
      k_max > FRS_MAX_K)
    return abi_set_err(FRS_E_INVALID, "bad exchange shape (world %d rank %d nq %d k %d)", world, rank, nq_max, k_max);
  EX_TRY(cudaSetDevice(device));
  frs_exchange* ex = new frs_exchange();
  ex->device = device;
  ex->world = world;
  ex->rank = rank;
  ex->nq_max = nq_max;
  ex->k_max = k_max;
  ex->plane_words = (size_t)nq_max * k_max;
  ex->block_words = 2 * ex->plane_words;
  const size_t gbytes = (size_t)kExchangeSlots * world * ex->block_words * 8;
  uint32_t* hs = nullptr;
  // plain cudaMalloc (not a caching-allocator sub-block): the IPC handle names exactly this allocation
  cudaError_t e = cudaMalloc(&ex->gather, gbytes);
  if (e == cudaSuccess) e = cudaMalloc(&ex->flags, (size_t)world * 4);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_peer_gather, (size_t)world * sizeof(void*));
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_peer_flags, (size_t)world * sizeof(void*));
  if (e == cudaSuccess) e = cudaMalloc(&ex->local, ex->block_words * 8);
  if (e == cudaSuccess) e = cudaMalloc(&ex->counter, 4);
  if (e == cudaSuccess) e = cudaMalloc(&ex->poison, 4);
  if (e == cudaSuccess) e = cudaHostAlloc(&hs, 4, cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *hs = 0;
    ex->h_status = hs;
    e = cudaHostGetDevicePointer(&ex->d_status, hs, 0);
  }
  cudaFuncAttributes fa;  // load the exchange kernels now, not at their first launch (see preload_search_kernels)
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, (const void*)exchange_push_kernel);
  if (e == cudaSuccess) e = frs::preload_search_kernels();
  if (e == cudaSuccess) e = cudaMemset(ex->gather, 0, gbytes);
  if (e == cudaSuccess) e = cudaMemset(ex->flags, 0, (size_t)world * 4);
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->block_words * 8);
  if (e == cudaSuccess) e = cudaMemset(ex->counter, 0, 4);
  if (e == cudaSuccess) e = cudaMemset(ex->poison, 0, 4);
  static_assert(kExchangeSlots == 4, "frs_exchange::merged has one event per gather slot");
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) {
    cudaEvent_t ev = nullptr;
    e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    ex->merged[i] = ev;
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    free_exchange(ex);
    return abi_set_err(FRS_E_CUDA, "exchange set-up failed: %s", cudaGetErrorString(e));
  }
  *out = ex;
  return FRS_OK;
}

extern "C" int frs_exchange_destroy(frs_exchange* ex) {
  if (!ex) return FRS_OK;
  cudaSetDevice(ex->device);
  cudaDeviceSynchronize();
  free_exchange(ex);
  return FRS_OK;
}

extern "C" int frs_exchange_set_timeout_ms(frs_exchange* ex, int64_t ms) {
  if (!ex || ms < 1) return abi_set_err(FRS_E_INVALID, "bad argument");
  ex->timeout_ns = (unsigned long long)ms * 1000000ull;
  return FRS_OK;
}

namespace frs {
int exchange_check(frs_exchange* ex) {
  const uint32_t bad = ex->h_status ? *ex->h_status : 0u;
  if (bad)
    return abi_set_err(FRS_E_TIMEOUT,
                       "exchange: a peer rank did not publish batch %u within %.1f s; results from that batch on are "
                       "empty — destroy and re-create the exchange on every rank", bad, ex->timeout_ns * 1e-9);
  return FRS_OK;
}
}  // namespace frs

extern "C" int frs_exchange_status(frs_exchange* ex) {
  if (!ex) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  return frs::exchange_check(ex);
}

// 128 bytes: IPC handle of the gather buffer | IPC handle of the flag array
extern "C" int frs_exchange_handle(frs_exchange* ex, uint8_t* out128) {
  if (!ex || !out128) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  EX_TRY(cudaSetDevice(ex->device));
  cudaIpcMemHandle_t h;
  EX_TRY(cudaIpcGetMemHandle(&h, ex->gather));
  memcpy(out128, &h, 64);
  EX_TRY(cudaIpcGetMemHandle(&h, ex->flags));
  memcpy(out128 + 64, &h, 64);
  return FRS_OK;
}

static int install_peers(frs_exchange* ex, const std::vector<uint64_t*>& g, const std::vector<uint32_t*>& f) {
  EX_TRY(cudaMemcpy(ex->d_peer_gather, g.data(), g.size() * sizeof(void*), cudaMemcpyHostToDevice));
  EX_TRY(cudaMemcpy(ex->d_peer_flags, f.data(), f.size() * sizeof(void*), cudaMemcpyHostToDevice));
  ex->connected = true;
  return FRS_OK;
}

// handles: [world][128] as produced by frs_exchange_handle on every rank (all-gathered by the caller)
extern "C" int frs_exchange_connect(frs_exchange* ex, const uint8_t* handles) {
  if (!ex || !handles) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  EX_TRY(cudaSetDevice(ex->device));
  std::vector<uint64_t*> g(ex->world);
  std::vector<uint32_t*> f(ex->world);
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank) {
      g[r] = ex->gather;
      f[r] = ex->flags;
      continue;
    }
    cudaIpcMemHandle_t h;
    void* p = nullptr;
    memcpy(&h, handles + (size_t)r * 128, 64);
    EX_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ex->opened.push_back(p);
    g[r] = static_cast<uint64_t*>(p);
    memcpy(&h, handles + (size_t)r * 128 + 64, 64);
    EX_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ex->opened.push_back(p);
    f[r] = static_cast<uint32_t*>(p);
  }
  return install_peers(ex, g, f);
}

// in-process form (several shards of one process, tests): the peers' buffers by pointer
extern "C" int frs_exchange_connect_local(frs_exchange* ex, frs_exchange* const* peers) {
  if (!ex || !peers) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  EX_TRY(cudaSetDevice(ex->device));
  std::vector<uint64_t*> g(ex->world);
  std::vector<uint32_t*> f(ex->world);
  for (int r = 0; r < ex->world; ++r) {
    if (!peers[r] || peers[r]->world != ex->world || peers[r]->block_words != ex->block_words)
      return abi_set_err(FRS_E_INVALID, "peer %d does not match this exchange", r);
    if (peers[r]->device != ex->device) {  // several GPUs of one process: map the peer's memory
      int can = 0;
      EX_TRY(cudaDeviceCanAccessPeer(&can, ex->device, peers[r]->device));
      if (!can) return abi_set_err(FRS_E_CUDA, "device %d cannot access device %d", ex->device, peers[r]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return abi_set_err(FRS_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
      cudaGetLastError();
    }
    g[r] = peers[r]->gather;
    f[r] = peers[r]->flags;
  }
  return install_peers(ex, g, f);
}

namespace frs {

// The push target of this rank's NEXT batch (sequence number seq + 1).  The sequence number itself advances in
// exchange_commit_push, after the kernel carrying the push has been enqueued successfully: a failed launch leaves
// the ranks' sequence numbers in step.
int exchange_begin_push(frs_exchange* ex, int nq, int k, PushTarget* t) {
  (void)nq;
  (void)k;
  int rc = exchange_check(ex);
  if (rc) return rc;
  t->peer_gather = ex->d_peer_gather;
  t->peer_flags = ex->d_peer_flags;
  t->counter = ex->counter;
  t->n_targets = ex->world;
  t->world = ex->world;
  t->rank = ex->rank;
  t->seq = ex->seq + 1;
  t->block_words = ex->block_words;
  t->plane_words = ex->plane_words;
  return FRS_OK;
}
void exchange_commit_push(frs_exchange* ex) { ++ex->seq; }

// waits for every rank's push of the current sequence number, then merges [world][2][nq_max][k_max] -> [nq][k]
int exchange_wait_merge(frs_exchange* ex, int nq, int k, float* out_s, int64_t* out_i, cudaStream_t st) {
  const uint64_t* slot = ex->gather + (size_t)(ex->seq % kExchangeSlots) * ex->world * ex->block_words;
  EX_TRY(launch_wait_merge_shards(ex->flags, ex->world, ex->seq, ex->timeout_ns, ex->poison, ex->d_status,
                                  reinterpret_cast<const double*>(slot), reinterpret_cast<const int64_t*>(slot) + ex->plane_words,
                                  nq, k, ex->block_words, out_s, out_i, st));
  return FRS_OK;
}

}  // namespace frs

// dev_local_packed: this rank's block in the exchange's layout — [2][nq_max][k_max] words, entry (q, r) of a plane at
// q * k + r (dense [2][nq][k] when nq, k are the ones the exchange was created with).  Asynchronous on `stream`.
extern "C" int frs_exchange_push(frs_exchange* ex, const int64_t* dev_local_packed, void* stream) {
  if (!ex || !dev_local_packed) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (!ex->connected) return abi_set_err(FRS_E_INVALID, "exchange is not connected");
  int rc = frs::exchange_check(ex);
  if (rc) return rc;
  EX_TRY(cudaSetDevice(ex->device));
  exchange_push_kernel<<<ex->world, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint64_t*>(dev_local_packed), ex->d_peer_gather, ex->d_peer_flags, ex->world, ex->rank,
      (uint32_t)ex->block_words, ex->block_words, ex->seq + 1);
  EX_TRY(cudaGetLastError());
  ++ex->seq;
  return FRS_OK;
}

extern "C" int frs_exchange_wait_merge(frs_exchange* ex, float* dev_out_scores, int64_t* dev_out_ids, void* stream) {
  return frs_exchange_wait_merge_n(ex, ex ? ex->nq_max : 0, ex ? ex->k_max : 0, dev_out_scores, dev_out_ids, stream);
}

extern "C" int frs_exchange_wait_merge_n(frs_exchange* ex, int nq, int k, float* dev_out_scores, int64_t* dev_out_ids,
                                         void* stream) {
  if (!ex || !dev_out_scores || !dev_out_ids) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (!ex->connected || ex->seq == 0) return abi_set_err(FRS_E_INVALID, "nothing was pushed");
  if (nq < 1 || nq > ex->nq_max || k < 1 || k > ex->k_max) return abi_set_err(FRS_E_INVALID, "nq / k exceed the exchange's");
  EX_TRY(cudaSetDevice(ex->device));
  int rc = frs::exchange_wait_merge(ex, nq, k, dev_out_scores, dev_out_ids, (cudaStream_t)stream);
  if (rc) return rc;
  // (the pipelined forms order the push of sequence number s + 2 behind this merge of s: gather slots are a ring)
  EX_TRY(cudaEventRecord((cudaEvent_t)ex->merged[ex->seq % kExchangeSlots], (cudaStream_t)stream));
  return FRS_OK;
}

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
//   wait   a one-warp kernel spins (bounded) until all `world` flags have reached the current sequence number;
//   merge  the existing cross-shard merge kernel runs over the local gather buffer.
//
// The data crosses NVSwitch exactly once per (source, destination) pair, no kernel of a collective library has to be
// resident next to the persistent scan kernel, and the latency is one store + one flag per peer.
//
// Buffer reuse: gather buffers are a ring of kExchangeSlots = 4 batches (scan.cuh): a rank may issue the push of
// batch t + 2 only after its own final merge of batch t (ShardedIndex.search_async orders its streams that way; the
// synchronous form pushes t + 1 after merging t), which keeps a slot from being overwritten while a peer still
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

#include "../../include/frs_b200.h"
#include "exchange.cuh"
#include "scan.cuh"

namespace frs {
int abi_set_err(int code, const char* fmt, ...);
}
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

// one warp: lane r waits for rank r's flag (bounded: a lost peer must trap, not hang the GPU)
__global__ void exchange_wait_kernel(const uint32_t* __restrict__ flags, int world, uint32_t seq) {
  for (int r = threadIdx.x; r < world; r += blockDim.x) {
    uint32_t v, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
      if ((int32_t)(v - seq) >= 0) break;
      __nanosleep(200);
    } while (++spins < (1u << 24));  // ~ several seconds
    if ((int32_t)(v - seq) < 0) __trap();
  }
}

}  // namespace

extern "C" int frs_exchange_create(int device, int world, int rank, int nq_max, int k_max, frs_exchange** out) {
  if (!out) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (world < 1 || world > 64 || rank < 0 || rank >= world || nq_max < 1 || k_max < 1 || k_max > FRS_MAX_K)
    return abi_set_err(FRS_E_INVALID, "bad exchange shape (world %d rank %d nq %d k %d)", world, rank, nq_max, k_max);
  EX_TRY(cudaSetDevice(device));
  frs_exchange* ex = new frs_exchange();
  ex->device = device;
  ex->world = world;
  ex->rank = rank;
  ex->nq_max = nq_max;
  ex->k_max = k_max;
  ex->block_words = 2 * (size_t)nq_max * k_max;
  const size_t gbytes = (size_t)kExchangeSlots * world * ex->block_words * 8;
  // plain cudaMalloc (not a caching-allocator sub-block): the IPC handle names exactly this allocation
  if (cudaMalloc(&ex->gather, gbytes) != cudaSuccess || cudaMalloc(&ex->flags, (size_t)world * 4) != cudaSuccess ||
      cudaMalloc(&ex->d_peer_gather, (size_t)world * sizeof(void*)) != cudaSuccess ||
      cudaMalloc(&ex->d_peer_flags, (size_t)world * sizeof(void*)) != cudaSuccess ||
      cudaMalloc(&ex->local, ex->block_words * 8) != cudaSuccess || cudaMalloc(&ex->counter, 4) != cudaSuccess) {
    delete ex;
    return abi_set_err(FRS_E_CUDA, "exchange buffer allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  EX_TRY(cudaMemset(ex->gather, 0, gbytes));
  EX_TRY(cudaMemset(ex->flags, 0, (size_t)world * 4));
  EX_TRY(cudaMemset(ex->counter, 0, 4));
  EX_TRY(cudaDeviceSynchronize());
  *out = ex;
  return FRS_OK;
}

extern "C" int frs_exchange_destroy(frs_exchange* ex) {
  if (!ex) return FRS_OK;
  cudaSetDevice(ex->device);
  cudaDeviceSynchronize();
  for (void* p : ex->opened) cudaIpcCloseMemHandle(p);
  cudaFree(ex->gather);
  cudaFree(ex->flags);
  cudaFree(ex->d_peer_gather);
  cudaFree(ex->d_peer_flags);
  cudaFree(ex->local);
  cudaFree(ex->counter);
  delete ex;
  return FRS_OK;
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
    g[r] = peers[r]->gather;
    f[r] = peers[r]->flags;
  }
  return install_peers(ex, g, f);
}

// dev_local_packed: this rank's [2][nq][k] words (nq, k as created).  Asynchronous on `stream`.
extern "C" int frs_exchange_push(frs_exchange* ex, const int64_t* dev_local_packed, void* stream) {
  if (!ex || !dev_local_packed) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (!ex->connected) return abi_set_err(FRS_E_INVALID, "exchange is not connected");
  EX_TRY(cudaSetDevice(ex->device));
  ++ex->seq;
  exchange_push_kernel<<<ex->world, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint64_t*>(dev_local_packed), ex->d_peer_gather, ex->d_peer_flags, ex->world, ex->rank,
      (uint32_t)ex->block_words, ex->block_words, ex->seq);
  EX_TRY(cudaGetLastError());
  return FRS_OK;
}

// waits for every rank's push of the current sequence number, then merges [world][2][nq][k] -> [nq][k]
extern "C" int frs_exchange_wait_merge(frs_exchange* ex, float* dev_out_scores, int64_t* dev_out_ids, void* stream) {
  if (!ex || !dev_out_scores || !dev_out_ids) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (!ex->connected || ex->seq == 0) return abi_set_err(FRS_E_INVALID, "nothing was pushed");
  EX_TRY(cudaSetDevice(ex->device));
  cudaStream_t st = (cudaStream_t)stream;
  exchange_wait_kernel<<<1, 64, 0, st>>>(ex->flags, ex->world, ex->seq);
  EX_TRY(cudaGetLastError());
  const uint64_t* slot = ex->gather + (size_t)(ex->seq % kExchangeSlots) * ex->world * ex->block_words;
  const size_t plane = (size_t)ex->nq_max * ex->k_max;
  EX_TRY(frs::launch_merge_shards(reinterpret_cast<const double*>(slot), reinterpret_cast<const int64_t*>(slot) + plane,
                                  ex->world, ex->nq_max, ex->k_max, 2 * plane, dev_out_scores, dev_out_ids, st));
  return FRS_OK;
}

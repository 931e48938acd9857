// exchange.cuh — the exchange object shared by exchange.cu (ABI) and index.cu (the push fused into the local merge)
#pragma once
#include <stdint.h>

#include <vector>

struct frs_exchange {
  int device = 0, world = 1, rank = 0, nq_max = 0, k_max = 0;
  size_t block_words = 0;             // 2 * nq_max * k_max 64-bit words per (slot, source rank)
  size_t plane_words = 0;             // nq_max * k_max: a block is [score plane | id plane], entry (q, r) at q * k + r
  uint64_t* gather = nullptr;         // [kExchangeSlots][world][block_words]   (local, written by the peers)
  uint32_t* flags = nullptr;          // [world]: last sequence number pushed by each rank (local, written by peers)
  uint64_t** d_peer_gather = nullptr; // device array [world]: every rank's gather buffer as seen from this GPU
  uint32_t** d_peer_flags = nullptr;  // device array [world]
  uint64_t* local = nullptr;          // [block_words]: this rank's block (output of the local merge)
  unsigned int* counter = nullptr;    // merge CTAs that have pushed (fused form)
  // Time-out reporting (a late or lost peer must not take this rank's CUDA context down with it): the wait kernel
  // sets `poison` (device; the merge that follows then emits an empty result) and writes the sequence number that
  // timed out into `h_status` (pinned, host-mapped), which the host entry points turn into FRS_E_TIMEOUT.
  uint32_t* poison = nullptr;
  volatile uint32_t* h_status = nullptr;
  uint32_t* d_status = nullptr;
  unsigned long long timeout_ns = 30ull * 1000 * 1000 * 1000;
  // merged[s % kSlots]: recorded (on the stream that ran it) after this rank's cross-shard merge of sequence number s.
  // The push of s + 2 waits for it: gather slots are a ring of four (see scan.cuh: D >= 2 d).
  void* merged[4] = {nullptr, nullptr, nullptr, nullptr};  // cudaEvent_t
  std::vector<void*> opened;          // IPC mappings to close
  uint32_t seq = 0;
  bool connected = false;
};

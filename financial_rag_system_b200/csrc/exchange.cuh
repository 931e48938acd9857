// exchange.cuh — the exchange object shared by exchange.cu (ABI) and index.cu (the push fused into the local merge)
#pragma once
#include <stdint.h>

#include <vector>

struct frs_exchange {
  int device = 0, world = 1, rank = 0, nq_max = 0, k_max = 0;
  size_t block_words = 0;             // 2 * nq_max * k_max 64-bit words per (slot, source rank)
  uint64_t* gather = nullptr;         // [kExchangeSlots][world][block_words]   (local, written by the peers)
  uint32_t* flags = nullptr;          // [world]: last sequence number pushed by each rank (local, written by peers)
  uint64_t** d_peer_gather = nullptr; // device array [world]: every rank's gather buffer as seen from this GPU
  uint32_t** d_peer_flags = nullptr;  // device array [world]
  uint64_t* local = nullptr;          // [block_words]: this rank's block (output of the local merge)
  unsigned int* counter = nullptr;    // merge CTAs that have pushed (fused form)
  std::vector<void*> opened;          // IPC mappings to close
  uint32_t seq = 0;
  bool connected = false;
};

// index.cu — C ABI of the chunk store (include/frs_b200.h): the sharded, GPU-resident replacement
// of the Qdrant collection used by the reference (create_collection ingest.py:86-96, upsert
// ingest.py:171-175, query_points main.py:232-237 / main2.py:163).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <mutex>
#include <new>

#include "index.cuh"

using namespace frs;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU_TRY(expr)                                                                          \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return set_err(FRS_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                     __FILE__, __LINE__);                                                     \
  } while (0)

namespace frs {
int abi_set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace frs

extern "C" const char* frs_last_error(void) { return g_err; }
extern "C" int frs_version(void) { return FRS_VERSION; }
extern "C" int frs_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return set_err(FRS_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

// ---------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched at run time: no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p)
      return set_err(FRS_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return FRS_OK;
}

namespace frs {
// generic 2-D bf16 row-major [rows, cols] map with a (box_cols x box_rows) box; the swizzle span equals the
// box's inner extent: 64 columns -> SWIZZLE_128B, 32 columns -> SWIZZLE_64B
int abi_make_tmap_bf16(CUtensorMap* m, void* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                       uint32_t box_rows) {
  EncodeTiledFn enc = nullptr;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (box_cols != 64 && box_cols != 32) return set_err(FRS_E_INVALID, "tensor map box must be 32 or 64 columns wide");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(FRS_E_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return FRS_OK;
}
}  // namespace frs

// [rows, 384] row-major matrix, box = one 128-byte-wide K-slab of `box_rows` rows, SWIZZLE_128B
static int make_tmap(CUtensorMap* m, void* base, bool f32, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = nullptr;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  const uint32_t esz = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kDim * esz};
  cuuint32_t box[2] = {128u / esz, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(FRS_E_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return FRS_OK;
}

// ---------------------------------------------------------------------------------------------
// index object (struct frs_index: index.cuh)
// ---------------------------------------------------------------------------------------------
static void free_index(frs_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  cudaFree(ix->rows);
  cudaFree(ix->codes);
  for (SearchWs& w : ix->ws) {
    cudaFree(w.qop);
    cudaFree(w.qrec);
    cudaFree(w.qcode);
    cudaFree(w.qmask);
    cudaFree(w.part_keys);
    cudaFree(w.part_cnt);
    cudaFree(w.stats);
    cudaFree(w.gmax);
    cudaFree(w.gsample);
    cudaFree(w.spill);
    if (w.free) cudaEventDestroy(w.free);
  }
  cudaFree(ix->timeline);
  if (ix->br_first) cudaEventDestroy(ix->br_first);
  if (ix->br_last) cudaEventDestroy(ix->br_last);
  if (ix->prof_ev) {
    for (int i = 0; i < frs_index::kProfRing * kProfEvents; ++i) cudaEventDestroy(ix->prof_ev[i]);
    delete[] ix->prof_ev;
  }
  for (HostSlot& h : ix->hs) {
    cudaFree(h.d_in);
    cudaFree(h.d_out);
    cudaFreeHost(h.h_in);
    cudaFreeHost(h.h_out);
    if (h.done) cudaEventDestroy(h.done);
  }
  for (int i = 0; i < kJobRing; ++i) {
    if (ix->job_in[i]) cudaEventDestroy(ix->job_in[i]);
    if (ix->job_prep[i]) cudaEventDestroy(ix->job_prep[i]);
    if (ix->job_scan[i]) cudaEventDestroy(ix->job_scan[i]);
    if (ix->job_done[i]) cudaEventDestroy(ix->job_done[i]);
    if (ix->job_merge[i]) cudaEventDestroy(ix->job_merge[i]);
  }
  if (ix->rows_ready) cudaEventDestroy(ix->rows_ready);
  if (ix->xchg_last) cudaEventDestroy(ix->xchg_last);
  for (cudaStream_t s : {ix->s_prep, ix->s_scan[0], ix->s_scan[1], ix->s_merge, ix->s_xchg, ix->stream})
    if (s) cudaStreamDestroy(s);
  delete ix;
}

extern "C" int frs_index_create(int device, int dim, int64_t capacity, int dtype, frs_index** out) {
  if (!out) return set_err(FRS_E_INVALID, "out is null");
  *out = nullptr;
  if (dim != kDim) return set_err(FRS_E_INVALID, "dim must be %d (got %d)", kDim, dim);
  if (dtype != FRS_DTYPE_F32 && dtype != FRS_DTYPE_BF16)
    return set_err(FRS_E_INVALID, "dtype must be FRS_DTYPE_F32 or FRS_DTYPE_BF16");
  if (capacity <= 0 || capacity >= (int64_t)0xFFFFFF00ll)
    return set_err(FRS_E_INVALID, "capacity out of range");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(FRS_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                   prop.major, prop.minor);
  frs_index* ix = new (std::nothrow) frs_index();
  if (!ix) return set_err(FRS_E_INVALID, "out of host memory");
  for (int i = 0; i < kJobRing; ++i) ix->job_in[i] = ix->job_prep[i] = ix->job_scan[i] = ix->job_done[i] = ix->job_merge[i] = nullptr;
  ix->device = device;
  ix->dtype = dtype;
  ix->capacity = capacity;
  ix->sm_count = prop.multiProcessorCount;
  ix->max_parts = ix->sm_count < kGmaxPad ? ix->sm_count : kGmaxPad;
  const bool f32 = ix->f32();
  // rows padded to a whole tile so a TMA box never straddles the allocation end
  const int64_t cap_pad = (capacity + kTileM - 1) / kTileM * kTileM;
#define IX_TRY(expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      set_err(FRS_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));                 \
      free_index(ix);                                                                      \
      return FRS_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)
  IX_TRY(cudaMalloc(&ix->rows, (size_t)cap_pad * ix->row_bytes()));
  // never-written rows must not hold NaN bit patterns: a tile is read whole and 0 * NaN would poison its neighbours' sums
  IX_TRY(cudaMemset(ix->rows, 0, (size_t)cap_pad * ix->row_bytes()));
  IX_TRY(cudaMalloc(&ix->codes, (size_t)cap_pad * 4));
  IX_TRY(cudaMemset(ix->codes, 0xFF, (size_t)cap_pad * 4));
  for (SearchWs& w : ix->ws) {
    IX_TRY(cudaMalloc(&w.qop, (size_t)kNQ * ix->row_bytes()));
    IX_TRY(cudaMalloc(&w.qrec, (size_t)kNQ * kDim * 4));
    IX_TRY(cudaMemset(w.qrec, 0, (size_t)kNQ * kDim * 4));
    IX_TRY(cudaMalloc(&w.qcode, kNQ * 4));
    IX_TRY(cudaMalloc(&w.qmask, kNQ * 4));
    IX_TRY(cudaMalloc(&w.part_keys, (size_t)ix->max_parts * kNQ * kListCap * 8));
    IX_TRY(cudaMalloc(&w.part_cnt, (size_t)ix->max_parts * kNQ * 4));
    IX_TRY(cudaMalloc(&w.stats, kStatSlots * 8));
    IX_TRY(cudaMemset(w.stats, 0, kStatSlots * 8));
    IX_TRY(cudaMalloc(&w.gmax, (size_t)kNQ * kGmaxPad * 4));
    IX_TRY(cudaMalloc(&w.gsample, (size_t)kSampleFloats * 4));
    IX_TRY(cudaMemset(w.gsample, 0, (size_t)kSampleFloats * 4));  // (the block counter starts at 0 and resets itself)
    IX_TRY(cudaMalloc(&w.spill, merge_spill_bytes()));
    IX_TRY(cudaEventCreateWithFlags(&w.free, cudaEventDisableTiming));
  }
  IX_TRY(cudaMalloc(&ix->timeline, (size_t)kGmaxPad * 16 * 8));
  IX_TRY(cudaMemset(ix->timeline, 0, (size_t)kGmaxPad * 16 * 8));
  for (HostSlot& h : ix->hs) {
    IX_TRY(cudaMalloc(&h.d_in, kHostInBytes));
    IX_TRY(cudaMalloc(&h.d_out, kHostOutBytes));
    IX_TRY(cudaMallocHost(&h.h_in, kHostInBytes));
    IX_TRY(cudaMallocHost(&h.h_out, kHostOutBytes));
    IX_TRY(cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming));
  }
  for (int i = 0; i < kJobRing; ++i) {
    IX_TRY(cudaEventCreateWithFlags(&ix->job_in[i], cudaEventDisableTiming));
    IX_TRY(cudaEventCreateWithFlags(&ix->job_prep[i], cudaEventDisableTiming));
    IX_TRY(cudaEventCreateWithFlags(&ix->job_scan[i], cudaEventDisableTiming));
    IX_TRY(cudaEventCreateWithFlags(&ix->job_done[i], cudaEventDisableTiming));
    IX_TRY(cudaEventCreateWithFlags(&ix->job_merge[i], cudaEventDisableTiming));
  }
  IX_TRY(cudaEventCreateWithFlags(&ix->rows_ready, cudaEventDisableTiming));
  // prep / merge / exchange kernels are short and gate the next scan: when an SM frees up they go first
  int prio_lo = 0, prio_hi = 0;
  IX_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  IX_TRY(cudaStreamCreateWithPriority(&ix->s_prep, cudaStreamNonBlocking, prio_hi));
  IX_TRY(cudaStreamCreateWithPriority(&ix->s_scan[0], cudaStreamNonBlocking, prio_lo));
  IX_TRY(cudaStreamCreateWithPriority(&ix->s_scan[1], cudaStreamNonBlocking, prio_lo));
  IX_TRY(cudaStreamCreateWithPriority(&ix->s_merge, cudaStreamNonBlocking, prio_hi));
  IX_TRY(cudaStreamCreateWithPriority(&ix->s_xchg, cudaStreamNonBlocking, prio_hi));
  IX_TRY(cudaEventCreateWithFlags(&ix->xchg_last, cudaEventDisableTiming));
  IX_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  IX_TRY(preload_search_kernels());
  IX_TRY(cudaDeviceSynchronize());  // the memsets above ran on the legacy stream; the internal streams do not sync with it
#undef IX_TRY
  int rc = make_tmap(&ix->tmap_rows, ix->rows, f32, (uint64_t)cap_pad, kTileM);
  for (SearchWs& w : ix->ws)
    if (!rc) rc = make_tmap(&w.tmap_q, w.qop, f32, kNQ, kNQ);
  if (rc) {
    free_index(ix);
    return rc;
  }
  *out = ix;
  return FRS_OK;
}

extern "C" int frs_index_destroy(frs_index* idx) {
  if (!idx) return FRS_OK;
  cudaSetDevice(idx->device);
  cudaDeviceSynchronize();
  free_index(idx);
  return FRS_OK;
}

extern "C" int64_t frs_index_size(const frs_index* idx) { return idx ? idx->size : 0; }
extern "C" int64_t frs_index_capacity(const frs_index* idx) { return idx ? idx->capacity : 0; }
extern "C" int frs_index_dtype(const frs_index* idx) { return idx ? idx->dtype : FRS_E_INVALID; }
extern "C" int frs_index_device(const frs_index* idx) { return idx ? idx->device : FRS_E_INVALID; }
extern "C" int frs_index_set_base(frs_index* idx, int64_t base) {
  if (!idx) return set_err(FRS_E_INVALID, "idx is null");
  idx->base = base;
  return FRS_OK;
}
extern "C" int frs_index_set_scan_grid(frs_index* idx, int grid) {
  if (!idx || grid < 0) return set_err(FRS_E_INVALID, "bad argument");
  idx->grid_override = grid;
  return FRS_OK;
}
// The scan kernel is persistent with one CTA per SM and nearly all of an SM's shared memory, so nothing else fits
// next to one of its CTAs.  The pipelined entry points therefore run it on (SMs - reserve) CTAs: the prep kernel of
// the next batch and the merge / exchange kernels of the previous one run on the SMs left over.
extern "C" int frs_index_set_pipeline_reserve(frs_index* idx, int sms) {
  if (!idx || sms < -1 || sms > 64) return set_err(FRS_E_INVALID, "bad argument");
  idx->pipe_reserve = sms;
  return FRS_OK;
}
// experiment knob: 1 = every scan on one stream (launches strictly one after the other), 2 = alternate (default)
extern "C" int frs_index_set_scan_streams(frs_index* idx, int n) {
  if (!idx || n < 1 || n > 2) return set_err(FRS_E_INVALID, "bad argument");
  idx->scan_streams = n;
  return FRS_OK;
}
extern "C" void* frs_index_rows_ptr(frs_index* idx) { return idx ? idx->rows : nullptr; }
extern "C" uint32_t* frs_index_codes_ptr(frs_index* idx) { return idx ? idx->codes : nullptr; }
extern "C" int frs_index_set_size(frs_index* idx, int64_t n) {
  if (!idx || n < 0 || n > idx->capacity) return set_err(FRS_E_INVALID, "size out of range");
  std::lock_guard<std::mutex> lk(idx->mu);
  idx->size = n;
  return FRS_OK;
}

// ---------------------------------------------------------------------------------------------
// write path
// ---------------------------------------------------------------------------------------------
// Ordering of writes against searches (the reference serves queries from up to 25 threads while ingest.py is
// upserting, main2.py:52-53): a write kernel is chained behind the previous write (`rows_ready`) and re-records
// that event; every search waits on it before its first kernel, so a search whose size snapshot includes new
// rows never reads them half-written.  An in-place overwrite additionally waits for the searches in flight.
static int write_begin(frs_index* ix, cudaStream_t st, bool in_place) {
  ix->writes_pending = true;
  CU_TRY(cudaStreamWaitEvent(st, ix->rows_ready, 0));
  if (in_place)
    for (SearchWs& w : ix->ws) CU_TRY(cudaStreamWaitEvent(st, w.free, 0));
  return FRS_OK;
}

extern "C" int frs_index_set_rows(frs_index* idx, int64_t row0, const float* dev_vecs,
                                  const uint32_t* dev_codes, int64_t n, void* stream) {
  if (!idx || n < 0 || row0 < 0 || (n > 0 && !dev_vecs)) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (row0 + n > idx->size) return set_err(FRS_E_INVALID, "rows [%lld,%lld) beyond size %lld", (long long)row0,
                                           (long long)(row0 + n), (long long)idx->size);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = write_begin(idx, st, true);
  if (rc) return rc;
  // dev_codes == null keeps the stored payload codes
  CU_TRY(launch_store_rows(idx->f32(), dev_vecs, dev_codes, n,
                           (char*)idx->rows + (size_t)row0 * idx->row_bytes(), dev_codes ? idx->codes + row0 : nullptr, st));
  CU_TRY(cudaEventRecord(idx->rows_ready, st));
  return FRS_OK;
}

extern "C" int frs_index_add(frs_index* idx, const float* dev_vecs, const uint32_t* dev_codes, int64_t n,
                             void* stream) {
  if (!idx || n < 0 || (n > 0 && !dev_vecs)) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (idx->size + n > idx->capacity)
    return set_err(FRS_E_CAPACITY, "index full: size %lld + %lld > capacity %lld", (long long)idx->size,
                   (long long)n, (long long)idx->capacity);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = write_begin(idx, st, false);
  if (rc) return rc;
  CU_TRY(launch_store_rows(idx->f32(), dev_vecs, dev_codes, n,
                           (char*)idx->rows + (size_t)idx->size * idx->row_bytes(),
                           idx->codes + idx->size, st));
  CU_TRY(cudaEventRecord(idx->rows_ready, st));
  idx->size += n;  // published under the lock; searches wait on rows_ready before reading the new rows
  return FRS_OK;
}

extern "C" int frs_index_add_host(frs_index* idx, const float* host_vecs, const uint32_t* host_codes,
                                  int64_t n) {
  if (!idx || n < 0 || (n > 0 && !host_vecs)) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  {
    std::lock_guard<std::mutex> lk(idx->mu);  // early verdict; frs_index_add re-checks under the same lock
    if (idx->size + n > idx->capacity)
      return set_err(FRS_E_CAPACITY, "index full: size %lld + %lld > capacity %lld", (long long)idx->size,
                     (long long)n, (long long)idx->capacity);
  }
  const int64_t chunk = 16384;
  float* d_v = nullptr;
  uint32_t* d_c = nullptr;
  cudaStream_t st = nullptr;  // a stream of this call: concurrent add_host callers do not share staging
  CU_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaError_t e = cudaMalloc(&d_v, (size_t)chunk * kDim * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_c, (size_t)chunk * 4);
  if (e != cudaSuccess) {
    cudaFree(d_v);
    cudaStreamDestroy(st);
    return set_err(FRS_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  int rc = FRS_OK;
  for (int64_t o = 0; o < n && rc == FRS_OK; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    e = cudaMemcpyAsync(d_v, host_vecs + o * kDim, (size_t)m * kDim * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && host_codes)
      e = cudaMemcpyAsync(d_c, host_codes + o, (size_t)m * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      rc = set_err(FRS_E_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
      break;
    }
    rc = frs_index_add(idx, d_v, host_codes ? d_c : nullptr, m, st);
    if (rc == FRS_OK) {
      e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = set_err(FRS_E_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    }
  }
  cudaFree(d_v);
  cudaFree(d_c);
  cudaStreamDestroy(st);
  return rc;
}

extern "C" int frs_index_set_codes(frs_index* idx, int64_t row0, const uint32_t* dev_codes, int64_t n,
                                   void* stream) {
  if (!idx || n < 0 || row0 < 0 || (n > 0 && !dev_codes)) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (row0 + n > idx->size) return set_err(FRS_E_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = write_begin(idx, st, true);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(idx->codes + row0, dev_codes, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  CU_TRY(cudaEventRecord(idx->rows_ready, st));
  return FRS_OK;
}

extern "C" int frs_index_read_rows(frs_index* idx, int64_t row0, int64_t n, float* dev_out, void* stream) {
  if (!idx || n < 0 || row0 < 0 || row0 + n > idx->size || (n > 0 && !dev_out))
    return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(cudaStreamWaitEvent((cudaStream_t)stream, idx->rows_ready, 0));
  CU_TRY(launch_read_rows(idx->f32(), (const char*)idx->rows + (size_t)row0 * idx->row_bytes(), n, dev_out,
                          (cudaStream_t)stream));
  return FRS_OK;
}

extern "C" int frs_index_read_rows_host(frs_index* idx, int64_t row0, int64_t n, float* host_out) {
  if (!idx || n < 0 || row0 < 0 || row0 + n > idx->size || (n > 0 && !host_out))
    return set_err(FRS_E_INVALID, "bad argument");
  if (n == 0) return FRS_OK;
  CU_TRY(cudaSetDevice(idx->device));
  float* d = nullptr;
  CU_TRY(cudaMalloc(&d, (size_t)n * kDim * 4));
  int rc = frs_index_read_rows(idx, row0, n, d, idx->stream);
  if (rc == FRS_OK) {
    cudaError_t e = cudaMemcpyAsync(host_out, d, (size_t)n * kDim * 4, cudaMemcpyDeviceToHost, idx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
    if (e != cudaSuccess) rc = set_err(FRS_E_CUDA, "read back: %s", cudaGetErrorString(e));
  }
  cudaFree(d);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// persistence: raw storage rows + payload codes
// ---------------------------------------------------------------------------------------------
extern "C" int frs_index_export_raw(frs_index* idx, int64_t row0, int64_t n, void* host_rows, uint32_t* host_codes) {
  if (!idx || n < 0 || row0 < 0 || row0 + n > idx->size || (n > 0 && (!host_rows || !host_codes)))
    return set_err(FRS_E_INVALID, "bad argument");
  if (n == 0) return FRS_OK;
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  CU_TRY(cudaEventSynchronize(idx->rows_ready));
  CU_TRY(cudaMemcpy(host_rows, (const char*)idx->rows + (size_t)row0 * idx->row_bytes(), (size_t)n * idx->row_bytes(),
                    cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(host_codes, idx->codes + row0, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return FRS_OK;
}

extern "C" int frs_index_import_raw(frs_index* idx, const void* host_rows, const uint32_t* host_codes, int64_t n) {
  if (!idx || n < 0 || (n > 0 && (!host_rows || !host_codes))) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (idx->size + n > idx->capacity)
    return set_err(FRS_E_CAPACITY, "index full: size %lld + %lld > capacity %lld", (long long)idx->size, (long long)n,
                   (long long)idx->capacity);
  if (n == 0) return FRS_OK;
  CU_TRY(cudaMemcpy((char*)idx->rows + (size_t)idx->size * idx->row_bytes(), host_rows, (size_t)n * idx->row_bytes(),
                    cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(idx->codes + idx->size, host_codes, (size_t)n * 4, cudaMemcpyHostToDevice));
  idx->size += n;
  return FRS_OK;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
static int scan_grid(const frs_index* ix, uint32_t num_tiles, int reserve = 0) {
  if (reserve < 0) {
    // auto: the prep and merge kernels of the neighbouring batches must finish inside one scan period on the SMs the
    // scan leaves free, so short scans (small shards) leave more (measured on B200: 10M rows/GPU best with 4,
    // 1.25M rows/GPU with 12)
    const uint32_t per_sm = num_tiles / (uint32_t)(ix->sm_count > 0 ? ix->sm_count : 1);
    reserve = per_sm >= 384 ? 4 : per_sm >= 160 ? 8 : 12;
  }
  int g = ix->grid_override > 0 ? ix->grid_override : ix->sm_count - (ix->sm_count > 4 * reserve ? reserve : 0);
  if (g > ix->max_parts) g = ix->max_parts;
  if ((uint32_t)g > num_tiles) g = (int)num_tiles;
  return g;
}

namespace frs {

// prep -> scan -> merge.  Caller holds ix->mu and has selected the device.  Exactly one of out_s32 / out_s64 may be null.
int search_enqueue(frs_index* ix, const SearchArgs& a, const SearchLaunch& L) {
  if (a.nq < 1 || a.nq > kNQ) return set_err(FRS_E_INVALID, "nq must be in [1,%d] (got %d)", kNQ, a.nq);
  if (a.k < 1 || a.k > kMaxK) return set_err(FRS_E_INVALID, "k must be in [1,%d] (got %d)", kMaxK, a.k);
  if (!a.q || !a.code || !a.mask || !a.out_ids || (!a.out_s32 && !a.out_s64))
    return set_err(FRS_E_INVALID, "null pointer argument");
  const bool f32 = ix->f32();
  const float eps = f32 ? kEpsTF32 : kEpsBF16;
  SearchWs& w = ix->ws[ix->ws_next];
  ix->ws_last = ix->ws_next;
  ix->ws_next = (ix->ws_next + 1) % kWsRing;
  cudaEvent_t* pev = ix->prof_mode == 3 ? nullptr : L.prof;
  const bool bracket = ix->prof_mode == 3 && ix->br_first;
  // the workspace's previous search (kWsRing calls ago) and the pending writes come first
  CU_TRY(cudaStreamWaitEvent(L.prep, w.free, 0));
  if (ix->writes_pending) {  // (skipped while the store is read-only: one driver call less per search)
    CU_TRY(cudaStreamWaitEvent(L.prep, ix->rows_ready, 0));
    if (cudaEventQuery(ix->rows_ready) == cudaSuccess) ix->writes_pending = false;
  }
  int launches = 0;
  if (pev) CU_TRY(cudaEventRecord(pev[0], L.prep));
  CU_TRY(launch_prep_queries(f32, a.q, a.code, a.mask, a.nq, w.qop, w.qrec, w.qcode, w.qmask, w.stats, w.gmax,
                             w.gsample, ix->rows, ix->codes, (uint32_t)ix->size, a.k, eps, L.prep));
  launches++;
  if (pev) CU_TRY(cudaEventRecord(pev[1], L.prep));
  if (L.scan != L.prep) {
    CU_TRY(cudaEventRecord(L.ev_prep, L.prep));
    CU_TRY(cudaStreamWaitEvent(L.scan, L.ev_prep, 0));
  }
  const uint32_t n = (uint32_t)ix->size;
  uint32_t num_tiles = (n + kTileM - 1) / kTileM;
  // restricted scan: only the listed tiles (the caller guarantees that every row matching any query's
  // predicate lies in one of them).  A list too long for the per-CTA table falls back to the full scan.
  const uint32_t* tile_ids = a.tile_ids;
  if (tile_ids) {
    const int g = scan_grid(ix, (uint32_t)a.n_tile_ids, L.reserve_sms);
    if (a.n_tile_ids >= 0 && a.n_tile_ids <= (int64_t)num_tiles && (g == 0 || (a.n_tile_ids + g - 1) / g <= kMaxTileSlots))
      num_tiles = (uint32_t)a.n_tile_ids;
    else
      tile_ids = nullptr;
  }
  const int grid = scan_grid(ix, num_tiles, L.reserve_sms);
  if (pev) CU_TRY(cudaEventRecord(pev[2], L.scan));
  if (bracket && ix->br_count == 0) CU_TRY(cudaEventRecord(ix->br_first, L.scan));
  if (grid > 0) {
    ScanParams sp{};
    sp.rows = ix->rows;
    sp.codes = ix->codes;
    sp.n = n;
    sp.num_tiles = num_tiles;
    sp.tile_ids = tile_ids;
    sp.qrec = w.qrec;
    sp.qcode = w.qcode;
    sp.qmask = w.qmask;
    sp.nq = a.nq;
    sp.k = a.k;
    sp.eps = eps;
    sp.part_keys = w.part_keys;
    sp.part_cnt = w.part_cnt;
    sp.dbg_scores = nullptr;
    sp.gmax = w.gmax;
    sp.tau0 = w.gsample + kSampleTau0;
    sp.stats = w.stats;
    sp.timeline = ix->prof_mode == 2 ? ix->timeline : nullptr;
    if (sp.timeline) CU_TRY(cudaMemsetAsync(ix->timeline, 0, (size_t)kGmaxPad * 16 * 8, L.scan));
    CU_TRY(launch_scan(f32, false, grid, ix->tmap_rows, w.tmap_q, sp, L.scan));
    launches++;
  }
  if (pev) CU_TRY(cudaEventRecord(pev[3], L.scan));
  if (bracket) {
    CU_TRY(cudaEventRecord(ix->br_last, L.scan));
    ix->br_count++;
  }
  if (L.merge != L.scan) {
    CU_TRY(cudaEventRecord(L.ev_scan, L.scan));
    CU_TRY(cudaStreamWaitEvent(L.merge, L.ev_scan, 0));
  }
  MergeParams mp{};
  mp.part_keys = w.part_keys;
  mp.part_cnt = w.part_cnt;
  mp.gmax = w.gmax;
  mp.nparts = grid;
  mp.rows = ix->rows;
  mp.qrec = w.qrec;
  mp.nq = a.nq;
  mp.k = a.k;
  mp.eps = eps;
  mp.base = ix->base;
  mp.id_block = ix->id_block;
  mp.id_shards = ix->id_shards;
  mp.id_shard = ix->id_shard;
  mp.out_s64 = a.out_s64;
  mp.out_s32 = a.out_s32;
  mp.out_ids = a.out_ids;
  mp.stats = w.stats;
  mp.spill = w.spill;
  if (a.push) mp.push = *a.push;  // the exchange step rides in the tail of the merge kernel
  if (pev) CU_TRY(cudaEventRecord(pev[4], L.merge));
  CU_TRY(launch_merge(f32, mp, L.merge));
  launches++;
  if (pev) CU_TRY(cudaEventRecord(pev[5], L.merge));
  CU_TRY(cudaEventRecord(w.free, L.merge));
  ix->last_grid = grid;
  ix->last_launches = launches;
  return FRS_OK;
}

static cudaEvent_t* prof_slot(frs_index* ix) {
  if (!ix->prof_mode || ix->prof_mode == 3 || !ix->prof_ev) return nullptr;
  return ix->prof_ev + (size_t)(ix->prof_calls % frs_index::kProfRing) * kProfEvents;
}

int pipelined_begin(frs_index* ix, bool has_in, cudaStream_t in_stream, int* slot, SearchLaunch* L) {
  const int s = (int)(ix->jobs++ % kJobRing);
  if (has_in) {
    CU_TRY(cudaEventRecord(ix->job_in[s], in_stream));
    CU_TRY(cudaStreamWaitEvent(ix->s_prep, ix->job_in[s], 0));
  }
  L->prep = ix->s_prep;
  L->scan = ix->s_scan[ix->scan_streams > 1 ? (int)((ix->jobs - 1) & 1) : 0];
  L->merge = ix->s_merge;
  L->ev_prep = ix->job_prep[s];
  L->ev_scan = ix->job_scan[s];
  L->prof = prof_slot(ix);
  L->reserve_sms = ix->pipe_reserve;
  *slot = s;
  return FRS_OK;
}

int host_slot_acquire(frs_index* ix, int* slot) {
  std::unique_lock<std::mutex> lk(ix->hs_mu);
  int found = -1;
  const bool ok = ix->hs_cv.wait_for(lk, std::chrono::seconds(30), [&] {
    for (int i = 0; i < kHostSlots; ++i)
      if (!ix->hs[i].busy) {
        found = i;
        return true;
      }
    return false;
  });
  if (!ok) return set_err(FRS_E_STATE, "all %d host-call slots of this index stayed busy for 30 s "
                          "(more than %d submits without a collect?)", kHostSlots, kHostSlots);
  ix->hs[found].busy = true;
  *slot = found;
  return FRS_OK;
}

void host_slot_release(frs_index* ix, int slot) {
  {
    std::lock_guard<std::mutex> lk(ix->hs_mu);
    ix->hs[slot].busy = false;
  }
  ix->hs_cv.notify_one();
}

// exchange.cu
int exchange_begin_push(frs_exchange* ex, int nq, int k, PushTarget* t);
void exchange_commit_push(frs_exchange* ex);
int exchange_wait_merge(frs_exchange* ex, int nq, int k, float* out_s, int64_t* out_i, cudaStream_t st);
int exchange_check(frs_exchange* ex);

}  // namespace frs

// Ring-slot rule of the gather buffers (scan.cuh): the push of sequence number s may be issued only after this rank's own
// cross-shard merge of s - 2.  Called before the local pass whose merge kernel carries the push of ex->seq + 1.
static int exchange_order_push(frs_index* ix, frs_exchange* ex) {
  const uint32_t next = ex->seq + 1;
  if (next > 2) CU_TRY(cudaStreamWaitEvent(ix->s_merge, (cudaEvent_t)ex->merged[(next - 2) % kExchangeSlots], 0));
  return FRS_OK;
}

// The cross-shard half of a pipelined sharded search, on its own stream: wait for every rank's push of ex->seq (already
// committed), merge into the caller's outputs.  Leaves s_xchg ordered behind the local merge of this job.
static int exchange_tail(frs_index* ix, frs_exchange* ex, int job, int nq, int k, float* out_s, int64_t* out_i) {
  CU_TRY(cudaEventRecord(ix->job_merge[job], ix->s_merge));
  CU_TRY(cudaStreamWaitEvent(ix->s_xchg, ix->job_merge[job], 0));
  int rc = exchange_wait_merge(ex, nq, k, out_s, out_i, ix->s_xchg);
  if (rc) return rc;
  CU_TRY(cudaEventRecord((cudaEvent_t)ex->merged[ex->seq % kExchangeSlots], ix->s_xchg));
  return FRS_OK;
}

static SearchLaunch in_stream_launch(frs_index* ix, cudaStream_t st) {
  SearchLaunch L{};
  L.prep = L.scan = L.merge = st;
  L.prof = prof_slot(ix);
  return L;
}

static int search_in_stream(frs_index* ix, const SearchArgs& a, cudaStream_t st) {
  if (!ix) return set_err(FRS_E_INVALID, "idx is null");
  CU_TRY(cudaSetDevice(ix->device));
  std::lock_guard<std::mutex> lk(ix->mu);
  SearchLaunch L = in_stream_launch(ix, st);
  int rc = search_enqueue(ix, a, L);
  if (rc == FRS_OK && L.prof) {
    CU_TRY(cudaEventRecord(L.prof[6], st));
    ix->prof_calls++;
  }
  return rc;
}

extern "C" int frs_index_search(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                const uint32_t* dev_q_mask, int nq, int k, float* dev_out_scores,
                                int64_t* dev_out_ids, void* stream) {
  if (!dev_out_scores) return set_err(FRS_E_INVALID, "null pointer argument");
  SearchArgs a;
  a.q = dev_queries; a.code = dev_q_code; a.mask = dev_q_mask; a.nq = nq; a.k = k;
  a.out_s32 = dev_out_scores; a.out_ids = dev_out_ids;
  return search_in_stream(idx, a, (cudaStream_t)stream);
}

extern "C" int frs_index_search_tiles(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                      const uint32_t* dev_q_mask, int nq, int k, const uint32_t* dev_tile_ids,
                                      int64_t n_tiles, float* dev_out_scores, int64_t* dev_out_ids, void* stream) {
  if (!idx || !dev_out_scores || n_tiles < 0 || (n_tiles > 0 && !dev_tile_ids)) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  // an empty list is still a restricted scan (nothing can match): any non-null pointer selects it
  SearchArgs a;
  a.q = dev_queries; a.code = dev_q_code; a.mask = dev_q_mask; a.nq = nq; a.k = k;
  a.out_s32 = dev_out_scores; a.out_ids = dev_out_ids;
  a.tile_ids = n_tiles ? dev_tile_ids : idx->codes;
  a.n_tile_ids = n_tiles;
  return search_in_stream(idx, a, (cudaStream_t)stream);
}

extern "C" int frs_index_search_local(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                      const uint32_t* dev_q_mask, int nq, int k, double* dev_out_scores64,
                                      int64_t* dev_out_ids, void* stream) {
  if (!dev_out_scores64) return set_err(FRS_E_INVALID, "null pointer argument");
  SearchArgs a;
  a.q = dev_queries; a.code = dev_q_code; a.mask = dev_q_mask; a.nq = nq; a.k = k;
  a.out_s64 = dev_out_scores64; a.out_ids = dev_out_ids;
  return search_in_stream(idx, a, (cudaStream_t)stream);
}

static int check_exchange_args(const frs_index* idx, const frs_exchange* ex, int nq, int k) {
  if (!idx || !ex) return set_err(FRS_E_INVALID, "null pointer argument");
  if (!ex->connected) return set_err(FRS_E_INVALID, "exchange is not connected");
  if (nq < 1 || nq > ex->nq_max || k < 1 || k > ex->k_max)
    return set_err(FRS_E_INVALID, "nq / k (%d / %d) exceed the exchange's (%d / %d)", nq, k, ex->nq_max, ex->k_max);
  if (idx->device != ex->device) return set_err(FRS_E_INVALID, "index and exchange live on different devices");
  return FRS_OK;
}

// Local pass of the sharded search WITH the exchange push fused into the merge kernel: the shard's exact top-k
// goes into the exchange's local block and, from the same kernel, into every peer's gather buffer.  Follow with
// frs_exchange_wait_merge.  nq <= nq_max and k <= k_max of the exchange; every rank passes the same nq and k.
extern "C" int frs_index_search_push(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                     const uint32_t* dev_q_mask, int nq, int k, frs_exchange* ex, void* stream) {
  int rc = check_exchange_args(idx, ex, nq, k);
  if (rc) return rc;
  PushTarget t{};
  rc = exchange_begin_push(ex, nq, k, &t);
  if (rc) return rc;
  SearchArgs a;
  a.q = dev_queries; a.code = dev_q_code; a.mask = dev_q_mask; a.nq = nq; a.k = k;
  a.out_s64 = reinterpret_cast<double*>(ex->local);
  a.out_ids = reinterpret_cast<int64_t*>(ex->local) + ex->plane_words;
  a.push = &t;
  rc = search_in_stream(idx, a, (cudaStream_t)stream);
  if (rc == FRS_OK) exchange_commit_push(ex);  // the sequence number advances only once the push is really enqueued
  return rc;
}

// Pipelined search (device buffers): prep / scan / merge (+ exchange) run on the index's internal streams, so the
// prep of call i+1 and the merge + exchange of call i-1 overlap the scan of call i.
extern "C" int frs_index_search_async(frs_index* idx, frs_exchange* ex, const float* dev_queries,
                                      const uint32_t* dev_q_code, const uint32_t* dev_q_mask, int nq, int k,
                                      float* dev_out_scores, int64_t* dev_out_ids, void* in_stream, int* ticket) {
  if (!idx || !dev_out_scores || !dev_out_ids || !ticket) return set_err(FRS_E_INVALID, "null pointer argument");
  int rc = ex ? check_exchange_args(idx, ex, nq, k) : FRS_OK;
  if (rc) return rc;
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  SearchArgs a;
  a.q = dev_queries; a.code = dev_q_code; a.mask = dev_q_mask; a.nq = nq; a.k = k;
  PushTarget t{};
  if (ex) {
    rc = exchange_begin_push(ex, nq, k, &t);  // refuses a poisoned exchange before anything is enqueued
    if (rc) return rc;
    a.out_s64 = reinterpret_cast<double*>(ex->local);
    a.out_ids = reinterpret_cast<int64_t*>(ex->local) + ex->plane_words;
    a.push = &t;
  } else {
    a.out_s32 = dev_out_scores;
    a.out_ids = dev_out_ids;
  }
  SearchLaunch L{};
  int slot = 0;
  rc = pipelined_begin(idx, true, (cudaStream_t)in_stream, &slot, &L);
  if (rc) return rc;
  if (ex && (rc = exchange_order_push(idx, ex))) return rc;
  rc = search_enqueue(idx, a, L);
  if (rc) return rc;
  cudaStream_t tail = idx->s_merge;  // the stream the job's last operation runs on
  if (ex) {
    exchange_commit_push(ex);
    rc = exchange_tail(idx, ex, slot, nq, k, dev_out_scores, dev_out_ids);
    if (rc) return rc;
    tail = idx->s_xchg;
  }
  if (L.prof) {
    CU_TRY(cudaEventRecord(L.prof[6], tail));
    idx->prof_calls++;
  }
  CU_TRY(cudaEventRecord(idx->job_done[slot], tail));
  if (ex) {
    CU_TRY(cudaEventRecord(idx->xchg_last, tail));
    idx->xchg_used = true;
  }
  idx->last_launches += ex ? 1 : 0;
  *ticket = slot;
  return FRS_OK;
}

extern "C" int frs_index_wait(frs_index* idx, int ticket, void* stream) {
  if (!idx || ticket >= kJobRing) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (ticket < 0) {
    if (idx->jobs == 0) return FRS_OK;
    ticket = (int)((idx->jobs - 1) % kJobRing);  // each tail stream is in order: its newest job finishes last
    if (idx->xchg_used) CU_TRY(cudaStreamWaitEvent((cudaStream_t)stream, idx->xchg_last, 0));
  }
  CU_TRY(cudaStreamWaitEvent((cudaStream_t)stream, idx->job_done[ticket], 0));
  return FRS_OK;
}

extern "C" int frs_index_sync(frs_index* idx, int ticket) {
  if (!idx || ticket >= kJobRing) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  cudaEvent_t ev, ev2 = nullptr;
  {
    std::lock_guard<std::mutex> lk(idx->mu);
    if (ticket < 0) {
      if (idx->jobs == 0) return FRS_OK;
      ticket = (int)((idx->jobs - 1) % kJobRing);
      if (idx->xchg_used) ev2 = idx->xchg_last;
    }
    ev = idx->job_done[ticket];
  }
  CU_TRY(cudaEventSynchronize(ev));
  if (ev2) CU_TRY(cudaEventSynchronize(ev2));
  return FRS_OK;
}

// Host-buffer search, split in two so that a caller (or several threads) can keep a few batches in flight: the
// H2D copy of batch i+1 and the D2H copy of batch i-1 overlap the scan of batch i.  With `ex` the local result
// is exchanged with the other ranks (every rank submits the same batches in the same order).
extern "C" int frs_index_search_host_submit(frs_index* idx, frs_exchange* ex, const float* host_queries,
                                            const uint32_t* host_q_code, const uint32_t* host_q_mask, int nq, int k,
                                            int* ticket) {
  if (!idx) return set_err(FRS_E_INVALID, "idx is null");
  if (nq < 1 || nq > kNQ) return set_err(FRS_E_INVALID, "nq must be in [1,%d] (got %d)", kNQ, nq);
  if (k < 1 || k > kMaxK) return set_err(FRS_E_INVALID, "k must be in [1,%d] (got %d)", kMaxK, k);
  if (!host_queries || !host_q_code || !host_q_mask || !ticket) return set_err(FRS_E_INVALID, "null pointer argument");
  int rc = ex ? check_exchange_args(idx, ex, nq, k) : FRS_OK;
  if (rc) return rc;
  CU_TRY(cudaSetDevice(idx->device));
  int hsi = 0;
  rc = host_slot_acquire(idx, &hsi);
  if (rc) return rc;
  HostSlot& h = idx->hs[hsi];
  const size_t qb = (size_t)nq * kDim * 4;
  memcpy(h.h_in, host_queries, qb);
  memcpy(h.h_in + qb, host_q_code, (size_t)nq * 4);
  memcpy(h.h_in + qb + (size_t)nq * 4, host_q_mask, (size_t)nq * 4);
  h.nq = nq;
  h.k = k;
  auto fail = [&](int code) {
    host_slot_release(idx, hsi);
    return code;
  };
  std::lock_guard<std::mutex> lk(idx->mu);
  int64_t* d_ids = reinterpret_cast<int64_t*>(h.d_out);
  float* d_scores = reinterpret_cast<float*>(h.d_out + (size_t)nq * k * 8);
  SearchArgs a;
  a.q = reinterpret_cast<const float*>(h.d_in);
  a.code = reinterpret_cast<const uint32_t*>(h.d_in + qb);
  a.mask = a.code + nq;
  a.nq = nq;
  a.k = k;
  PushTarget t{};
  if (ex) {
    rc = exchange_begin_push(ex, nq, k, &t);
    if (rc) return fail(rc);
    a.out_s64 = reinterpret_cast<double*>(ex->local);
    a.out_ids = reinterpret_cast<int64_t*>(ex->local) + ex->plane_words;
    a.push = &t;
  } else {
    a.out_s32 = d_scores;
    a.out_ids = d_ids;
  }
  SearchLaunch L{};
  int slot = 0;
  rc = pipelined_begin(idx, false, nullptr, &slot, &L);
  if (rc) return fail(rc);
  cudaError_t e = cudaMemcpyAsync(h.d_in, h.h_in, qb + (size_t)nq * 8, cudaMemcpyHostToDevice, idx->s_prep);
  if (e != cudaSuccess) return fail(set_err(FRS_E_CUDA, "cudaMemcpyAsync (queries): %s", cudaGetErrorString(e)));
  if (ex && (rc = exchange_order_push(idx, ex))) return fail(rc);
  rc = search_enqueue(idx, a, L);
  if (rc) return fail(rc);
  cudaStream_t tail = idx->s_merge;
  if (ex) {
    exchange_commit_push(ex);
    rc = exchange_tail(idx, ex, slot, nq, k, d_scores, d_ids);
    if (rc) return fail(rc);
    tail = idx->s_xchg;
  }
  if (L.prof) {
    cudaEventRecord(L.prof[6], tail);
    idx->prof_calls++;
  }
  e = cudaMemcpyAsync(h.h_out, h.d_out, (size_t)nq * k * 12, cudaMemcpyDeviceToHost, tail);
  if (e == cudaSuccess) e = cudaEventRecord(h.done, tail);
  if (e == cudaSuccess) e = cudaEventRecord(idx->job_done[slot], tail);
  if (e == cudaSuccess && ex) {
    e = cudaEventRecord(idx->xchg_last, tail);
    idx->xchg_used = true;
  }
  if (e != cudaSuccess) return fail(set_err(FRS_E_CUDA, "cudaMemcpyAsync (results): %s", cudaGetErrorString(e)));
  idx->last_launches += ex ? 1 : 0;
  *ticket = hsi;
  return FRS_OK;
}

extern "C" int frs_index_search_host_collect(frs_index* idx, frs_exchange* ex, int ticket, float* host_out_scores,
                                             int64_t* host_out_ids) {
  if (!idx || ticket < 0 || ticket >= kHostSlots || !host_out_scores || !host_out_ids)
    return set_err(FRS_E_INVALID, "bad argument");
  HostSlot& h = idx->hs[ticket];
  if (!h.busy) return set_err(FRS_E_STATE, "ticket %d is not in flight", ticket);
  CU_TRY(cudaSetDevice(idx->device));
  cudaError_t e = cudaEventSynchronize(h.done);
  int rc = FRS_OK;
  if (e != cudaSuccess) rc = set_err(FRS_E_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
  if (rc == FRS_OK && ex) rc = exchange_check(ex);
  if (rc == FRS_OK) {
    memcpy(host_out_ids, h.h_out, (size_t)h.nq * h.k * 8);
    memcpy(host_out_scores, h.h_out + (size_t)h.nq * h.k * 8, (size_t)h.nq * h.k * 4);
  }
  host_slot_release(idx, ticket);
  return rc;
}

extern "C" int frs_index_search_host(frs_index* idx, const float* host_queries, const uint32_t* host_q_code,
                                     const uint32_t* host_q_mask, int nq, int k, float* host_out_scores,
                                     int64_t* host_out_ids) {
  if (!host_out_scores || !host_out_ids) return set_err(FRS_E_INVALID, "null pointer argument");
  int ticket = -1;
  int rc = frs_index_search_host_submit(idx, nullptr, host_queries, host_q_code, host_q_mask, nq, k, &ticket);
  if (rc) return rc;
  return frs_index_search_host_collect(idx, nullptr, ticket, host_out_scores, host_out_ids);
}

extern "C" int frs_merge_shards(int device, const double* dev_scores64, const int64_t* dev_ids, int n_shards,
                                int nq, int k, float* dev_out_scores, int64_t* dev_out_ids, void* stream) {
  if (!dev_scores64 || !dev_ids || !dev_out_scores || !dev_out_ids) return set_err(FRS_E_INVALID, "null pointer argument");
  if (n_shards < 1 || nq < 1 || k < 1 || k > kMaxK) return set_err(FRS_E_INVALID, "bad shape");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(launch_merge_shards(dev_scores64, dev_ids, n_shards, nq, k, (size_t)nq * k, dev_out_scores, dev_out_ids,
                             (cudaStream_t)stream));
  return FRS_OK;
}

// Same merge over the exchange buffer of ShardedIndex: [n_shards][2][nq][k] 64-bit words, plane 0 =
// fp64 score bits, plane 1 = int64 global ids (one all-gather moves both).
extern "C" int frs_merge_shards_packed(int device, const int64_t* dev_packed, int n_shards, int nq, int k,
                                       float* dev_out_scores, int64_t* dev_out_ids, void* stream) {
  if (!dev_packed || !dev_out_scores || !dev_out_ids) return set_err(FRS_E_INVALID, "null pointer argument");
  if (n_shards < 1 || nq < 1 || k < 1 || k > kMaxK) return set_err(FRS_E_INVALID, "bad shape");
  CU_TRY(cudaSetDevice(device));
  const size_t plane = (size_t)nq * k;
  CU_TRY(launch_merge_shards(reinterpret_cast<const double*>(dev_packed), dev_packed + plane, n_shards, nq, k,
                             2 * plane, dev_out_scores, dev_out_ids, (cudaStream_t)stream));
  return FRS_OK;
}

extern "C" int frs_index_last_queries(frs_index* idx, float* dev_out, void* stream) {
  if (!idx || !dev_out) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  const SearchWs& w = idx->ws[idx->ws_last];
  CU_TRY(cudaStreamWaitEvent((cudaStream_t)stream, w.free, 0));
  CU_TRY(cudaMemcpyAsync(dev_out, w.qrec, (size_t)kNQ * kDim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return FRS_OK;
}

extern "C" int frs_index_debug_scores(frs_index* idx, const float* dev_queries, int nq, float* dev_out,
                                      void* stream) {
  if (!idx || !dev_queries || !dev_out) return set_err(FRS_E_INVALID, "bad argument");
  if (nq < 1 || nq > kNQ) return set_err(FRS_E_INVALID, "nq must be in [1,%d]", kNQ);
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  cudaStream_t st = (cudaStream_t)stream;
  const bool f32 = idx->f32();
  SearchWs& w = idx->ws[idx->ws_next];
  idx->ws_last = idx->ws_next;
  idx->ws_next = (idx->ws_next + 1) % kWsRing;
  CU_TRY(cudaStreamWaitEvent(st, w.free, 0));
  CU_TRY(cudaStreamWaitEvent(st, idx->rows_ready, 0));
  CU_TRY(cudaMemsetAsync(w.qcode, 0, kNQ * 4, st));  // all-zero (code, mask): every row passes the predicate
  CU_TRY(launch_prep_queries(f32, dev_queries, w.qcode, w.qcode, nq, w.qop, w.qrec, w.qcode, w.qmask, w.stats, w.gmax,
                             w.gsample, idx->rows, idx->codes, (uint32_t)idx->size, 1, 0.f, st));
  const uint32_t n = (uint32_t)idx->size;
  const uint32_t num_tiles = (n + kTileM - 1) / kTileM;
  const int grid = scan_grid(idx, num_tiles);
  if (grid > 0) {
    ScanParams sp{};
    sp.rows = idx->rows;
    sp.codes = idx->codes;
    sp.n = n;
    sp.num_tiles = num_tiles;
    sp.qrec = w.qrec;
    sp.qcode = w.qcode;
    sp.qmask = w.qmask;
    sp.nq = nq;
    sp.k = 1;
    sp.eps = 0.f;
    sp.dbg_scores = dev_out;
    sp.gmax = w.gmax;
    sp.tau0 = w.gsample + kSampleTau0;
    sp.stats = w.stats;
    CU_TRY(launch_scan(f32, true, grid, idx->tmap_rows, w.tmap_q, sp, st));
  }
  CU_TRY(cudaEventRecord(w.free, st));
  return FRS_OK;
}

extern "C" int frs_index_last_stats(frs_index* idx, int64_t* host_out6) {
  if (!idx || !host_out6) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long h[kStatSlots];
  const SearchWs* w;
  {
    std::lock_guard<std::mutex> lk(idx->mu);
    w = &idx->ws[idx->ws_last];
  }
  CU_TRY(cudaEventSynchronize(w->free));
  CU_TRY(cudaMemcpy(h, w->stats, sizeof(h), cudaMemcpyDeviceToHost));
  host_out6[0] = (int64_t)h[kStatAppended];
  host_out6[1] = (int64_t)h[kStatCompactions];
  host_out6[2] = (int64_t)h[kStatResolutions];
  host_out6[3] = (int64_t)h[kStatRescored];
  host_out6[4] = idx->last_grid;
  host_out6[5] = idx->last_launches;
  return FRS_OK;
}

// ---------------------------------------------------------------------------------------------
// profiling
// ---------------------------------------------------------------------------------------------
extern "C" int frs_index_set_profiling(frs_index* idx, int mode) {
  if (!idx || mode < 0 || mode > 3) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (mode == 3 && !idx->br_first) {
    CU_TRY(cudaEventCreate(&idx->br_first));
    CU_TRY(cudaEventCreate(&idx->br_last));
  }
  idx->br_count = 0;
  if (mode && mode != 3 && !idx->prof_ev) {
    idx->prof_ev = new (std::nothrow) cudaEvent_t[frs_index::kProfRing * kProfEvents];
    if (!idx->prof_ev) return set_err(FRS_E_INVALID, "out of host memory");
    for (int i = 0; i < frs_index::kProfRing * kProfEvents; ++i) CU_TRY(cudaEventCreate(&idx->prof_ev[i]));
  }
  idx->prof_mode = mode;
  idx->prof_calls = 0;
  return FRS_OK;
}

// Bracket profiling (mode 3) relative to the caller's own events: out2 = {ms from `ev_before` to the start of the first scan
// kernel, ms from the end of the last scan kernel to `ev_after`} — the fill and the drain of a pipelined run.  Both events
// must have been recorded on this device (cudaEvent_t handles, e.g. torch.cuda.Event.cuda_event) and completed.
extern "C" int frs_index_read_profile_bracket_rel(frs_index* idx, void* ev_before, void* ev_after, double* host_out2) {
  if (!idx || !ev_before || !ev_after || !host_out2) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  host_out2[0] = host_out2[1] = 0.0;
  if (idx->prof_mode != 3 || !idx->br_first || idx->br_count == 0) return set_err(FRS_E_INVALID, "no bracket recorded (set_profiling(3), then search)");
  CU_TRY(cudaEventSynchronize(idx->br_last));
  CU_TRY(cudaEventSynchronize((cudaEvent_t)ev_after));
  float a = 0.f, b = 0.f;
  CU_TRY(cudaEventElapsedTime(&a, (cudaEvent_t)ev_before, idx->br_first));
  CU_TRY(cudaEventElapsedTime(&b, idx->br_last, (cudaEvent_t)ev_after));
  host_out2[0] = a;
  host_out2[1] = b;
  return FRS_OK;
}

// out8: {searches, prep ms, scan ms, merge ms, exchange ms (cross-shard wait + merge; 0 for a plain search),
//        scan-stream gap ms (end of one scan kernel -> start of the next, summed over consecutive searches),
//        span ms (first prep start -> last search end), 0}
extern "C" int frs_index_read_profile_ex(frs_index* idx, double* host_out8) {
  if (!idx || !host_out8) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  for (int i = 0; i < 8; ++i) host_out8[i] = 0.0;
  if (idx->prof_mode == 3) {  // bracket: {scans, 0, bracket ms (first scan start -> last scan end), 0...}
    if (idx->br_count == 0) return FRS_OK;
    CU_TRY(cudaEventSynchronize(idx->br_last));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, idx->br_first, idx->br_last));
    host_out8[0] = idx->br_count;
    host_out8[2] = ms;
    host_out8[6] = ms;
    idx->br_count = 0;
    return FRS_OK;
  }
  if (!idx->prof_ev || idx->prof_calls == 0) return FRS_OK;
  const int n = idx->prof_calls < frs_index::kProfRing ? idx->prof_calls : frs_index::kProfRing;
  auto evs = [&](int c) {  // c = 0 is the newest recorded search
    return idx->prof_ev + (size_t)((idx->prof_calls - 1 - c) % frs_index::kProfRing) * kProfEvents;
  };
  CU_TRY(cudaEventSynchronize(evs(0)[6]));
  for (int c = 0; c < n; ++c) {
    cudaEvent_t* ev = evs(c);
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, ev[0], ev[1]));
    host_out8[1] += ms;
    CU_TRY(cudaEventElapsedTime(&ms, ev[2], ev[3]));
    host_out8[2] += ms;
    CU_TRY(cudaEventElapsedTime(&ms, ev[4], ev[5]));
    host_out8[3] += ms;
    CU_TRY(cudaEventElapsedTime(&ms, ev[5], ev[6]));
    host_out8[4] += ms;
    if (c + 1 < n) {
      CU_TRY(cudaEventElapsedTime(&ms, evs(c + 1)[3], ev[2]));
      host_out8[5] += ms > 0.f ? ms : 0.f;
    }
  }
  float span = 0.f;
  CU_TRY(cudaEventElapsedTime(&span, evs(n - 1)[0], evs(0)[6]));
  host_out8[6] = span;
  host_out8[0] = n;
  idx->prof_calls = 0;
  return FRS_OK;
}

// Raw time line of the recorded searches (oldest first): per search 7 event times in ms relative to the first search's
// first event — {prep start, prep end, scan start, scan end, merge start, merge end, exchange end}.  Returns the count.
extern "C" int frs_index_read_profile_raw(frs_index* idx, double* host_out, int max_searches) {
  if (!idx || !host_out || max_searches < 1) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (!idx->prof_ev || idx->prof_calls == 0 || idx->prof_mode == 3) return 0;
  int n = idx->prof_calls < frs_index::kProfRing ? idx->prof_calls : frs_index::kProfRing;
  if (n > max_searches) n = max_searches;
  auto evs = [&](int c) { return idx->prof_ev + (size_t)((idx->prof_calls - n + c) % frs_index::kProfRing) * kProfEvents; };
  CU_TRY(cudaEventSynchronize(evs(n - 1)[6]));
  for (int c = 0; c < n; ++c)
    for (int j = 0; j < 7; ++j) {
      float ms = 0.f;
      CU_TRY(cudaEventElapsedTime(&ms, evs(0)[0], evs(c)[j]));
      host_out[c * 7 + j] = ms;
    }
  return n;
}

extern "C" int frs_index_read_profile(frs_index* idx, double* host_out4) {
  if (!host_out4) return set_err(FRS_E_INVALID, "bad argument");
  double o[8];
  int rc = frs_index_read_profile_ex(idx, o);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) host_out4[i] = o[i];
  return FRS_OK;
}

extern "C" int frs_index_read_timeline(frs_index* idx, uint64_t* host_out, int n_ctas) {
  if (!idx || !host_out || n_ctas < 0 || n_ctas > kGmaxPad) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(host_out, idx->timeline, (size_t)n_ctas * 16 * 8, cudaMemcpyDeviceToHost));
  return FRS_OK;
}

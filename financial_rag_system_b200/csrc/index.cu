// index.cu — C ABI of the chunk store (include/frs_b200.h): the sharded, GPU-resident replacement
// of the Qdrant collection used by the reference (create_collection ingest.py:86-96, upsert
// ingest.py:171-175, query_points main.py:232-237 / main2.py:163).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/frs_b200.h"
#include "exchange.cuh"
#include "scan.cuh"

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
// index object
// ---------------------------------------------------------------------------------------------
struct frs_index {
  int device = 0;
  int dtype = FRS_DTYPE_BF16;
  int64_t capacity = 0;
  int64_t size = 0;
  int64_t base = 0;
  int sm_count = 0;
  int grid_override = 0;
  void* rows = nullptr;
  uint32_t* codes = nullptr;
  CUtensorMap tmap_rows, tmap_q;
  // search workspace (device)
  void* qop = nullptr;
  float* qrec = nullptr;
  uint32_t* qcode = nullptr;
  uint32_t* qmask = nullptr;
  uint64_t* part_keys = nullptr;
  uint32_t* part_cnt = nullptr;
  unsigned long long* stats = nullptr;
  float* gmax = nullptr;
  float* gsample = nullptr;
  unsigned long long* timeline = nullptr;
  int max_parts = 0;
  // profiling (off by default): CUDA events around each kernel of a search
  static constexpr int kProfRing = 256;
  int prof_mode = 0;          // 0 off, 1 events, 2 events + in-kernel timeline
  int prof_calls = 0;         // searches recorded since the last read
  cudaEvent_t* prof_ev = nullptr;  // [kProfRing][4]: before prep, after prep, after scan, after merge
  // staging for the *_host entry points
  float* d_q = nullptr;
  uint32_t* d_code = nullptr;
  uint32_t* d_mask = nullptr;
  float* d_out_s = nullptr;
  int64_t* d_out_ids = nullptr;
  float* h_q = nullptr;
  uint32_t* h_code = nullptr;
  uint32_t* h_mask = nullptr;
  float* h_out_s = nullptr;
  int64_t* h_out_ids = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ws_free = nullptr;  // recorded after the last kernel that uses the workspace
  std::mutex mu;
  int last_grid = 0;
  int last_launches = 0;
  bool f32() const { return dtype == FRS_DTYPE_F32; }
  size_t row_bytes() const { return (size_t)kDim * (f32() ? 4 : 2); }
};

static void free_index(frs_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  cudaFree(ix->rows);
  cudaFree(ix->codes);
  cudaFree(ix->qop);
  cudaFree(ix->qrec);
  cudaFree(ix->qcode);
  cudaFree(ix->qmask);
  cudaFree(ix->part_keys);
  cudaFree(ix->part_cnt);
  cudaFree(ix->stats);
  cudaFree(ix->gmax);
  cudaFree(ix->gsample);
  cudaFree(ix->timeline);
  if (ix->prof_ev) {
    for (int i = 0; i < frs_index::kProfRing * 4; ++i) cudaEventDestroy(ix->prof_ev[i]);
    delete[] ix->prof_ev;
  }
  cudaFree(ix->d_q);
  cudaFree(ix->d_code);
  cudaFree(ix->d_mask);
  cudaFree(ix->d_out_s);
  cudaFree(ix->d_out_ids);
  cudaFreeHost(ix->h_q);
  cudaFreeHost(ix->h_code);
  cudaFreeHost(ix->h_mask);
  cudaFreeHost(ix->h_out_s);
  cudaFreeHost(ix->h_out_ids);
  if (ix->ws_free) cudaEventDestroy(ix->ws_free);
  if (ix->stream) cudaStreamDestroy(ix->stream);
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
  IX_TRY(cudaMalloc(&ix->codes, (size_t)cap_pad * 4));
  IX_TRY(cudaMemset(ix->codes, 0xFF, (size_t)cap_pad * 4));
  IX_TRY(cudaMalloc(&ix->qop, (size_t)kNQ * ix->row_bytes()));
  IX_TRY(cudaMalloc(&ix->qrec, (size_t)kNQ * kDim * 4));
  IX_TRY(cudaMemset(ix->qrec, 0, (size_t)kNQ * kDim * 4));
  IX_TRY(cudaMalloc(&ix->qcode, kNQ * 4));
  IX_TRY(cudaMalloc(&ix->qmask, kNQ * 4));
  IX_TRY(cudaMalloc(&ix->part_keys, (size_t)ix->max_parts * kNQ * kListCap * 8));
  IX_TRY(cudaMalloc(&ix->part_cnt, (size_t)ix->max_parts * kNQ * 4));
  IX_TRY(cudaMalloc(&ix->stats, kStatSlots * 8));
  IX_TRY(cudaMemset(ix->stats, 0, kStatSlots * 8));
  IX_TRY(cudaMalloc(&ix->gmax, (size_t)kNQ * kGmaxPad * 4));
  IX_TRY(cudaMalloc(&ix->gsample, (size_t)kNQ * kSampleBlocks * 4));
  IX_TRY(cudaMalloc(&ix->timeline, (size_t)kGmaxPad * 16 * 8));
  IX_TRY(cudaMemset(ix->timeline, 0, (size_t)kGmaxPad * 16 * 8));
  IX_TRY(cudaMalloc(&ix->d_q, (size_t)kNQ * kDim * 4));
  IX_TRY(cudaMalloc(&ix->d_code, kNQ * 4));
  IX_TRY(cudaMalloc(&ix->d_mask, kNQ * 4));
  IX_TRY(cudaMalloc(&ix->d_out_s, (size_t)kNQ * kMaxK * 4));
  IX_TRY(cudaMalloc(&ix->d_out_ids, (size_t)kNQ * kMaxK * 8));
  IX_TRY(cudaMallocHost(&ix->h_q, (size_t)kNQ * kDim * 4));
  IX_TRY(cudaMallocHost(&ix->h_code, kNQ * 4));
  IX_TRY(cudaMallocHost(&ix->h_mask, kNQ * 4));
  IX_TRY(cudaMallocHost(&ix->h_out_s, (size_t)kNQ * kMaxK * 4));
  IX_TRY(cudaMallocHost(&ix->h_out_ids, (size_t)kNQ * kMaxK * 8));
  IX_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  IX_TRY(cudaEventCreateWithFlags(&ix->ws_free, cudaEventDisableTiming));
#undef IX_TRY
  int rc = make_tmap(&ix->tmap_rows, ix->rows, f32, (uint64_t)cap_pad, kTileM);
  if (!rc) rc = make_tmap(&ix->tmap_q, ix->qop, f32, kNQ, kNQ);
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
extern "C" void* frs_index_rows_ptr(frs_index* idx) { return idx ? idx->rows : nullptr; }
extern "C" uint32_t* frs_index_codes_ptr(frs_index* idx) { return idx ? idx->codes : nullptr; }
extern "C" int frs_index_set_size(frs_index* idx, int64_t n) {
  if (!idx || n < 0 || n > idx->capacity) return set_err(FRS_E_INVALID, "size out of range");
  idx->size = n;
  return FRS_OK;
}

// ---------------------------------------------------------------------------------------------
// write path
// ---------------------------------------------------------------------------------------------
extern "C" int frs_index_set_rows(frs_index* idx, int64_t row0, const float* dev_vecs,
                                  const uint32_t* dev_codes, int64_t n, void* stream) {
  if (!idx || n < 0 || row0 < 0 || (n > 0 && !dev_vecs)) return set_err(FRS_E_INVALID, "bad argument");
  if (row0 + n > idx->size) return set_err(FRS_E_INVALID, "rows [%lld,%lld) beyond size %lld", (long long)row0,
                                           (long long)(row0 + n), (long long)idx->size);
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(launch_store_rows(idx->f32(), dev_vecs, dev_codes, n,
                           (char*)idx->rows + (size_t)row0 * idx->row_bytes(), idx->codes + row0,
                           (cudaStream_t)stream));
  return FRS_OK;
}

extern "C" int frs_index_add(frs_index* idx, const float* dev_vecs, const uint32_t* dev_codes, int64_t n,
                             void* stream) {
  if (!idx || n < 0 || (n > 0 && !dev_vecs)) return set_err(FRS_E_INVALID, "bad argument");
  std::lock_guard<std::mutex> lk(idx->mu);
  if (idx->size + n > idx->capacity)
    return set_err(FRS_E_CAPACITY, "index full: size %lld + %lld > capacity %lld", (long long)idx->size,
                   (long long)n, (long long)idx->capacity);
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(launch_store_rows(idx->f32(), dev_vecs, dev_codes, n,
                           (char*)idx->rows + (size_t)idx->size * idx->row_bytes(),
                           idx->codes + idx->size, (cudaStream_t)stream));
  idx->size += n;
  return FRS_OK;
}

extern "C" int frs_index_add_host(frs_index* idx, const float* host_vecs, const uint32_t* host_codes,
                                  int64_t n) {
  if (!idx || n < 0 || (n > 0 && !host_vecs)) return set_err(FRS_E_INVALID, "bad argument");
  if (idx->size + n > idx->capacity)
    return set_err(FRS_E_CAPACITY, "index full: size %lld + %lld > capacity %lld", (long long)idx->size,
                   (long long)n, (long long)idx->capacity);
  CU_TRY(cudaSetDevice(idx->device));
  const int64_t chunk = 16384;
  float* d_v = nullptr;
  uint32_t* d_c = nullptr;
  CU_TRY(cudaMalloc(&d_v, (size_t)chunk * kDim * 4));
  cudaError_t e = cudaMalloc(&d_c, (size_t)chunk * 4);
  if (e != cudaSuccess) {
    cudaFree(d_v);
    return set_err(FRS_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  int rc = FRS_OK;
  for (int64_t o = 0; o < n && rc == FRS_OK; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    e = cudaMemcpyAsync(d_v, host_vecs + o * kDim, (size_t)m * kDim * 4, cudaMemcpyHostToDevice, idx->stream);
    if (e == cudaSuccess && host_codes)
      e = cudaMemcpyAsync(d_c, host_codes + o, (size_t)m * 4, cudaMemcpyHostToDevice, idx->stream);
    if (e != cudaSuccess) {
      rc = set_err(FRS_E_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
      break;
    }
    rc = frs_index_add(idx, d_v, host_codes ? d_c : nullptr, m, idx->stream);
    if (rc == FRS_OK) {
      e = cudaStreamSynchronize(idx->stream);
      if (e != cudaSuccess) rc = set_err(FRS_E_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    }
  }
  cudaFree(d_v);
  cudaFree(d_c);
  return rc;
}

extern "C" int frs_index_set_codes(frs_index* idx, int64_t row0, const uint32_t* dev_codes, int64_t n,
                                   void* stream) {
  if (!idx || n < 0 || row0 < 0 || row0 + n > idx->size || (n > 0 && !dev_codes))
    return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(cudaMemcpyAsync(idx->codes + row0, dev_codes, (size_t)n * 4, cudaMemcpyDeviceToDevice,
                         (cudaStream_t)stream));
  return FRS_OK;
}

extern "C" int frs_index_read_rows(frs_index* idx, int64_t row0, int64_t n, float* dev_out, void* stream) {
  if (!idx || n < 0 || row0 < 0 || row0 + n > idx->size || (n > 0 && !dev_out))
    return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
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
  CU_TRY(cudaStreamSynchronize(idx->stream));
  CU_TRY(cudaMemcpy(host_rows, (const char*)idx->rows + (size_t)row0 * idx->row_bytes(), (size_t)n * idx->row_bytes(),
                    cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(host_codes, idx->codes + row0, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return FRS_OK;
}

extern "C" int frs_index_import_raw(frs_index* idx, const void* host_rows, const uint32_t* host_codes, int64_t n) {
  if (!idx || n < 0 || (n > 0 && (!host_rows || !host_codes))) return set_err(FRS_E_INVALID, "bad argument");
  std::lock_guard<std::mutex> lk(idx->mu);
  if (idx->size + n > idx->capacity)
    return set_err(FRS_E_CAPACITY, "index full: size %lld + %lld > capacity %lld", (long long)idx->size, (long long)n,
                   (long long)idx->capacity);
  if (n == 0) return FRS_OK;
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(cudaMemcpy((char*)idx->rows + (size_t)idx->size * idx->row_bytes(), host_rows, (size_t)n * idx->row_bytes(),
                    cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(idx->codes + idx->size, host_codes, (size_t)n * 4, cudaMemcpyHostToDevice));
  idx->size += n;
  return FRS_OK;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
static int scan_grid(const frs_index* ix, uint32_t num_tiles) {
  int g = ix->grid_override > 0 ? ix->grid_override : ix->sm_count;
  if (g > ix->max_parts) g = ix->max_parts;
  if ((uint32_t)g > num_tiles) g = (int)num_tiles;
  return g;
}

// prep -> scan -> merge on `st`.  Exactly one of (out_s32) / (out_s64) may be null.
static int search_impl(frs_index* ix, const float* q, const uint32_t* code, const uint32_t* mask, int nq,
                       int k, float* out_s32, double* out_s64, int64_t* out_ids, cudaStream_t st,
                       const uint32_t* tile_ids = nullptr, int64_t n_tile_ids = 0, frs_exchange* push = nullptr) {
  if (!ix) return set_err(FRS_E_INVALID, "idx is null");
  if (nq < 1 || nq > kNQ) return set_err(FRS_E_INVALID, "nq must be in [1,%d] (got %d)", kNQ, nq);
  if (k < 1 || k > kMaxK) return set_err(FRS_E_INVALID, "k must be in [1,%d] (got %d)", kMaxK, k);
  if (!q || !code || !mask || !out_ids || (!out_s32 && !out_s64)) return set_err(FRS_E_INVALID, "null pointer argument");
  CU_TRY(cudaSetDevice(ix->device));
  std::lock_guard<std::mutex> lk(ix->mu);
  const bool f32 = ix->f32();
  const float eps = f32 ? kEpsTF32 : kEpsBF16;
  // the workspace is shared by all calls on this index: order this call after the previous one
  CU_TRY(cudaStreamWaitEvent(st, ix->ws_free, 0));
  int launches = 0;
  cudaEvent_t* pev = nullptr;
  if (ix->prof_mode && ix->prof_ev) pev = ix->prof_ev + (size_t)(ix->prof_calls % frs_index::kProfRing) * 4;
  if (pev) CU_TRY(cudaEventRecord(pev[0], st));
  CU_TRY(launch_prep_queries(f32, q, code, mask, nq, ix->qop, ix->qrec, ix->qcode, ix->qmask, ix->stats, ix->gmax,
                             ix->gsample, ix->rows, ix->codes, (uint32_t)ix->size, st));
  launches++;
  if (pev) CU_TRY(cudaEventRecord(pev[1], st));
  const uint32_t n = (uint32_t)ix->size;
  uint32_t num_tiles = (n + kTileM - 1) / kTileM;
  // restricted scan: only the listed tiles (the caller guarantees that every row matching any query's
  // predicate lies in one of them).  A list too long for the per-CTA table falls back to the full scan.
  if (tile_ids) {
    const int g = scan_grid(ix, (uint32_t)n_tile_ids);
    if (n_tile_ids >= 0 && n_tile_ids <= (int64_t)num_tiles && (g == 0 || (n_tile_ids + g - 1) / g <= kMaxTileSlots))
      num_tiles = (uint32_t)n_tile_ids;
    else
      tile_ids = nullptr;
  }
  const int grid = scan_grid(ix, num_tiles);
  if (grid > 0) {
    ScanParams sp{};
    sp.rows = ix->rows;
    sp.codes = ix->codes;
    sp.n = n;
    sp.num_tiles = num_tiles;
    sp.tile_ids = tile_ids;
    sp.qrec = ix->qrec;
    sp.qcode = ix->qcode;
    sp.qmask = ix->qmask;
    sp.nq = nq;
    sp.k = k;
    sp.eps = eps;
    sp.part_keys = ix->part_keys;
    sp.part_cnt = ix->part_cnt;
    sp.dbg_scores = nullptr;
    sp.gmax = ix->gmax;
    sp.gsample = ix->gsample;
    sp.stats = ix->stats;
    sp.timeline = ix->prof_mode == 2 ? ix->timeline : nullptr;
    if (sp.timeline) CU_TRY(cudaMemsetAsync(ix->timeline, 0, (size_t)kGmaxPad * 16 * 8, st));
    CU_TRY(launch_scan(f32, false, grid, ix->tmap_rows, ix->tmap_q, sp, st));
    launches++;
  }
  if (pev) CU_TRY(cudaEventRecord(pev[2], st));
  MergeParams mp{};
  mp.part_keys = ix->part_keys;
  mp.part_cnt = ix->part_cnt;
  mp.gmax = ix->gmax;
  mp.nparts = grid;
  mp.rows = ix->rows;
  mp.qrec = ix->qrec;
  mp.nq = nq;
  mp.k = k;
  mp.eps = eps;
  mp.base = ix->base;
  mp.out_s64 = out_s64;
  mp.out_s32 = out_s32;
  mp.out_ids = out_ids;
  mp.stats = ix->stats;
  if (push) {  // the exchange step rides in the tail of the merge kernel
    ++push->seq;
    mp.push.peer_gather = push->d_peer_gather;
    mp.push.peer_flags = push->d_peer_flags;
    mp.push.counter = push->counter;
    mp.push.world = push->world;
    mp.push.rank = push->rank;
    mp.push.seq = push->seq;
    mp.push.block_words = push->block_words;
    mp.push.nq_stride = push->nq_max;
  }
  CU_TRY(launch_merge(f32, mp, st));
  launches++;
  if (pev) {
    CU_TRY(cudaEventRecord(pev[3], st));
    ix->prof_calls++;
  }
  CU_TRY(cudaEventRecord(ix->ws_free, st));
  ix->last_grid = grid;
  ix->last_launches = launches;
  return FRS_OK;
}

extern "C" int frs_index_search(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                const uint32_t* dev_q_mask, int nq, int k, float* dev_out_scores,
                                int64_t* dev_out_ids, void* stream) {
  if (!dev_out_scores) return set_err(FRS_E_INVALID, "null pointer argument");
  return search_impl(idx, dev_queries, dev_q_code, dev_q_mask, nq, k, dev_out_scores, nullptr, dev_out_ids,
                     (cudaStream_t)stream);
}

extern "C" int frs_index_search_tiles(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                      const uint32_t* dev_q_mask, int nq, int k, const uint32_t* dev_tile_ids,
                                      int64_t n_tiles, float* dev_out_scores, int64_t* dev_out_ids, void* stream) {
  if (!dev_out_scores || n_tiles < 0 || (n_tiles > 0 && !dev_tile_ids)) return set_err(FRS_E_INVALID, "bad argument");
  static uint32_t* dummy = nullptr;  // an empty list is still a restricted scan (nothing can match)
  if (n_tiles == 0 && !dummy) CU_TRY(cudaMalloc(&dummy, 4));
  return search_impl(idx, dev_queries, dev_q_code, dev_q_mask, nq, k, dev_out_scores, nullptr, dev_out_ids,
                     (cudaStream_t)stream, n_tiles ? dev_tile_ids : dummy, n_tiles);
}

extern "C" int frs_index_search_local(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                      const uint32_t* dev_q_mask, int nq, int k, double* dev_out_scores64,
                                      int64_t* dev_out_ids, void* stream) {
  if (!dev_out_scores64) return set_err(FRS_E_INVALID, "null pointer argument");
  return search_impl(idx, dev_queries, dev_q_code, dev_q_mask, nq, k, nullptr, dev_out_scores64, dev_out_ids,
                     (cudaStream_t)stream);
}

// Local pass of the sharded search WITH the exchange push fused into the merge kernel: the shard's exact top-k
// goes into the exchange's local block and, from the same kernel, into every peer's gather buffer.  Follow with
// frs_exchange_wait_merge.  nq and k must be the ones the exchange was created with.
extern "C" int frs_index_search_push(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                                     const uint32_t* dev_q_mask, int nq, int k, frs_exchange* ex, void* stream) {
  if (!ex) return set_err(FRS_E_INVALID, "null pointer argument");
  if (!ex->connected) return set_err(FRS_E_INVALID, "exchange is not connected");
  if (nq != ex->nq_max || k != ex->k_max)
    return set_err(FRS_E_INVALID, "nq / k (%d / %d) differ from the exchange's (%d / %d)", nq, k, ex->nq_max, ex->k_max);
  if (idx && idx->device != ex->device) return set_err(FRS_E_INVALID, "index and exchange live on different devices");
  const size_t plane = (size_t)nq * k;
  return search_impl(idx, dev_queries, dev_q_code, dev_q_mask, nq, k, nullptr, reinterpret_cast<double*>(ex->local),
                     reinterpret_cast<int64_t*>(ex->local) + plane, (cudaStream_t)stream, nullptr, 0, ex);
}

extern "C" int frs_index_search_host(frs_index* idx, const float* host_queries, const uint32_t* host_q_code,
                                     const uint32_t* host_q_mask, int nq, int k, float* host_out_scores,
                                     int64_t* host_out_ids) {
  if (!idx) return set_err(FRS_E_INVALID, "idx is null");
  if (nq < 1 || nq > kNQ) return set_err(FRS_E_INVALID, "nq must be in [1,%d] (got %d)", kNQ, nq);
  if (k < 1 || k > kMaxK) return set_err(FRS_E_INVALID, "k must be in [1,%d] (got %d)", kMaxK, k);
  if (!host_queries || !host_q_code || !host_q_mask || !host_out_scores || !host_out_ids)
    return set_err(FRS_E_INVALID, "null pointer argument");
  CU_TRY(cudaSetDevice(idx->device));
  // the pinned staging buffers belong to the index: one host-call at a time
  static std::mutex host_mu;
  std::lock_guard<std::mutex> lk(host_mu);
  memcpy(idx->h_q, host_queries, (size_t)nq * kDim * 4);
  memcpy(idx->h_code, host_q_code, (size_t)nq * 4);
  memcpy(idx->h_mask, host_q_mask, (size_t)nq * 4);
  cudaStream_t st = idx->stream;
  CU_TRY(cudaMemcpyAsync(idx->d_q, idx->h_q, (size_t)nq * kDim * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(idx->d_code, idx->h_code, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(idx->d_mask, idx->h_mask, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
  int rc = search_impl(idx, idx->d_q, idx->d_code, idx->d_mask, nq, k, idx->d_out_s, nullptr, idx->d_out_ids, st);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(idx->h_out_s, idx->d_out_s, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(idx->h_out_ids, idx->d_out_ids, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  memcpy(host_out_scores, idx->h_out_s, (size_t)nq * k * 4);
  memcpy(host_out_ids, idx->h_out_ids, (size_t)nq * k * 8);
  return FRS_OK;
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
  CU_TRY(cudaMemcpyAsync(dev_out, idx->qrec, (size_t)kNQ * kDim * 4, cudaMemcpyDeviceToDevice,
                         (cudaStream_t)stream));
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
  CU_TRY(cudaStreamWaitEvent(st, idx->ws_free, 0));
  CU_TRY(cudaMemsetAsync(idx->d_code, 0, kNQ * 4, st));
  CU_TRY(launch_prep_queries(f32, dev_queries, idx->d_code, idx->d_code, nq, idx->qop, idx->qrec, idx->qcode,
                             idx->qmask, idx->stats, idx->gmax, idx->gsample, idx->rows, idx->codes, (uint32_t)idx->size, st));
  const uint32_t n = (uint32_t)idx->size;
  const uint32_t num_tiles = (n + kTileM - 1) / kTileM;
  const int grid = scan_grid(idx, num_tiles);
  if (grid > 0) {
    ScanParams sp{};
    sp.rows = idx->rows;
    sp.codes = idx->codes;
    sp.n = n;
    sp.num_tiles = num_tiles;
    sp.qrec = idx->qrec;
    sp.qcode = idx->qcode;
    sp.qmask = idx->qmask;
    sp.nq = nq;
    sp.k = 1;
    sp.eps = 0.f;
    sp.dbg_scores = dev_out;
    sp.gmax = idx->gmax;
    sp.gsample = idx->gsample;
    sp.stats = idx->stats;
    CU_TRY(launch_scan(f32, true, grid, idx->tmap_rows, idx->tmap_q, sp, st));
  }
  CU_TRY(cudaEventRecord(idx->ws_free, st));
  return FRS_OK;
}

extern "C" int frs_index_last_stats(frs_index* idx, int64_t* host_out6) {
  if (!idx || !host_out6) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long h[kStatSlots];
  CU_TRY(cudaMemcpy(h, idx->stats, sizeof(h), cudaMemcpyDeviceToHost));
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
  if (!idx || mode < 0 || mode > 2) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  if (mode && !idx->prof_ev) {
    idx->prof_ev = new (std::nothrow) cudaEvent_t[frs_index::kProfRing * 4];
    if (!idx->prof_ev) return set_err(FRS_E_INVALID, "out of host memory");
    for (int i = 0; i < frs_index::kProfRing * 4; ++i) CU_TRY(cudaEventCreate(&idx->prof_ev[i]));
  }
  idx->prof_mode = mode;
  idx->prof_calls = 0;
  return FRS_OK;
}

extern "C" int frs_index_read_profile(frs_index* idx, double* host_out4) {
  if (!idx || !host_out4) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  std::lock_guard<std::mutex> lk(idx->mu);
  host_out4[0] = host_out4[1] = host_out4[2] = host_out4[3] = 0.0;
  if (!idx->prof_ev || idx->prof_calls == 0) return FRS_OK;
  const int n = idx->prof_calls < frs_index::kProfRing ? idx->prof_calls : frs_index::kProfRing;
  for (int c = 0; c < n; ++c) {
    cudaEvent_t* ev = idx->prof_ev + (size_t)((idx->prof_calls - 1 - c) % frs_index::kProfRing) * 4;
    CU_TRY(cudaEventSynchronize(ev[3]));
    for (int j = 0; j < 3; ++j) {
      float ms = 0.f;
      CU_TRY(cudaEventElapsedTime(&ms, ev[j], ev[j + 1]));
      host_out4[1 + j] += ms;
    }
  }
  host_out4[0] = n;
  idx->prof_calls = 0;
  return FRS_OK;
}

extern "C" int frs_index_read_timeline(frs_index* idx, uint64_t* host_out, int n_ctas) {
  if (!idx || !host_out || n_ctas < 0 || n_ctas > kGmaxPad) return set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(cudaMemcpy(host_out, idx->timeline, (size_t)n_ctas * 16 * 8, cudaMemcpyDeviceToHost));
  return FRS_OK;
}

// encoder.cu — C ABI of the two BERT encoders (include/frs_b200.h): what get_embedder() /
// get_reranker() of the reference return (main.py:80-90, main2.py:88-103) — SentenceTransformer(
// "BAAI/bge-small-en-v1.5").encode and CrossEncoder("cross-encoder/ms-marco-MiniLM-L-6-v2").predict —
// as one forward pass of hand-written sm_100a kernels (bert.cu) over packed token ids.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/frs_b200.h"
#include "bert.cuh"

namespace frs {
int abi_set_err(int code, const char* fmt, ...);
int abi_make_tmap_bf16(CUtensorMap* m, void* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                       uint32_t box_rows);
}  // namespace frs
using namespace frs;

#define CU_TRY(expr)                                                                             \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return abi_set_err(FRS_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                         __FILE__, __LINE__);                                                    \
  } while (0)

namespace {
enum ProfClass { kPEmbed = 0, kPQkv, kPAttn, kPOut, kPUp, kPDown, kPHead, kPClasses };

struct Layer {
  __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *wqkv32 = nullptr, *wo32 = nullptr, *w132 = nullptr, *w232 = nullptr;  // fp32 mode
  float *bqkv = nullptr, *bo = nullptr, *ln1g = nullptr, *ln1b = nullptr, *b1 = nullptr, *b2 = nullptr,
        *ln2g = nullptr, *ln2b = nullptr;
  CUtensorMap t_wqkv, t_wo, t_w1, t_w2;
};
}  // namespace

struct frs_encoder {
  int device = 0;
  frs_bert_cfg cfg{};
  int sm_count = 0;
  int max_tokens = 0;  // multiple of 128
  int max_seqs = 0;
  int max_qblk = 0;
  // weights
  float *word = nullptr, *pos = nullptr, *type = nullptr, *emb_g = nullptr, *emb_b = nullptr;
  Layer layers[FRS_MAX_LAYERS];
  float *pool_w = nullptr, *pool_b = nullptr, *cls_w = nullptr, *cls_b = nullptr;
  std::vector<void*> owned;  // every device allocation, freed in destroy
  // activations (packed tokens)
  __nv_bfloat16 *x0 = nullptr, *x1 = nullptr, *qk = nullptr, *vt = nullptr, *ctx = nullptr, *h = nullptr;
  float *fx0 = nullptr, *fx1 = nullptr, *fqkv = nullptr, *fctx = nullptr, *fh = nullptr, *ftmp = nullptr;  // fp32 mode
  bool f32() const { return cfg.precision == FRS_PRECISION_F32; }
  int32_t *pos_of_row = nullptr, *src_tok = nullptr, *row_of_tok = nullptr, *cls_slot = nullptr;
  float* cls_f32 = nullptr;  // [max_seqs, 384] fp32 copies of the last layer's [CLS] rows (bf16 path)
  int32_t *d_cu = nullptr, *d_rs = nullptr;  // caller cu_seqlens | internal row starts (multiples of 8)
  QBlock* d_qblk = nullptr;
  CUtensorMap t_x0, t_x1, t_ctx, t_h, t_qk, t_vt;  // box 64 x 128 (vt: 64 x 64): GEMM A operands, attention loads
  CUtensorMap t_k;  // qk with a box of attn_key_block() rows: the attention kernel's K tiles
  CUtensorMap s_x0, s_x1, s_h, s_qk, s_vt;         // epilogue stores: box 32 x 128 SWIZZLE_64B (vt: 64 tokens x 32 dims)
  // staging
  int32_t *h_cu = nullptr, *h_rs = nullptr, *h_ids = nullptr, *h_type = nullptr;
  QBlock* h_qblk = nullptr;
  float* h_out = nullptr;
  int32_t *d_ids = nullptr, *d_type = nullptr;
  float* d_out = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t staged = nullptr;   // the async copies out of h_cu / h_qblk have completed
  cudaEvent_t ws_free = nullptr;  // last kernel using the workspace
  bool staged_pending = false;
  std::mutex mu;
  // profiling
  bool prof = false;
  std::vector<cudaEvent_t> pev;
  std::vector<int> pclass;
  int plaunches = 0;
  int last_tokens = 0;
};

static void free_encoder(frs_encoder* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (void* p : e->owned) cudaFree(p);
  cudaFreeHost(e->h_cu);
  cudaFreeHost(e->h_rs);
  cudaFreeHost(e->h_ids);
  cudaFreeHost(e->h_type);
  cudaFreeHost(e->h_qblk);
  cudaFreeHost(e->h_out);
  for (cudaEvent_t ev : e->pev) cudaEventDestroy(ev);
  if (e->staged) cudaEventDestroy(e->staged);
  if (e->ws_free) cudaEventDestroy(e->ws_free);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

template <typename T>
static cudaError_t dev_alloc(frs_encoder* e, T** out, size_t bytes, bool zero) {
  void* p = nullptr;
  cudaError_t r = cudaMalloc(&p, bytes);
  if (r != cudaSuccess) return r;
  e->owned.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return zero ? cudaMemset(p, 0, bytes) : cudaSuccess;
}

// copy an fp32 tensor to the device (from host or device memory)
static cudaError_t upload_f32(frs_encoder* e, float** out, const float* src, size_t n, bool on_device) {
  cudaError_t r = dev_alloc(e, out, n * 4, false);
  if (r != cudaSuccess) return r;
  return cudaMemcpy(*out, src, n * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
}

// fp32 [rows, cols] tensor(s) -> one bf16 matrix at row offset `row0` of dst
static cudaError_t upload_bf16(frs_encoder* e, __nv_bfloat16* dst, const float* src, size_t n, bool on_device,
                               float* scratch) {
  const float* d = src;
  if (!on_device) {
    cudaError_t r = cudaMemcpy(scratch, src, n * 4, cudaMemcpyHostToDevice);
    if (r != cudaSuccess) return r;
    d = scratch;
  }
  cudaError_t r = launch_f32_to_bf16(d, (int64_t)n, dst, nullptr);
  if (r != cudaSuccess) return r;
  return cudaDeviceSynchronize();
}

extern "C" int frs_encoder_create(int device, const frs_bert_cfg* cfg, const float* const* weights, int n_weights,
                                  int on_device, int max_tokens, frs_encoder** out) {
  if (!out) return abi_set_err(FRS_E_INVALID, "out is null");
  *out = nullptr;
  if (!cfg || !weights) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (cfg->hidden != kHid || cfg->heads != kHeads || cfg->intermediate != kFfn || cfg->max_pos != kMaxSeq ||
      cfg->type_vocab != 2 || cfg->layers < 1 || cfg->layers > FRS_MAX_LAYERS || cfg->vocab_size < 1)
    return abi_set_err(FRS_E_INVALID,
                       "unsupported BERT shape: need hidden 384, heads 12, intermediate 1536, max_pos 512, "
                       "type_vocab 2, 1..%d layers", FRS_MAX_LAYERS);
  if (cfg->precision != FRS_PRECISION_BF16 && cfg->precision != FRS_PRECISION_F32)
    return abi_set_err(FRS_E_INVALID, "precision must be FRS_PRECISION_BF16 or FRS_PRECISION_F32");
  if (n_weights != FRS_BERT_WEIGHTS(cfg->layers, cfg->has_head))
    return abi_set_err(FRS_E_INVALID, "expected %d weight tensors, got %d", FRS_BERT_WEIGHTS(cfg->layers, cfg->has_head),
                       n_weights);
  for (int i = 0; i < n_weights; ++i)
    if (!weights[i]) return abi_set_err(FRS_E_INVALID, "weight %d is null", i);
  if (max_tokens < 128 || max_tokens > (1 << 22)) return abi_set_err(FRS_E_INVALID, "max_tokens out of range");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return abi_set_err(FRS_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                       prop.major, prop.minor);
  frs_encoder* e = new (std::nothrow) frs_encoder();
  if (!e) return abi_set_err(FRS_E_INVALID, "out of host memory");
  e->device = device;
  e->cfg = *cfg;
  e->sm_count = prop.multiProcessorCount;
  e->max_tokens = (max_tokens + kBM - 1) / kBM * kBM;
  e->max_seqs = e->max_tokens / 2 < 64 ? 64 : (e->max_tokens / 2 > 32768 ? 32768 : e->max_tokens / 2);
  e->max_qblk = e->max_tokens / kBM + e->max_seqs;
  const bool dev = on_device != 0;
  const size_t T = (size_t)e->max_tokens;
#define EN_TRY(expr)                                                                        \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      abi_set_err(FRS_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));              \
      free_encoder(e);                                                                      \
      return FRS_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)
#define EN_RC(expr)          \
  do {                       \
    int _rc = (expr);        \
    if (_rc) {               \
      free_encoder(e);       \
      return _rc;            \
    }                        \
  } while (0)
  float* scratch = nullptr;  // host->device staging for the bf16 conversion
  EN_TRY(dev_alloc(e, &scratch, (size_t)kFfn * kHid * 4, false));
  const float* const* w = weights;
  EN_TRY(upload_f32(e, &e->word, w[0], (size_t)cfg->vocab_size * kHid, dev));
  EN_TRY(upload_f32(e, &e->pos, w[1], (size_t)kMaxSeq * kHid, dev));
  EN_TRY(upload_f32(e, &e->type, w[2], (size_t)2 * kHid, dev));
  EN_TRY(upload_f32(e, &e->emb_g, w[3], kHid, dev));
  EN_TRY(upload_f32(e, &e->emb_b, w[4], kHid, dev));
  for (int l = 0; l < cfg->layers; ++l) {
    Layer& L = e->layers[l];
    const float* const* lw = w + 5 + 16 * l;
    const cudaMemcpyKind mk = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    EN_TRY(dev_alloc(e, &L.bqkv, (size_t)kQkvN * 4, false));
    for (int j = 0; j < 3; ++j) EN_TRY(cudaMemcpy(L.bqkv + j * kHid, lw[2 * j + 1], kHid * 4, mk));
    if (cfg->precision == FRS_PRECISION_F32) {
      EN_TRY(dev_alloc(e, &L.wqkv32, (size_t)kQkvN * kHid * 4, false));
      for (int j = 0; j < 3; ++j) EN_TRY(cudaMemcpy(L.wqkv32 + (size_t)j * kHid * kHid, lw[2 * j], (size_t)kHid * kHid * 4, mk));
      EN_TRY(upload_f32(e, &L.wo32, lw[6], (size_t)kHid * kHid, dev));
      EN_TRY(upload_f32(e, &L.w132, lw[10], (size_t)kFfn * kHid, dev));
      EN_TRY(upload_f32(e, &L.w232, lw[12], (size_t)kHid * kFfn, dev));
      EN_TRY(upload_f32(e, &L.bo, lw[7], kHid, dev));
      EN_TRY(upload_f32(e, &L.ln1g, lw[8], kHid, dev));
      EN_TRY(upload_f32(e, &L.ln1b, lw[9], kHid, dev));
      EN_TRY(upload_f32(e, &L.b1, lw[11], kFfn, dev));
      EN_TRY(upload_f32(e, &L.b2, lw[13], kHid, dev));
      EN_TRY(upload_f32(e, &L.ln2g, lw[14], kHid, dev));
      EN_TRY(upload_f32(e, &L.ln2b, lw[15], kHid, dev));
      continue;
    }
    EN_TRY(dev_alloc(e, &L.wqkv, (size_t)kQkvN * kHid * 2, false));
    for (int j = 0; j < 3; ++j)
      EN_TRY(upload_bf16(e, L.wqkv + (size_t)j * kHid * kHid, lw[2 * j], (size_t)kHid * kHid, dev, scratch));
    EN_TRY(dev_alloc(e, &L.wo, (size_t)kHid * kHid * 2, false));
    EN_TRY(upload_bf16(e, L.wo, lw[6], (size_t)kHid * kHid, dev, scratch));
    EN_TRY(upload_f32(e, &L.bo, lw[7], kHid, dev));
    EN_TRY(upload_f32(e, &L.ln1g, lw[8], kHid, dev));
    EN_TRY(upload_f32(e, &L.ln1b, lw[9], kHid, dev));
    EN_TRY(dev_alloc(e, &L.w1, (size_t)kFfn * kHid * 2, false));
    EN_TRY(upload_bf16(e, L.w1, lw[10], (size_t)kFfn * kHid, dev, scratch));
    EN_TRY(upload_f32(e, &L.b1, lw[11], kFfn, dev));
    EN_TRY(dev_alloc(e, &L.w2, (size_t)kHid * kFfn * 2, false));
    EN_TRY(upload_bf16(e, L.w2, lw[12], (size_t)kHid * kFfn, dev, scratch));
    EN_TRY(upload_f32(e, &L.b2, lw[13], kHid, dev));
    EN_TRY(upload_f32(e, &L.ln2g, lw[14], kHid, dev));
    EN_TRY(upload_f32(e, &L.ln2b, lw[15], kHid, dev));
    EN_RC(abi_make_tmap_bf16(&L.t_wqkv, L.wqkv, kQkvN, kHid, 64, 96));  // CTA pair: half a weight slab per CTA
    EN_RC(abi_make_tmap_bf16(&L.t_wo, L.wo, kHid, kHid, 64, 96));
    EN_RC(abi_make_tmap_bf16(&L.t_w1, L.w1, kFfn, kHid, 64, 96));
    EN_RC(abi_make_tmap_bf16(&L.t_w2, L.w2, kHid, kFfn, 64, 96));
  }
  if (cfg->has_head) {
    const float* const* hw = w + 5 + 16 * cfg->layers;
    EN_TRY(upload_f32(e, &e->pool_w, hw[0], (size_t)kHid * kHid, dev));
    EN_TRY(upload_f32(e, &e->pool_b, hw[1], kHid, dev));
    EN_TRY(upload_f32(e, &e->cls_w, hw[2], kHid, dev));
    EN_TRY(upload_f32(e, &e->cls_b, hw[3], 1, dev));
  }
  // activations: zero-initialised so that rows beyond the live tokens are always finite
  if (cfg->precision == FRS_PRECISION_F32) {
    EN_TRY(dev_alloc(e, &e->fx0, T * kHid * 4, true));
    EN_TRY(dev_alloc(e, &e->fx1, T * kHid * 4, true));
    EN_TRY(dev_alloc(e, &e->fqkv, T * kQkvN * 4, true));
    EN_TRY(dev_alloc(e, &e->fctx, T * kHid * 4, true));
    EN_TRY(dev_alloc(e, &e->fh, T * kFfn * 4, true));
    EN_TRY(dev_alloc(e, &e->ftmp, T * kHid * 4, true));
  } else {
    EN_TRY(dev_alloc(e, &e->x0, T * kHid * 2, true));
    EN_TRY(dev_alloc(e, &e->x1, T * kHid * 2, true));
    EN_TRY(dev_alloc(e, &e->qk, T * 2 * kHid * 2, true));
    EN_TRY(dev_alloc(e, &e->vt, T * kHid * 2, true));
    EN_TRY(dev_alloc(e, &e->ctx, T * kHid * 2, true));
    EN_TRY(dev_alloc(e, &e->h, T * kFfn * 2, true));
  }
  EN_TRY(dev_alloc(e, &e->pos_of_row, T * 4, true));
  EN_TRY(dev_alloc(e, &e->src_tok, T * 4, true));
  EN_TRY(dev_alloc(e, &e->row_of_tok, T * 4, true));
  EN_TRY(dev_alloc(e, &e->cls_slot, T * 4, true));
  EN_TRY(dev_alloc(e, &e->cls_f32, (size_t)e->max_seqs * kHid * 4, true));
  EN_TRY(dev_alloc(e, &e->d_cu, ((size_t)e->max_seqs + 1) * 4, true));
  EN_TRY(dev_alloc(e, &e->d_rs, ((size_t)e->max_seqs + 1) * 4, true));
  EN_TRY(dev_alloc(e, &e->d_qblk, (size_t)e->max_qblk * sizeof(QBlock), true));
  EN_TRY(dev_alloc(e, &e->d_ids, T * 4, true));
  EN_TRY(dev_alloc(e, &e->d_type, T * 4, true));
  EN_TRY(dev_alloc(e, &e->d_out, (size_t)e->max_seqs * kHid * 4, true));
  EN_TRY(cudaMallocHost(&e->h_cu, ((size_t)e->max_seqs + 1) * 4));
  EN_TRY(cudaMallocHost(&e->h_rs, ((size_t)e->max_seqs + 1) * 4));
  EN_TRY(cudaMallocHost(&e->h_qblk, (size_t)e->max_qblk * sizeof(QBlock)));
  EN_TRY(cudaMallocHost(&e->h_ids, T * 4));
  EN_TRY(cudaMallocHost(&e->h_type, T * 4));
  EN_TRY(cudaMallocHost(&e->h_out, (size_t)e->max_seqs * kHid * 4));
  EN_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  EN_TRY(cudaEventCreateWithFlags(&e->staged, cudaEventDisableTiming));
  EN_TRY(cudaEventCreateWithFlags(&e->ws_free, cudaEventDisableTiming));
  if (cfg->precision == FRS_PRECISION_F32) {
    *out = e;
    return FRS_OK;
  }
  EN_RC(abi_make_tmap_bf16(&e->t_x0, e->x0, T, kHid, 64, kBM));
  EN_RC(abi_make_tmap_bf16(&e->t_x1, e->x1, T, kHid, 64, kBM));
  EN_RC(abi_make_tmap_bf16(&e->t_ctx, e->ctx, T, kHid, 64, kBM));
  EN_RC(abi_make_tmap_bf16(&e->t_h, e->h, T, kFfn, 64, kBM));
  EN_RC(abi_make_tmap_bf16(&e->t_qk, e->qk, T, 2 * kHid, 64, kBM));
  EN_RC(abi_make_tmap_bf16(&e->t_vt, e->vt, kHid, T, 64, 64));
  EN_RC(abi_make_tmap_bf16(&e->t_k, e->qk, T, 2 * kHid, 64, (uint32_t)attn_key_block()));
  EN_RC(abi_make_tmap_bf16(&e->s_x0, e->x0, T, kHid, 32, kBM));
  EN_RC(abi_make_tmap_bf16(&e->s_x1, e->x1, T, kHid, 32, kBM));
  EN_RC(abi_make_tmap_bf16(&e->s_h, e->h, T, kFfn, 32, kBM));
  EN_RC(abi_make_tmap_bf16(&e->s_qk, e->qk, T, 2 * kHid, 32, kBM));
  EN_RC(abi_make_tmap_bf16(&e->s_vt, e->vt, kHid, T, 64, 32));
#undef EN_TRY
#undef EN_RC
  *out = e;
  return FRS_OK;
}

extern "C" int frs_encoder_destroy(frs_encoder* enc) {
  if (!enc) return FRS_OK;
  cudaSetDevice(enc->device);
  cudaDeviceSynchronize();
  free_encoder(enc);
  return FRS_OK;
}

extern "C" int frs_encoder_max_tokens(const frs_encoder* enc) { return enc ? enc->max_tokens : 0; }

// ---------------------------------------------------------------------------------------------
// forward pass
// ---------------------------------------------------------------------------------------------
static int prof_mark(frs_encoder* e, int cls, cudaStream_t st) {
  static const bool debug_sync = getenv("FRS_DEBUG_SYNC") != nullptr;
  if (debug_sync) {  // fault isolation: name the kernel class that failed
    static const char* names[] = {"embed_ln", "gemm<QKV>", "attention", "gemm<ResLN> out-proj", "gemm<Gelu>",
                                  "gemm<ResLN> ffn-down", "pool/head"};
    cudaError_t err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) {
      const uint32_t* ti = bert_trap_info_host();
      return abi_set_err(FRS_E_CUDA, "kernel %s failed: %s (trap info: wait %u block %u parity %u thread %u)", names[cls],
                         cudaGetErrorString(err), ti ? ti[0] : 0u, ti ? ti[1] : 0u, ti ? ti[2] : 0u, ti ? ti[3] : 0u);
    }
  }
  if (!e->prof) return FRS_OK;
  if (e->plaunches + 1 >= (int)e->pev.size()) return FRS_OK;
  e->pclass[e->plaunches] = cls;
  e->plaunches++;
  CU_TRY(cudaEventRecord(e->pev[e->plaunches], st));
  return FRS_OK;
}

// caller holds e->mu.  Leaves last_hidden_state in x0.
static int forward(frs_encoder* e, const int32_t* d_ids, const int32_t* d_type, const int32_t* host_cu, int n_seqs,
                   cudaStream_t st) {
  if (n_seqs < 1 || n_seqs > e->max_seqs)
    return abi_set_err(FRS_E_INVALID, "n_seqs must be in [1,%d] (got %d)", e->max_seqs, n_seqs);
  if (host_cu[0] != 0) return abi_set_err(FRS_E_INVALID, "cu_seqlens[0] must be 0");
  // the pinned staging buffers are reused: wait for the previous call's copies
  if (e->staged_pending) CU_TRY(cudaEventSynchronize(e->staged));
  // internal layout: every sequence starts on a row that is a multiple of 8 (see bert.cuh)
  int nqb = 0, row = 0;
  for (int s = 0; s < n_seqs; ++s) {
    const int len = host_cu[s + 1] - host_cu[s];
    if (len < 1 || len > kMaxSeq)
      return abi_set_err(FRS_E_INVALID, "sequence %d has %d tokens; must be in [1,%d]", s, len, kMaxSeq);
    e->h_rs[s] = row;
    if (row + len > e->max_tokens)
      return abi_set_err(FRS_E_CAPACITY, "batch does not fit the workspace (max_tokens = %d, sequences are padded to a multiple of 8 rows)", e->max_tokens);
    for (int q0 = 0; q0 < len; q0 += kBM) {
      QBlock& b = e->h_qblk[nqb++];
      b.q_tok0 = row + q0;
      b.seq_tok0 = row;
      b.seq_len = len;
      b.kv_tok0 = row;
    }
    row += (len + 7) & ~7;
  }
  e->h_rs[n_seqs] = row;
  const int M = row < e->max_tokens ? row : e->max_tokens;  // internal rows
  memcpy(e->h_cu, host_cu, ((size_t)n_seqs + 1) * 4);
  CU_TRY(cudaStreamWaitEvent(st, e->ws_free, 0));
  CU_TRY(cudaMemcpyAsync(e->d_cu, e->h_cu, ((size_t)n_seqs + 1) * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(e->d_rs, e->h_rs, ((size_t)n_seqs + 1) * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(e->d_qblk, e->h_qblk, (size_t)nqb * sizeof(QBlock), cudaMemcpyHostToDevice, st));
  CU_TRY(cudaEventRecord(e->staged, st));
  e->staged_pending = true;
  e->plaunches = 0;
  if (getenv("FRS_DEBUG_SYNC")) bert_trap_info_host();  // arm the wait-timeout recorder
  if (e->prof) CU_TRY(cudaEventRecord(e->pev[0], st));
  int rc;
  CU_TRY(launch_row_map(e->d_cu, e->d_rs, n_seqs, e->src_tok, e->pos_of_row, e->row_of_tok, e->cls_slot, st));
  if (e->f32()) {
    // fp32 mode: the same graph with fp32 FFMA kernels (bert_fp32.cu); LayerNorm is a separate kernel here
    CU_TRY(launch_embed_ln_f32(d_ids, d_type, e->src_tok, e->pos_of_row, M, e->cfg.vocab_size, e->word, e->pos, e->type,
                               e->emb_g, e->emb_b, e->cfg.ln_eps, e->fx0, st));
    if ((rc = prof_mark(e, kPEmbed, st))) return rc;
    for (int l = 0; l < e->cfg.layers; ++l) {
      const Layer& L = e->layers[l];
      CU_TRY(launch_sgemm(0, e->fx0, L.wqkv32, L.bqkv, nullptr, e->fqkv, M, kQkvN, kHid, st));
      if ((rc = prof_mark(e, kPQkv, st))) return rc;
      CU_TRY(launch_attention_f32(e->fqkv, e->d_qblk, nqb, e->fctx, st));
      if ((rc = prof_mark(e, kPAttn, st))) return rc;
      CU_TRY(launch_sgemm(2, e->fctx, L.wo32, L.bo, e->fx0, e->ftmp, M, kHid, kHid, st));
      CU_TRY(launch_layernorm_f32(e->ftmp, M, L.ln1g, L.ln1b, e->cfg.ln_eps, e->fx1, st));
      if ((rc = prof_mark(e, kPOut, st))) return rc;
      CU_TRY(launch_sgemm(1, e->fx1, L.w132, L.b1, nullptr, e->fh, M, kFfn, kHid, st));
      if ((rc = prof_mark(e, kPUp, st))) return rc;
      CU_TRY(launch_sgemm(2, e->fh, L.w232, L.b2, e->fx1, e->ftmp, M, kHid, kFfn, st));
      CU_TRY(launch_layernorm_f32(e->ftmp, M, L.ln2g, L.ln2b, e->cfg.ln_eps, e->fx0, st));
      if ((rc = prof_mark(e, kPDown, st))) return rc;
    }
    e->last_tokens = host_cu[n_seqs];
    return FRS_OK;
  }
  CU_TRY(launch_embed_ln(d_ids, d_type, e->src_tok, e->pos_of_row, M, e->cfg.vocab_size, e->word, e->pos, e->type, e->emb_g,
                         e->emb_b, e->cfg.ln_eps, e->x0, st));
  if ((rc = prof_mark(e, kPEmbed, st))) return rc;
  const int mtiles = (M + kBM - 1) / kBM;
  for (int l = 0; l < e->cfg.layers; ++l) {
    const Layer& L = e->layers[l];
    GemmParams g{};
    g.M = M;
    g.num_mtiles = mtiles;
    g.eps = e->cfg.ln_eps;
    // q|k|v projection
    g.N = kQkvN;
    g.K = kHid;
    g.bias = L.bqkv;
    g.qscale = 1.4426950408889634f / sqrtf((float)kHeadDim);
    CU_TRY(launch_gemm(kEpiQKV, e->sm_count, e->t_x0, L.t_wqkv, e->s_qk, e->s_vt, g, st));
    if ((rc = prof_mark(e, kPQkv, st))) return rc;
    AttnParams a{};
    a.qblk = e->d_qblk;
    a.nqb = nqb;
    a.ctx = e->ctx;
    CU_TRY(launch_attention(e->sm_count, e->t_qk, e->t_k, e->t_vt, a, st));
    if ((rc = prof_mark(e, kPAttn, st))) return rc;
    // attention output projection + residual + LayerNorm
    g.N = kHid;
    g.K = kHid;
    g.bias = L.bo;
    g.resid = e->x0;
    g.gamma = L.ln1g;
    g.beta = L.ln1b;
    CU_TRY(launch_gemm(kEpiResLN, e->sm_count, e->t_ctx, L.t_wo, e->s_x1, e->t_x0, g, st));
    if ((rc = prof_mark(e, kPOut, st))) return rc;
    // FFN.  The two GEMMs can run in row chunks (FRS_FFN_CHUNK_TILES=128: FFN-up then FFN-down over the same 16384 rows, so
    // that the chunk of h — 50 MB of the pass's 201 MB — is still in the 126 MB L2 when FFN-down reads it back).  MEASURED
    // on B200 (128 x 512 tokens, ms per 12-layer pass, FFN-up / FFN-down): whole batch 0.84 / 1.07, chunks of 256 tiles
    // 0.98 / 1.19, 128 tiles 1.29 / 1.29, 64 tiles 1.86 / 2.43 — the ramp and tail of every extra launch of these persistent
    // kernels cost more than the L2 hits give back.  Off by default (0 = one launch per GEMM); chunks are an even number of
    // row tiles (CTA pairs own two).
    static const int chunk_tiles_env = getenv("FRS_FFN_CHUNK_TILES") ? atoi(getenv("FRS_FFN_CHUNK_TILES")) : 0;
    const int chunk_tiles = chunk_tiles_env > 0 ? (chunk_tiles_env + 1) / 2 * 2 : mtiles;
    for (int t0 = 0; t0 < mtiles; t0 += chunk_tiles) {
      const int nt = mtiles - t0 < chunk_tiles ? mtiles - t0 : chunk_tiles;
      g.mtile0 = t0;
      g.num_mtiles = nt;
      g.N = kFfn;
      g.K = kHid;
      g.bias = L.b1;
      g.resid = nullptr;
      g.cls_slot = nullptr;
      g.cls_out = nullptr;
      CU_TRY(launch_gemm(kEpiGelu, e->sm_count, e->t_x1, L.t_w1, e->s_h, e->s_h, g, st));
      if ((rc = prof_mark(e, kPUp, st))) return rc;
      g.N = kHid;
      g.K = kFfn;
      g.bias = L.b2;
      g.resid = e->x1;
      g.gamma = L.ln2g;
      g.beta = L.ln2b;
      if (l == e->cfg.layers - 1) {  // the last layer's [CLS] rows also leave in fp32 (head / CLS pooling operand)
        g.cls_slot = e->cls_slot;
        g.cls_out = e->cls_f32;
      }
      CU_TRY(launch_gemm(kEpiResLN, e->sm_count, e->t_h, L.t_w2, e->s_x0, e->t_x1, g, st));
      if ((rc = prof_mark(e, kPDown, st))) return rc;
    }
    g.mtile0 = 0;
    g.num_mtiles = mtiles;
    g.cls_slot = nullptr;
    g.cls_out = nullptr;
  }
  e->last_tokens = host_cu[n_seqs];
  return FRS_OK;
}

extern "C" int frs_encoder_embed(frs_encoder* enc, const int32_t* dev_ids, const int32_t* host_cu_seqlens, int n_seqs,
                                 int pool_mode, float* dev_out, void* stream) {
  if (!enc || !dev_ids || !host_cu_seqlens || !dev_out) return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (pool_mode != FRS_POOL_CLS && pool_mode != FRS_POOL_MEAN) return abi_set_err(FRS_E_INVALID, "bad pool_mode");
  CU_TRY(cudaSetDevice(enc->device));
  std::lock_guard<std::mutex> lk(enc->mu);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = forward(enc, dev_ids, nullptr, host_cu_seqlens, n_seqs, st);
  if (rc) return rc;
  if (enc->f32())
    CU_TRY(launch_pool_normalize_f32(enc->fx0, enc->d_cu, enc->d_rs, n_seqs, pool_mode, dev_out, st));
  else if (pool_mode == FRS_POOL_CLS)  // fp32 [CLS] rows kept by the last LayerNorm epilogue, dense [n_seqs, 384]
    CU_TRY(launch_pool_normalize_f32(enc->cls_f32, enc->d_cu, nullptr, n_seqs, pool_mode, dev_out, st));
  else
    CU_TRY(launch_pool_normalize(enc->x0, enc->d_cu, enc->d_rs, n_seqs, pool_mode, dev_out, st));
  if ((rc = prof_mark(enc, kPHead, st))) return rc;
  CU_TRY(cudaEventRecord(enc->ws_free, st));
  return FRS_OK;
}

extern "C" int frs_encoder_score_pairs(frs_encoder* enc, const int32_t* dev_ids, const int32_t* dev_type_ids,
                                       const int32_t* host_cu_seqlens, int n_seqs, float* dev_logits, void* stream) {
  if (!enc || !dev_ids || !dev_type_ids || !host_cu_seqlens || !dev_logits)
    return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (!enc->cfg.has_head) return abi_set_err(FRS_E_STATE, "this encoder has no classifier head");
  CU_TRY(cudaSetDevice(enc->device));
  std::lock_guard<std::mutex> lk(enc->mu);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = forward(enc, dev_ids, dev_type_ids, host_cu_seqlens, n_seqs, st);
  if (rc) return rc;
  if (enc->f32())
    CU_TRY(launch_ce_head_f32(enc->fx0, enc->d_rs, n_seqs, enc->pool_w, enc->pool_b, enc->cls_w, enc->cls_b, dev_logits, st));
  else  // pooler + classifier in fp32 from the fp32 [CLS] rows (dense [n_seqs, 384]: no row table)
    CU_TRY(launch_ce_head_f32(enc->cls_f32, nullptr, n_seqs, enc->pool_w, enc->pool_b, enc->cls_w, enc->cls_b, dev_logits, st));
  if ((rc = prof_mark(enc, kPHead, st))) return rc;
  CU_TRY(cudaEventRecord(enc->ws_free, st));
  return FRS_OK;
}

// host buffers of any size: greedy passes of <= max_tokens packed tokens, split on sequence boundaries
static int run_host(frs_encoder* enc, const int32_t* host_ids, const int32_t* host_type, const int32_t* host_cu,
                    int n_seqs, int pool_mode, bool pairs, float* host_out) {
  if (!enc || !host_ids || !host_cu || !host_out || (pairs && !host_type))
    return abi_set_err(FRS_E_INVALID, "null pointer argument");
  if (n_seqs < 0) return abi_set_err(FRS_E_INVALID, "n_seqs < 0");
  if (n_seqs == 0) return FRS_OK;
  CU_TRY(cudaSetDevice(enc->device));
  static std::mutex host_mu;  // the staging buffers belong to the encoder: one host call at a time
  std::lock_guard<std::mutex> hl(host_mu);
  cudaStream_t st = enc->stream;
  const int width = pairs ? 1 : kHid;
  std::vector<int32_t> cu;
  int s0 = 0;
  while (s0 < n_seqs) {
    int s1 = s0, rows = 0;
    const int base = host_cu[s0];
    while (s1 < n_seqs && s1 - s0 < enc->max_seqs) {
      const int len = host_cu[s1 + 1] - host_cu[s1];
      if (rows + len > enc->max_tokens) break;
      rows += (len + 7) & ~7;  // internal layout: sequences start on multiples of 8 rows
      ++s1;
    }
    if (s1 == s0) return abi_set_err(FRS_E_INVALID, "sequence %d is longer than max_tokens", s0);
    const int ntok = host_cu[s1] - base;
    cu.resize((size_t)(s1 - s0) + 1);
    for (int s = s0; s <= s1; ++s) cu[(size_t)(s - s0)] = host_cu[s] - base;
    memcpy(enc->h_ids, host_ids + base, (size_t)ntok * 4);
    CU_TRY(cudaMemcpyAsync(enc->d_ids, enc->h_ids, (size_t)ntok * 4, cudaMemcpyHostToDevice, st));
    if (pairs) {
      memcpy(enc->h_type, host_type + base, (size_t)ntok * 4);
      CU_TRY(cudaMemcpyAsync(enc->d_type, enc->h_type, (size_t)ntok * 4, cudaMemcpyHostToDevice, st));
    }
    int rc = pairs ? frs_encoder_score_pairs(enc, enc->d_ids, enc->d_type, cu.data(), s1 - s0, enc->d_out, st)
                   : frs_encoder_embed(enc, enc->d_ids, cu.data(), s1 - s0, pool_mode, enc->d_out, st);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(enc->h_out, enc->d_out, (size_t)(s1 - s0) * width * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    memcpy(host_out + (size_t)s0 * width, enc->h_out, (size_t)(s1 - s0) * width * 4);
    s0 = s1;
  }
  return FRS_OK;
}

extern "C" int frs_encoder_embed_host(frs_encoder* enc, const int32_t* host_ids, const int32_t* host_cu_seqlens,
                                      int n_seqs, int pool_mode, float* host_out) {
  if (pool_mode != FRS_POOL_CLS && pool_mode != FRS_POOL_MEAN) return abi_set_err(FRS_E_INVALID, "bad pool_mode");
  return run_host(enc, host_ids, nullptr, host_cu_seqlens, n_seqs, pool_mode, false, host_out);
}

extern "C" int frs_encoder_score_pairs_host(frs_encoder* enc, const int32_t* host_ids, const int32_t* host_type_ids,
                                            const int32_t* host_cu_seqlens, int n_seqs, float* host_logits) {
  if (enc && !enc->cfg.has_head) return abi_set_err(FRS_E_STATE, "this encoder has no classifier head");
  return run_host(enc, host_ids, host_type_ids, host_cu_seqlens, n_seqs, 0, true, host_logits);
}

extern "C" int frs_encoder_last_hidden(frs_encoder* enc, float* dev_out, int n_tokens, void* stream) {
  if (!enc || !dev_out || n_tokens < 0 || n_tokens > enc->max_tokens) return abi_set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(enc->device));
  std::lock_guard<std::mutex> lk(enc->mu);
  cudaStream_t st = (cudaStream_t)stream;
  CU_TRY(cudaStreamWaitEvent(st, enc->ws_free, 0));
  if (n_tokens > enc->last_tokens) return abi_set_err(FRS_E_INVALID, "the last pass had %d tokens", enc->last_tokens);
  if (enc->f32())
    CU_TRY(launch_gather_rows_f32f32(enc->fx0, enc->row_of_tok, n_tokens, dev_out, st));
  else
    CU_TRY(launch_gather_rows_f32(enc->x0, enc->row_of_tok, n_tokens, dev_out, st));
  CU_TRY(cudaEventRecord(enc->ws_free, st));
  return FRS_OK;
}

extern "C" int frs_encoder_debug_read(frs_encoder* enc, int which, float* dev_out, int64_t n_elems, void* stream) {
  if (!enc || !dev_out || n_elems < 0) return abi_set_err(FRS_E_INVALID, "bad argument");
  const int64_t T = enc->max_tokens;
  const __nv_bfloat16* src[6] = {enc->x0, enc->x1, enc->qk, enc->vt, enc->ctx, enc->h};
  const int64_t cap[6] = {T * kHid, T * kHid, T * 2 * kHid, T * kHid, T * kHid, T * kFfn};
  if (which < 0 || which > 5 || n_elems > cap[which]) return abi_set_err(FRS_E_INVALID, "bad buffer / size");
  if (enc->f32()) return abi_set_err(FRS_E_STATE, "debug_read reads the bf16 workspace; this encoder runs in fp32 mode");
  CU_TRY(cudaSetDevice(enc->device));
  std::lock_guard<std::mutex> lk(enc->mu);
  cudaStream_t st = (cudaStream_t)stream;
  CU_TRY(cudaStreamWaitEvent(st, enc->ws_free, 0));
  CU_TRY(launch_bf16_to_f32(src[which], n_elems, dev_out, st));
  CU_TRY(cudaEventRecord(enc->ws_free, st));
  return FRS_OK;
}

extern "C" int frs_encoder_set_profiling(frs_encoder* enc, int on) {
  if (!enc) return abi_set_err(FRS_E_INVALID, "enc is null");
  CU_TRY(cudaSetDevice(enc->device));
  std::lock_guard<std::mutex> lk(enc->mu);
  if (on && enc->pev.empty()) {
    const int n = (3 + 2 * 64) * FRS_MAX_LAYERS + 8;  // QKV, attention, out-proj + (FFN-up, FFN-down) per row chunk
    enc->pev.resize(n);
    enc->pclass.assign(n, 0);
    for (int i = 0; i < n; ++i) CU_TRY(cudaEventCreate(&enc->pev[i]));
  }
  enc->prof = on != 0;
  enc->plaunches = 0;
  return FRS_OK;
}

extern "C" int frs_encoder_read_profile(frs_encoder* enc, double* host_out8) {
  if (!enc || !host_out8) return abi_set_err(FRS_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(enc->device));
  std::lock_guard<std::mutex> lk(enc->mu);
  for (int i = 0; i < 8; ++i) host_out8[i] = 0.0;
  if (!enc->prof || enc->plaunches == 0) return FRS_OK;
  CU_TRY(cudaEventSynchronize(enc->pev[enc->plaunches]));
  for (int i = 0; i < enc->plaunches; ++i) {
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, enc->pev[i], enc->pev[i + 1]));
    host_out8[enc->pclass[i]] += ms;
  }
  // embeddings = 2 kernels (positions + gather/LN); every other mark is one kernel
  host_out8[7] = enc->plaunches + 1;
  return FRS_OK;
}

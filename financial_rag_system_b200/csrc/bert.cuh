// bert.cuh — launch interface of the BERT encoder kernels (kernels in bert.cu).
//
// These replace the torch/ATen ops that SentenceTransformer.encode (reference main.py:148,213,
// main2.py:171) and CrossEncoder.predict (main.py:245, main2.py:166) execute through
// transformers' BertModel (modeling_bert.py): embeddings+LayerNorm (:102-112), self-attention
// (:115-207), output projections with residual+LayerNorm (:294-298, :352-356), erf-GELU FFN
// (:339-342), CLS/mean pooling + L2 normalisation, and the pooler+classifier head (:462-468,
// :1111-1124).  Tokens of a batch are PACKED (no padding): sequence s owns rows
// [cu_seqlens[s], cu_seqlens[s+1]) of every activation matrix.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace frs {

constexpr int kHid = 384;        // hidden size of both checkpoints
constexpr int kHeads = 12;
constexpr int kHeadDim = 32;
constexpr int kHeadPairs = kHeads / 2;  // attention works on pairs of heads: 64 dims = one 128-byte row
constexpr int kFfn = 1536;
constexpr int kQkvN = 3 * kHid;  // fused q|k|v projection
constexpr int kBM = 128;         // token rows per GEMM / attention tile == MMA M == TMEM lanes
constexpr int kMaxSeq = 512;     // max_position_embeddings

enum GemmEpi {
  kEpiQKV = 0,    // + bias; q scaled by log2(e)/sqrt(32); q|k -> qk[M,768] bf16, v -> vt[384,M] bf16 (transposed)
  kEpiGelu = 1,   // gelu(acc + bias) -> bf16 [M, N]
  kEpiResLN = 2,  // LayerNorm(acc + bias + residual) * gamma + beta -> bf16 [M, 384]
};

struct GemmParams {
  int M;           // live token rows
  int num_mtiles;  // row tiles this launch covers (ceil(M / 128) for the whole batch)
  int mtile0;      // first row tile of this launch (even): a launch may cover a slice of the rows (FFN row chunks)
  int N, K;
  const float* bias;   // [N]
  float qscale;
  const __nv_bfloat16* resid;  // ResLN only: [rows, 384]
  const float* gamma;
  const float* beta;
  float eps;
  const int32_t* cls_slot;  // ResLN, last layer only: [rows] sequence index of a [CLS] row, else -1 (null = off)
  float* cls_out;           //   fp32 copies of the [CLS] rows' outputs, [n_seqs, 384]
  int debug;  // FRS_GEMM_DEBUG (measurement aid): 1 = epilogue only drains TMEM (no math, no stores)
  long long* trace;  // -DFRS_GEMM_TRACE builds: event timeline of CTA 0 (see launch_gemm_t), else null
};

// one entry per (sequence, 128-query block)
struct QBlock {
  int32_t q_tok0;    // first token row of this query block
  int32_t seq_tok0;  // first token row of the sequence
  int32_t seq_len;
  int32_t kv_tok0;   // seq_tok0 rounded down to a multiple of 8: key blocks start here (a TMA box must
                     // start on a 16-byte boundary and tokens are the contiguous dimension of vt)
};

struct AttnParams {
  const QBlock* qblk;
  int nqb;             // work items = nqb * kHeadPairs
  __nv_bfloat16* ctx;  // [rows, 384]
  long long* timing;   // -DFRS_ATTN_TIMING builds: [8 warps][8 phases] clock sums of CTA 0, else null
};

size_t gemm_smem_bytes(int epi);
size_t attn_smem_bytes();

// Internal row layout: sequence s owns rows [row_start[s], row_start[s] + len_s) and row_start[s] is a
// multiple of 8 (so that every TMA box of the transposed V starts on a 16-byte boundary, and a
// sequence's arithmetic does not depend on where it sits in the batch).  cu_seqlens is the caller's
// packed layout.  One block per sequence fills, for every internal row of the sequence's slot,
// src_tok[row] (caller token index, -1 for the <= 7 alignment rows), pos_of_row[row] (position in the
// sequence), cls_slot[row] (the sequence index on its first row, -1 elsewhere) and, for every caller token,
// row_of_tok[token].
cudaError_t launch_row_map(const int32_t* cu_seqlens, const int32_t* row_start, int n_seqs, int32_t* src_tok,
                           int32_t* pos_of_row, int32_t* row_of_tok, int32_t* cls_slot, cudaStream_t st);
// x[row] = LayerNorm(word[id] + pos[p] + type[tt]) -> bf16 for the M internal rows (alignment rows = 0);
// type_ids may be null (all 0)
cudaError_t launch_embed_ln(const int32_t* ids, const int32_t* type_ids, const int32_t* src_tok,
                            const int32_t* pos_of_row, int M, int vocab, const float* word, const float* pos,
                            const float* type, const float* gamma, const float* beta, float eps, __nv_bfloat16* x,
                            cudaStream_t st);
// C = A[M,K] * W[N,K]^T (+ epilogue).  tmap_a: activations, box 64 x 128, SWIZZLE_128B;
// tmap_b: weights, box 64 x 192, SWIZZLE_128B.  Outputs leave through TMA stores:
//   QKV    tmap_out  = qk [rows, 768], box 32 x 128 SWIZZLE_64B;  tmap_out2 = vt [384, rows], box 64 x 32 SWIZZLE_128B
//   GELU   tmap_out  = h [rows, 1536], box 32 x 128 SWIZZLE_64B   (tmap_out2 unused)
//   ResLN  tmap_out  = x [rows, 384],  box 32 x 128 SWIZZLE_64B;  tmap_out2 = the RESIDUAL as an A operand (box 64 x 128 SWIZZLE_128B)
cudaError_t launch_gemm(int epi, int sm_count, const CUtensorMap& tmap_a, const CUtensorMap& tmap_b,
                        const CUtensorMap& tmap_out, const CUtensorMap& tmap_out2, const GemmParams& p,
                        cudaStream_t st);
// tmap_qk: qk [rows, 768], box 64 x 128; tmap_vt: vt [384, rows], box 64 x 64; both SWIZZLE_128B
int attn_key_block();  // keys per block of the compiled attention kernel = rows of the K tensor map's box
cudaError_t launch_attention(int sm_count, const CUtensorMap& tmap_qk, const CUtensorMap& tmap_k, const CUtensorMap& tmap_vt,
                             const AttnParams& p, cudaStream_t st);
// pool_mode 0 = CLS row, 1 = mean over the sequence; then x / max(||x||, 1e-12)
cudaError_t launch_pool_normalize(const __nv_bfloat16* x, const int32_t* cu_seqlens, const int32_t* row_start,
                                  int n_seqs, int pool_mode, float* out, cudaStream_t st);
// logits[s] = wc . tanh(Wp * x[cls_s] + bp) + bc

// out[t] = fp32(x[row_of_tok[t]]) for the caller's packed tokens
cudaError_t launch_gather_rows_f32(const __nv_bfloat16* x, const int32_t* row_of_tok, int n_tokens, float* out,
                                   cudaStream_t st);
// ---- fp32 mode (bert_fp32.cu): fp32 weights / activations / FFMA arithmetic, for the 1e-5 parity bound ----
// epi: 0 = +bias, 1 = +bias then erf GELU, 2 = +bias +residual.  C = A[M,K] . W[N,K]^T; N % 128 == 0, K % 16 == 0
cudaError_t launch_sgemm(int epi, const float* A, const float* W, const float* bias, const float* resid, float* C,
                         int M, int N, int K, cudaStream_t st);
cudaError_t launch_embed_ln_f32(const int32_t* ids, const int32_t* type_ids, const int32_t* src_tok,
                                const int32_t* pos_of_row, int M, int vocab, const float* word, const float* pos,
                                const float* type, const float* gamma, const float* beta, float eps, float* x,
                                cudaStream_t st);
cudaError_t launch_layernorm_f32(const float* x, int M, const float* gamma, const float* beta, float eps, float* y,
                                 cudaStream_t st);
// qkv: [rows, 1152] = q | k | v (unscaled); one block per (query block, head)
cudaError_t launch_attention_f32(const float* qkv, const QBlock* qblk, int nqb, float* ctx, cudaStream_t st);
cudaError_t launch_pool_normalize_f32(const float* x, const int32_t* cu_seqlens, const int32_t* row_start, int n_seqs,
                                      int pool_mode, float* out, cudaStream_t st);
cudaError_t launch_ce_head_f32(const float* x, const int32_t* row_start, int n_seqs, const float* wp, const float* bp,
                               const float* wc, const float* bc, float* logits, cudaStream_t st);
cudaError_t launch_gather_rows_f32f32(const float* x, const int32_t* row_of_tok, int n_tokens, float* out,
                                      cudaStream_t st);

// mapped host words {wait code, blockIdx, parity, threadIdx} written by a barrier wait that timed out
uint32_t* bert_trap_info_host();
cudaError_t launch_bf16_to_f32(const __nv_bfloat16* src, int64_t n, float* dst, cudaStream_t st);
cudaError_t launch_f32_to_bf16(const float* src, int64_t n, __nv_bfloat16* dst, cudaStream_t st);

}  // namespace frs

// bert_fp32.cu — the fp32 mode of the encoders: the same forward pass as bert.cu with fp32 weights,
// fp32 activations and fp32 FFMA arithmetic throughout (no tensor cores: TF32 / bf16 operands cannot
// meet the 1e-5 tolerance north_star sets for the fp32 mode).  It exists for parity, not for speed —
// it is ~20x slower than the bf16 tensor-core path — and is the on-GPU witness that the bf16 path's
// deviation from the reference is rounding, not structure.
//
//   embed_ln_f32_kernel     BertEmbeddings.forward          modeling_bert.py:102-112
//   sgemm_kernel<EPI>       every Linear (+bias, +erf GELU, +residual)   :179-181, :294-298, :339-342, :352-356
//   layernorm_f32_kernel    LayerNorm of BertSelfOutput / BertOutput     :296-298, :354-356
//   attention_f32_kernel    softmax(QK^T/sqrt(32) + mask) V              :115-140
//   pool / head             as in bert.cu, on fp32 hidden states
#include <math.h>

#include "bert.cuh"
#include "common.cuh"

namespace frs {

__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// embeddings + LayerNorm, one warp per internal row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_ln_f32_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ type_ids,
                    const int32_t* __restrict__ src_tok, const int32_t* __restrict__ pos_of_row, int M, int vocab,
                    const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ x) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int tok = src_tok[row];
  float* out = x + (size_t)row * kHid;
  if (tok < 0) {
    for (int i = lane; i < kHid; i += 32) out[i] = 0.f;
    return;
  }
  int id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const int tt = type_ids ? (type_ids[tok] != 0) : 0;
  int ps = pos_of_row[row];
  ps = ps < 0 ? 0 : (ps >= kMaxSeq ? kMaxSeq - 1 : ps);
  float v[12], sum = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const int c = lane + 32 * i;
    v[i] = word[(size_t)id * kHid + c] + type[(size_t)tt * kHid + c] + pos[(size_t)ps * kHid + c];
    sum += v[i];
  }
  const float mean = warp_sum32(sum) * (1.0f / kHid);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) sq += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum32(sq) * (1.0f / kHid) + eps);
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const int c = lane + 32 * i;
    out[c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

// ---------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] . W[N,K]^T + bias (+ epilogue), fp32 FFMA, 128 x 128 x 16 tiles, 8 x 8 per thread
// ---------------------------------------------------------------------------------------------
enum SgemmEpi { kSBias = 0, kSGelu = 1, kSResidual = 2 };

template <int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
             const float* __restrict__ resid, float* __restrict__ C, int M, int N, int K) {
  __shared__ float As[2][16][128 + 4];
  __shared__ float Ws[2][16][128 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8 x 8 outputs each
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader: rows lrow and lrow + 64, k offset lk..lk+3
  // Two-level accumulation: fp32 FFMA over one 16-wide K tile, then the tile's partial sum is added in fp64.  The rounding
  // error of a K = 1536 dot product then looks like that of a 16-term one (this is the PARITY mode: its job is to sit inside
  // 1e-5 of an fp32 reference whose own summation order is different).
  float acc[8][8];
  double dacc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[i][j] = 0.f;
      dacc[i][j] = 0.0;
    }
  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = m0 + lrow + 64 * h;
      ra[h] = r < M ? *reinterpret_cast<const float4*>(A + (size_t)r * K + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lrow + 64 * h) * K + k0 + lk);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + 64 * h;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / 16;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dacc[i][j] += (double)acc[i][j];
        acc[i][j] = 0.f;
      }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + ty * 8 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + tx * 8 + j;
      float v = (float)(dacc[i][j] + (double)bias[c]);
      if (EPI == kSGelu) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
      if (EPI == kSResidual) v += resid[(size_t)r * N + c];
      C[(size_t)r * N + c] = v;
    }
  }
}

// y = LayerNorm(x) * gamma + beta, one warp per row of 384 (in place allowed)
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, int M, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, float* __restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float v[12], sum = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    v[i] = x[(size_t)row * kHid + lane + 32 * i];
    sum += v[i];
  }
  const float mean = warp_sum32(sum) * (1.0f / kHid);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) sq += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum32(sq) * (1.0f / kHid) + eps);
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const int c = lane + 32 * i;
    y[(size_t)row * kHid + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

// ---------------------------------------------------------------------------------------------
// attention: one block per (128-query block, head); K and V of the sequence for that head in shared
// memory; a thread owns a query row and runs the online softmax over all keys in fp32
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ qkv, const QBlock* __restrict__ qblk, float* __restrict__ ctx) {
  extern __shared__ float sm[];
  const QBlock qb = qblk[blockIdx.x];
  const int head = blockIdx.y;
  const int S = qb.seq_len;
  float* Ks = sm;                       // [S][32]
  float* Vs = sm + (size_t)kMaxSeq * kHeadDim;  // [S][32]
  for (int i = threadIdx.x; i < S * kHeadDim; i += blockDim.x) {
    const int t = i >> 5, d = i & 31;
    const float* base = qkv + (size_t)(qb.seq_tok0 + t) * kQkvN + head * kHeadDim + d;
    Ks[i] = base[kHid];
    Vs[i] = base[2 * kHid];
  }
  __syncthreads();
  const int qi = qb.q_tok0 - qb.seq_tok0 + (int)threadIdx.x;
  if (qi >= S) return;
  const size_t row = (size_t)qb.q_tok0 + threadIdx.x;
  float q[kHeadDim], acc[kHeadDim];
  const float scale = 0.17677669529663687f;  // 1 / sqrt(32)
#pragma unroll
  for (int d = 0; d < kHeadDim; ++d) {
    q[d] = qkv[row * kQkvN + head * kHeadDim + d] * scale;
    acc[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < S; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)j * kHeadDim);
    float s = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < kHeadDim / 4; ++d4) {
      const float4 kv = kr[d4];
      s = fmaf(q[4 * d4], kv.x, s);
      s = fmaf(q[4 * d4 + 1], kv.y, s);
      s = fmaf(q[4 * d4 + 2], kv.z, s);
      s = fmaf(q[4 * d4 + 3], kv.w, s);
    }
    if (s > m) {
      const float a = expf(m - s);
      l *= a;
#pragma unroll
      for (int d = 0; d < kHeadDim; ++d) acc[d] *= a;
      m = s;
    }
    const float p = expf(s - m);
    l += p;
    const float4* vr = reinterpret_cast<const float4*>(Vs + (size_t)j * kHeadDim);
#pragma unroll
    for (int d4 = 0; d4 < kHeadDim / 4; ++d4) {
      const float4 vv = vr[d4];
      acc[4 * d4] = fmaf(p, vv.x, acc[4 * d4]);
      acc[4 * d4 + 1] = fmaf(p, vv.y, acc[4 * d4 + 1]);
      acc[4 * d4 + 2] = fmaf(p, vv.z, acc[4 * d4 + 2]);
      acc[4 * d4 + 3] = fmaf(p, vv.w, acc[4 * d4 + 3]);
    }
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int d = 0; d < kHeadDim; ++d) ctx[row * kHid + head * kHeadDim + d] = acc[d] * inv;
}

// ---------------------------------------------------------------------------------------------
// pooling + L2 normalise, cross-encoder head (fp32 hidden states)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pool_normalize_f32_kernel(const float* __restrict__ x, const int32_t* __restrict__ cu, const int32_t* __restrict__ row_start,
                          int n_seqs, int pool_mode, float* __restrict__ out) {
  __shared__ float red[4];
  const int s = blockIdx.x;
  if (s >= n_seqs) return;
  // row_start == null: x holds one row per sequence (the fp32 [CLS] rows of the bf16 path)
  const int t0 = row_start ? row_start[s] : s, t1 = t0 + (cu[s + 1] - cu[s]);
  float v[3] = {0.f, 0.f, 0.f};
  if (pool_mode == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = x[(size_t)t0 * kHid + threadIdx.x + 128 * i];
  } else {
    for (int t = t0; t < t1; ++t)
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] += x[(size_t)t * kHid + threadIdx.x + 128 * i];
    const float inv = 1.0f / (float)max(t1 - t0, 1);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] *= inv;
  }
  const float sq = warp_sum32(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  const float inv = 1.0f / fmaxf(sqrtf((red[0] + red[1]) + (red[2] + red[3])), 1e-12f);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[(size_t)s * kHid + threadIdx.x + 128 * i] = v[i] * inv;
}

__global__ void __launch_bounds__(384)
ce_head_f32_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_start, int n_seqs,
                   const float* __restrict__ wp, const float* __restrict__ bp, const float* __restrict__ wc,
                   const float* __restrict__ bc, float* __restrict__ logits) {
  __shared__ float xs[kHid];
  __shared__ float pooled[kHid];
  __shared__ float red[12];
  const int s = blockIdx.x;
  if (s >= n_seqs) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  xs[threadIdx.x] = x[(size_t)(row_start ? row_start[s] : s) * kHid + threadIdx.x];
  __syncthreads();
  for (int o = warp * 32; o < warp * 32 + 32; ++o) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) a = fmaf(wp[(size_t)o * kHid + lane + 32 * i], xs[lane + 32 * i], a);
    a = warp_sum32(a);
    if (lane == 0) pooled[o] = tanhf(a + bp[o]);
  }
  __syncthreads();
  const float a = warp_sum32(pooled[threadIdx.x] * wc[threadIdx.x]);
  if (lane == 0) red[warp] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 12; ++i) t += red[i];
    logits[s] = t + bc[0];
  }
}

__global__ void gather_rows_f32f32_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_of_tok, int n_tokens,
                                          float* __restrict__ out) {
  const int t = blockIdx.x;
  if (t >= n_tokens) return;
  const size_t r = (size_t)row_of_tok[t];
  for (int i = threadIdx.x; i < kHid; i += blockDim.x) out[(size_t)t * kHid + i] = x[r * kHid + i];
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_embed_ln_f32(const int32_t* ids, const int32_t* type_ids, const int32_t* src_tok,
                                const int32_t* pos_of_row, int M, int vocab, const float* word, const float* pos,
                                const float* type, const float* gamma, const float* beta, float eps, float* x,
                                cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  embed_ln_f32_kernel<<<(M + 7) / 8, 256, 0, st>>>(ids, type_ids, src_tok, pos_of_row, M, vocab, word, pos, type, gamma,
                                                   beta, eps, x);
  return cudaGetLastError();
}

cudaError_t launch_sgemm(int epi, const float* A, const float* W, const float* bias, const float* resid, float* C,
                         int M, int N, int K, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  if (N % 128 != 0 || K % 16 != 0) return cudaErrorInvalidValue;
  const dim3 grid(N / 128, (M + 127) / 128);
  switch (epi) {
    case kSBias: sgemm_kernel<kSBias><<<grid, 256, 0, st>>>(A, W, bias, resid, C, M, N, K); break;
    case kSGelu: sgemm_kernel<kSGelu><<<grid, 256, 0, st>>>(A, W, bias, resid, C, M, N, K); break;
    case kSResidual: sgemm_kernel<kSResidual><<<grid, 256, 0, st>>>(A, W, bias, resid, C, M, N, K); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_layernorm_f32(const float* x, int M, const float* gamma, const float* beta, float eps, float* y,
                                 cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  layernorm_f32_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, M, gamma, beta, eps, y);
  return cudaGetLastError();
}

cudaError_t launch_attention_f32(const float* qkv, const QBlock* qblk, int nqb, float* ctx, cudaStream_t st) {
  if (nqb <= 0) return cudaSuccess;
  static DeviceOnce once;
  const size_t smem = (size_t)2 * kMaxSeq * kHeadDim * sizeof(float);
  cudaError_t e = once_per_device(once, [&] {
    return cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (e != cudaSuccess) return e;
  attention_f32_kernel<<<dim3(nqb, kHeads), 128, smem, st>>>(qkv, qblk, ctx);
  return cudaGetLastError();
}

cudaError_t launch_pool_normalize_f32(const float* x, const int32_t* cu_seqlens, const int32_t* row_start, int n_seqs,
                                      int pool_mode, float* out, cudaStream_t st) {
  if (n_seqs <= 0) return cudaSuccess;
  pool_normalize_f32_kernel<<<n_seqs, 128, 0, st>>>(x, cu_seqlens, row_start, n_seqs, pool_mode, out);
  return cudaGetLastError();
}

cudaError_t launch_ce_head_f32(const float* x, const int32_t* row_start, int n_seqs, const float* wp, const float* bp,
                               const float* wc, const float* bc, float* logits, cudaStream_t st) {
  if (n_seqs <= 0) return cudaSuccess;
  ce_head_f32_kernel<<<n_seqs, 384, 0, st>>>(x, row_start, n_seqs, wp, bp, wc, bc, logits);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows_f32f32(const float* x, const int32_t* row_of_tok, int n_tokens, float* out,
                                      cudaStream_t st) {
  if (n_tokens <= 0) return cudaSuccess;
  gather_rows_f32f32_kernel<<<n_tokens, 128, 0, st>>>(x, row_of_tok, n_tokens, out);
  return cudaGetLastError();
}

}  // namespace frs

/*
 * frs_b200.h — C ABI of the B200-native retrieval hot path.
 *
 * This is the drop-in boundary for the three calls the reference makes into
 * third-party libraries on its hot path (all citations are into the reference
 * repository, pythonmailer/financial-rag-system):
 *
 *   SentenceTransformer.encode(texts)     main.py:148, main.py:213, main2.py:171
 *   QdrantClient.query_points(...)        main.py:232-237, main2.py:163
 *   QdrantClient.upsert(points=...)       ingest.py:171-175
 *   CrossEncoder.predict(pairs)           main.py:245, main2.py:166
 *
 * The reference has no FFI of its own (it is pure Python); this header is what
 * a ctypes binding in main.py/ingest.py binds instead of the three packages.
 * INTEGRATION.md shows that binding.
 *
 * Conventions
 *   - every entry point returns 0 (FRS_OK) or a negative FRS_E_* code; the
 *     message is available from frs_last_error() (thread local).  Nothing
 *     throws across the boundary.
 *   - "dev" pointers are device pointers owned by the caller, "host" pointers
 *     are ordinary host memory.  Handles are owned by the library.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default
 *     stream).  Device-pointer entry points are asynchronous on that stream.
 *   - strings never cross the boundary: ticker / document_type are integer
 *     codes (see FRS_CODE_* below), ids are row numbers; the uuid / payload
 *     tables live in the Python shim.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with FRS_E_CUDA.
 */
#ifndef FRS_B200_H_
#define FRS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRS_VERSION 200 /* 0.2.0 */

#define FRS_OK 0
#define FRS_E_INVALID (-1) /* bad argument                               */
#define FRS_E_CUDA (-2)    /* CUDA runtime / driver error                */
#define FRS_E_CAPACITY (-3) /* index full                                */
#define FRS_E_STATE (-4)   /* call not valid in this state               */
#define FRS_E_TIMEOUT (-5) /* a peer rank of a sharded search did not answer within the exchange's time-out */

#define FRS_DIM 384      /* VectorParams(size=384)  ingest.py:89-95        */
#define FRS_MAX_BATCH 32 /* MAX_BATCH_SIZE          main2.py:51            */
#define FRS_MAX_K 32     /* largest limit; the reference asks for 15 (main.py:215) and 5 (evaluate.py:86) */

/* storage / arithmetic modes of the chunk store */
#define FRS_DTYPE_F32 0  /* rows kept in fp32, TF32 tensor-core pre-filter + fp64 rescoring */
#define FRS_DTYPE_BF16 1 /* rows kept in bf16, bf16 tensor-core pre-filter + fp64 rescoring */

/*
 * Per-row payload code (uint32): the keyword payloads the reference filters on
 * (`ticker`, `document_type`; main.py:218-230) folded into one word.
 *   bits  0..23  ticker id        (dictionary kept by the Python shim)
 *   bits 24..30  document_type id
 *   bit  31      tombstone (row deleted / superseded by an upsert)
 * A query carries (code, mask): a row matches iff ((row_code ^ code) & mask) == 0.
 * The shim always sets bit 31 in `mask` so tombstones never match.
 */
#define FRS_CODE_TICKER_MASK 0x00FFFFFFu
#define FRS_CODE_DOCTYPE_SHIFT 24
#define FRS_CODE_DOCTYPE_MASK 0x7F000000u
#define FRS_CODE_TOMBSTONE 0x80000000u

typedef struct frs_index frs_index;
typedef struct frs_encoder frs_encoder;
typedef struct frs_exchange frs_exchange;

/* ---- library ---------------------------------------------------------- */
int frs_version(void);
const char* frs_last_error(void);
/* number of visible CUDA devices, or a negative error code */
int frs_device_count(void);

/* ---- chunk store: replaces the Qdrant collection ----------------------
 * create_collection(VectorParams(size=384, distance=COSINE))  ingest.py:86-96, database.py:111-143 */
int frs_index_create(int device, int dim, int64_t capacity, int dtype, frs_index** out);
int frs_index_destroy(frs_index* idx);
int64_t frs_index_size(const frs_index* idx);
int64_t frs_index_capacity(const frs_index* idx);
int frs_index_dtype(const frs_index* idx);
int frs_index_device(const frs_index* idx);
/* first global row id of this shard (ids returned by search = base + local row) */
int frs_index_set_base(frs_index* idx, int64_t base);
/* number of CTAs of the scan kernel (0 = one per SM).  Results do not depend on it; exposed so
 * the tests can prove that. */
int frs_index_set_scan_grid(frs_index* idx, int grid);
/* SMs the persistent scan kernel leaves free in the PIPELINED entry points, so that the neighbouring batches'
 * prepare / merge / exchange kernels can run beside it; -1 (default) = chosen from the shard size (4 for long
 * scans, up to 12 for short ones).  Results do not depend on it. */
int frs_index_set_pipeline_reserve(frs_index* idx, int sms);
/* 2 (default): consecutive pipelined scans alternate between two streams, so the first CTAs of scan i+1 start on
 * the SMs the last CTAs of scan i leave; 1: strictly one scan after the other (measurement knob). */
int frs_index_set_scan_streams(frs_index* idx, int n);

/* qdrant.upsert(points)  ingest.py:171-175: L2-normalise (cosine collection),
 * convert to the storage dtype and append.  vecs: [n, 384] fp32, codes: [n]. */
int frs_index_add(frs_index* idx, const float* dev_vecs, const uint32_t* dev_codes, int64_t n,
                  void* stream);
int frs_index_add_host(frs_index* idx, const float* host_vecs, const uint32_t* host_codes,
                       int64_t n);
/* overwrite rows [row0, row0+n) (idempotent upsert on an existing id); dev_codes == NULL keeps the stored codes.
 * Writes and searches may be issued concurrently from different threads / streams (the reference serves queries
 * from 25 threads while ingest.py upserts, main2.py:52-53): a search never reads a half-written row. */
int frs_index_set_rows(frs_index* idx, int64_t row0, const float* dev_vecs,
                       const uint32_t* dev_codes, int64_t n, void* stream);
/* overwrite payload codes only (tombstoning) */
int frs_index_set_codes(frs_index* idx, int64_t row0, const uint32_t* dev_codes, int64_t n,
                        void* stream);
/* read stored rows back as fp32 (exact widening of the stored values) */
int frs_index_read_rows(frs_index* idx, int64_t row0, int64_t n, float* dev_out, void* stream);
int frs_index_read_rows_host(frs_index* idx, int64_t row0, int64_t n, float* host_out);
/* raw device pointers (persistence / zero-copy fill); rows are [capacity, 384] of the dtype */
void* frs_index_rows_ptr(frs_index* idx);
uint32_t* frs_index_codes_ptr(frs_index* idx);
/* declare that rows [0, n) were filled in place through frs_index_rows_ptr() */
int frs_index_set_size(frs_index* idx, int64_t n);

/* Persistence (the role of the Qdrant volume, docker-compose.yml:26-27): the stored rows exactly as
 * they sit in HBM (storage dtype, already normalised) and their payload codes.  export copies rows
 * [row0, row0+n) to host memory (n * 384 * (2|4) bytes, n * 4 bytes); import appends n such rows
 * without touching them, so a reloaded index answers every query bit-identically. */
int frs_index_export_raw(frs_index* idx, int64_t row0, int64_t n, void* host_rows, uint32_t* host_codes);
int frs_index_import_raw(frs_index* idx, const void* host_rows, const uint32_t* host_codes, int64_t n);

/* qdrant.query_points(query=vec, limit=k, query_filter=Filter(must=[...]))
 * main.py:215-239, main2.py:160-163 — exact cosine top-k with payload filter.
 *   queries      [nq, 384] fp32, nq <= FRS_MAX_BATCH (need not be normalised)
 *   q_code/mask  [nq] per-query payload predicate (see FRS_CODE_*)
 *   out_scores   [nq, k] fp32, descending; -inf where fewer than k rows match
 *   out_ids      [nq, k] int64 global row ids (base + row); -1 where no row
 * Ordering is (score desc, id asc) on the fp64 dot product of the stored rows
 * with the prepared query — independent of grid size, GPU count and timing. */
int frs_index_search(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                     const uint32_t* dev_q_mask, int nq, int k, float* dev_out_scores,
                     int64_t* dev_out_ids, void* stream);
int frs_index_search_host(frs_index* idx, const float* host_queries, const uint32_t* host_q_code,
                          const uint32_t* host_q_mask, int nq, int k, float* host_out_scores,
                          int64_t* host_out_ids);

/* Pipelined forms.  The search is three kernels (prepare queries -> scan -> merge); these entry points run them
 * on three internal streams of the index, so that consecutive calls overlap: the preparation of batch i+1 and
 * the merge (+ cross-shard exchange) of batch i-1 run while batch i is being scanned, and the scan kernels run
 * back to back.  This is how the 32-query dynamic batches of main2.py:281-295 are meant to be fed.
 *
 *   frs_index_search_async   device buffers.  Inputs are read in `in_stream` order; results are complete once
 *                            the ticket has been waited for: frs_index_wait (makes a stream wait, ticket -1 =
 *                            everything submitted so far) or frs_index_sync (blocks the host).  A ticket stays
 *                            valid for the next 15 submissions.  `ex` != NULL: sharded search, the shard's local
 *                            top-k is exchanged with the other ranks (see frs_exchange_*) and the outputs are the
 *                            global top-k; every rank submits the same batches in the same order.
 *   frs_index_search_host_submit / _collect
 *                            host buffers: one pinned H2D copy and one D2H copy per batch, both overlapping the
 *                            neighbouring batches' scans.  Up to 4 batches in flight per index (a 5th submit
 *                            blocks until a collect).  frs_index_search_host == submit + collect; it may be
 *                            called from many threads at once (main2.py:52-53: 25 concurrent requests), each call
 *                            taking its own staging slot. */
int frs_index_search_async(frs_index* idx, frs_exchange* ex, const float* dev_queries, const uint32_t* dev_q_code,
                           const uint32_t* dev_q_mask, int nq, int k, float* dev_out_scores,
                           int64_t* dev_out_ids, void* in_stream, int* ticket);
int frs_index_wait(frs_index* idx, int ticket, void* stream);
int frs_index_sync(frs_index* idx, int ticket);
int frs_index_search_host_submit(frs_index* idx, frs_exchange* ex, const float* host_queries,
                                 const uint32_t* host_q_code, const uint32_t* host_q_mask, int nq, int k,
                                 int* ticket);
int frs_index_search_host_collect(frs_index* idx, frs_exchange* ex, int ticket, float* host_out_scores,
                                  int64_t* host_out_ids);

/* Ticker-segmented search: the same exact search restricted to the listed 128-row tiles
 * (tile t = rows [128 t, 128 t + 128); ascending, unique).  The caller guarantees that every row that
 * can match ANY query's predicate lies in a listed tile — ingest is per ticker (ingest.py:109-177), so
 * a ticker's rows occupy few tiles and a filtered batch reads only those.  Rows inside the tiles are
 * still filtered by the per-query predicate, so the result is identical to frs_index_search. */
int frs_index_search_tiles(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                           const uint32_t* dev_q_mask, int nq, int k, const uint32_t* dev_tile_ids,
                           int64_t n_tiles, float* dev_out_scores, int64_t* dev_out_ids, void* stream);

/* Sharded search (one shard per GPU / process): local pass that leaves the
 * shard's exact top-k as (fp64 score, int64 global id) pairs for the exchange
 * step, and the final merge over the gathered [n_shards, nq, k] candidates. */
int frs_index_search_local(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                           const uint32_t* dev_q_mask, int nq, int k, double* dev_out_scores64,
                           int64_t* dev_out_ids, void* stream);
int frs_merge_shards(int device, const double* dev_scores64, const int64_t* dev_ids, int n_shards,
                     int nq, int k, float* dev_out_scores, int64_t* dev_out_ids, void* stream);
/* same, over the exchange buffer [n_shards][2][nq][k] of 64-bit words (plane 0 = fp64 score bits,
 * plane 1 = int64 ids) so that one all-gather moves both halves of the candidates */
int frs_merge_shards_packed(int device, const int64_t* dev_packed, int n_shards, int nq, int k,
                            float* dev_out_scores, int64_t* dev_out_ids, void* stream);

/* Cross-shard exchange over NVLink peer memory (replaces the all-gather of the sharded search; the reference
 * has one Qdrant server and no counterpart, main.py:215-239).  One process per GPU: every rank creates an exchange
 * (nq_max <= 32, k_max <= 32: ONE exchange serves every batch size and limit up to those), the 128-byte handles
 * (two CUDA IPC handles) are all-gathered by the host, frs_exchange_connect maps the peers' buffers.
 * Per batch: the local pass pushes this rank's exact top-k into every peer's gather buffer and publishes a
 * sequence number (fused into the merge kernel: frs_index_search_push / frs_index_search_async /
 * frs_index_search_host_submit with `ex`; frs_exchange_push is the stand-alone form for a block in the exchange's
 * layout [2][nq_max][k_max], entry (q, r) of a plane at q * k + r); frs_exchange_wait_merge[_n] waits for all
 * ranks' pushes of that sequence number and merges to the global top-k.  Every rank issues the same batches
 * (same nq, k) in the same order.  frs_exchange_connect_local links several exchanges of ONE process by pointer
 * (several shards on one GPU, or one shard per GPU with peer access).
 * A peer that does not publish within the time-out (default 30 s) poisons the exchange: that batch and all later
 * ones return empty results, the host-buffer entry points and frs_exchange_status return FRS_E_TIMEOUT, and the
 * exchange has to be destroyed and re-created on every rank.  Nothing traps; the shard stays resident. */
int frs_exchange_create(int device, int world, int rank, int nq_max, int k_max, frs_exchange** out);
int frs_exchange_destroy(frs_exchange* ex);
int frs_exchange_handle(frs_exchange* ex, uint8_t* out128);
int frs_exchange_connect(frs_exchange* ex, const uint8_t* handles_world_by_128);
int frs_exchange_connect_local(frs_exchange* ex, frs_exchange* const* peers);
int frs_exchange_push(frs_exchange* ex, const int64_t* dev_local_packed, void* stream);
int frs_exchange_wait_merge(frs_exchange* ex, float* dev_out_scores, int64_t* dev_out_ids, void* stream);
int frs_exchange_wait_merge_n(frs_exchange* ex, int nq, int k, float* dev_out_scores, int64_t* dev_out_ids,
                              void* stream);
int frs_exchange_set_timeout_ms(frs_exchange* ex, int64_t ms);
int frs_exchange_status(frs_exchange* ex);
/* the local pass with the push FUSED into its merge kernel (one CTA per query writes its k results into every
 * peer's gather buffer, the last CTA publishes the flags): replaces frs_index_search_local + frs_exchange_push */
int frs_index_search_push(frs_index* idx, const float* dev_queries, const uint32_t* dev_q_code,
                          const uint32_t* dev_q_mask, int nq, int k, frs_exchange* ex, void* stream);

/* ---- one process, several GPUs ------------------------------------------
 * get_qdrant() hands the reference ONE client object inside ONE server process (main.py:92-95,
 * main2.py:104-108) and retrieve_from_qdrant calls it from a thread pool (main.py:215-239, main2.py:160-163):
 * frs_sharded is that object for a store spread over the GPUs of a box.  Rows are placed block-cyclically
 * (blocks of frs_sharded_block_rows() rows: global row g -> shard (g / B) % n), ids are global row numbers in
 * insertion order exactly as for a single frs_index, and a search returns the same ids and scores as one
 * frs_index holding all rows.  Every shard's merge kernel writes its exact top-k into the collecting GPU's
 * buffer over NVLink peer memory; CUDA events order the GPUs (no collective library, no flag polling).
 * search_host may be called from many threads; submit / collect keep up to 4 batches in flight. */
typedef struct frs_sharded frs_sharded;
int frs_sharded_create(int n_devices, const int* devices /* NULL = 0..n-1 */, int dim, int64_t capacity_total,
                       int dtype, frs_sharded** out);
int frs_sharded_destroy(frs_sharded* sh);
int frs_sharded_n_shards(const frs_sharded* sh);
int64_t frs_sharded_size(const frs_sharded* sh);
int64_t frs_sharded_capacity(const frs_sharded* sh);
int64_t frs_sharded_block_rows(const frs_sharded* sh);
/* borrow shard s (device-side fill through frs_index_add in block order, diagnostics); owned by `sh` */
frs_index* frs_sharded_shard(frs_sharded* sh, int s);
int frs_sharded_set_size(frs_sharded* sh, int64_t n);
int frs_sharded_add_host(frs_sharded* sh, const float* host_vecs, const uint32_t* host_codes, int64_t n);
/* overwrite global rows [row0, row0+n): vectors (+ codes), or codes only when host_vecs == NULL */
int frs_sharded_set_rows_host(frs_sharded* sh, int64_t row0, const float* host_vecs, const uint32_t* host_codes,
                              int64_t n);
int frs_sharded_read_rows_host(frs_sharded* sh, int64_t row0, int64_t n, float* host_out);
int frs_sharded_export_raw(frs_sharded* sh, int64_t row0, int64_t n, void* host_rows, uint32_t* host_codes);
int frs_sharded_import_raw(frs_sharded* sh, const void* host_rows, const uint32_t* host_codes, int64_t n);
int frs_sharded_search_host(frs_sharded* sh, const float* host_queries, const uint32_t* host_q_code,
                            const uint32_t* host_q_mask, int nq, int k, float* host_out_scores,
                            int64_t* host_out_ids);
int frs_sharded_search_host_submit(frs_sharded* sh, const float* host_queries, const uint32_t* host_q_code,
                                   const uint32_t* host_q_mask, int nq, int k, int* ticket);
int frs_sharded_search_host_collect(frs_sharded* sh, int ticket, float* host_out_scores, int64_t* host_out_ids);

/* the prepared (normalised, storage-dtype-rounded) queries of the last search,
 * widened to fp32: what the scores are dot products with.  [FRS_MAX_BATCH, 384] */
int frs_index_last_queries(frs_index* idx, float* dev_out, void* stream);
/* diagnostics: raw tensor-core pre-filter scores of every row, [nq_pad=32, n] fp32
 * (row-major by query).  Test / profiling aid, not a product path. */
int frs_index_debug_scores(frs_index* idx, const float* dev_queries, int nq, float* dev_out,
                           void* stream);
/* counters of the last search on this index: [0] candidates appended in the scan,
 * [1] list compactions, [2] in-scan exact resolutions, [3] rows rescored in the merge,
 * [4] scan grid size, [5] kernels launched by the last search call */
int frs_index_last_stats(frs_index* idx, int64_t* host_out6);

/* Profiling (off by default).  mode 1: CUDA events are recorded on the caller's stream around each
 * kernel of a search (prep, scan, merge); mode 2: additionally the scan kernel stamps a per-CTA
 * timeline.  read_profile synchronises and returns {searches, prep ms, scan ms, merge ms} summed
 * over the searches recorded since the last read (at most the last 256). */
/* mode 3 (bracket): only ONE event before the first scan kernel and one after the latest, on the stream the scan
 * runs on: read_profile[_ex] then returns {scans, 0, ms from the first scan's start to the last scan's end, ...} —
 * the scan's average launch duration (gaps between launches included) with no events between a step's kernels. */
int frs_index_set_profiling(frs_index* idx, int mode);
int frs_index_read_profile(frs_index* idx, double* host_out4);
/* {searches, prep ms, scan ms, merge ms, exchange ms (cross-shard wait + merge), scan-stream gap ms (end of one
 * scan kernel to the start of the next, summed), span ms (first prep start to last search end), 0} */
int frs_index_read_profile_ex(frs_index* idx, double* host_out8);
/* bracket mode (3) relative to the caller's events (cudaEvent_t handles recorded on this device): out2 = {ms from ev_before
 * to the first scan kernel's start, ms from the last scan kernel's end to ev_after} = fill and drain of a pipelined run;
 * call before frs_index_read_profile_ex (which resets the bracket) */
int frs_index_read_profile_bracket_rel(frs_index* idx, void* ev_before, void* ev_after, double* host_out2);
/* raw time line (diagnostics): per recorded search, oldest first, 7 event times in ms relative to the first search's first
 * event {prep start, prep end, scan start, scan end, merge start, merge end, exchange end}; returns the number of searches
 * written (<= max_searches), does not reset the recording */
int frs_index_read_profile_raw(frs_index* idx, double* host_out, int max_searches);
/* [n_ctas, 16] globaltimer ns: start, first slab landed, last MMA issued, first tile consumed,
 * last tile consumed, exit, then stamps of the first rare-path invocation (diagnostics) */
int frs_index_read_timeline(frs_index* idx, uint64_t* host_out, int n_ctas);

/* ---- encoders: replace SentenceTransformer.encode / CrossEncoder.predict -
 * BertModel forward (transformers/models/bert/modeling_bert.py) for the two
 * checkpoints of main.py:84,90 (get_embedder / get_reranker).  Tokenisation
 * (WordPiece) stays on the host; the library takes PACKED token ids: sequence s
 * owns ids[cu_seqlens[s] .. cu_seqlens[s+1]), no padding anywhere. */
typedef struct frs_bert_cfg {
  int32_t vocab_size;    /* 30522 */
  int32_t hidden;        /* 384   */
  int32_t layers;        /* 12 (bge-small-en-v1.5) / 6 (ms-marco-MiniLM-L-6-v2) */
  int32_t heads;         /* 12    */
  int32_t intermediate;  /* 1536  */
  int32_t max_pos;       /* 512   */
  int32_t type_vocab;    /* 2     */
  int32_t has_head;      /* 1 = pooler + 1-logit classifier (cross-encoder) */
  float ln_eps;          /* 1e-12 */
  int32_t precision;     /* FRS_PRECISION_BF16 (tensor cores) or FRS_PRECISION_F32 (fp32 FFMA, parity mode) */
} frs_bert_cfg;

#define FRS_PRECISION_BF16 0 /* bf16 operands + activations, fp32 accumulation: |embedding error| ~1e-3  */
#define FRS_PRECISION_F32 1  /* fp32 everywhere, no tensor cores, ~20x slower: |error| < 1e-5          */

#define FRS_MAX_LAYERS 12
#define FRS_MAX_SEQ 512 /* max_position_embeddings; SentenceTransformer / CrossEncoder truncate here */
#define FRS_POOL_CLS 0  /* bge-small-en-v1.5: pooling_mode_cls_token */
#define FRS_POOL_MEAN 1 /* masked mean (all-MiniLM-L6-v2 of evaluate.py:22) */

/* Weight table of frs_encoder_create: fp32 tensors in HF state_dict layout ([out, in] Linear
 * weights), in this order (n = 5 + 16 * layers + 4 * has_head):
 *   0 embeddings.word_embeddings [vocab,384]   1 position_embeddings [512,384]
 *   2 token_type_embeddings [2,384]            3,4 embeddings.LayerNorm weight, bias
 *   per layer l, base 5 + 16 l:
 *     +0,+1 attention.self.query w,b    +2,+3 key w,b      +4,+5 value w,b
 *     +6,+7 attention.output.dense w,b  +8,+9 attention.output.LayerNorm w,b
 *     +10,+11 intermediate.dense w [1536,384], b           +12,+13 output.dense w [384,1536], b
 *     +14,+15 output.LayerNorm w,b
 *   head: pooler.dense w,b ; classifier w [1,384], b [1]
 * GEMM weights are converted to bf16 once at creation; biases / LayerNorm / embeddings stay fp32. */
#define FRS_BERT_WEIGHTS(layers, has_head) (5 + 16 * (layers) + ((has_head) ? 4 : 0))

/* weights: array of n_weights pointers (host memory if !on_device, else device memory of `device`).
 * max_tokens: token rows one forward pass can hold (workspace is ~7.7 KB per row); internally every
 * sequence starts on a row that is a multiple of 8, so a pass holds sum roundup8(len_s) <= max_tokens. */
int frs_encoder_create(int device, const frs_bert_cfg* cfg, const float* const* weights, int n_weights,
                       int on_device, int max_tokens, frs_encoder** out);
int frs_encoder_destroy(frs_encoder* enc);
int frs_encoder_max_tokens(const frs_encoder* enc);

/* SentenceTransformer.encode(texts)  main.py:148,213 main2.py:171 — BERT forward, pooling, L2 normalise.
 *   dev_ids          [total] int32 packed WordPiece ids ([CLS] ... [SEP] per sequence)
 *   host_cu_seqlens  [n_seqs+1] int32 prefix sums (HOST memory), every length in [1, FRS_MAX_SEQ]
 *   dev_out          [n_seqs, 384] fp32, rows L2-normalised
 * The device-pointer form needs total <= max_tokens; the *_host form takes host buffers of any
 * size and runs as many passes as needed. */
int frs_encoder_embed(frs_encoder* enc, const int32_t* dev_ids, const int32_t* host_cu_seqlens, int n_seqs,
                      int pool_mode, float* dev_out, void* stream);
int frs_encoder_embed_host(frs_encoder* enc, const int32_t* host_ids, const int32_t* host_cu_seqlens,
                           int n_seqs, int pool_mode, float* host_out);
/* CrossEncoder.predict(pairs)  main.py:245 main2.py:166 — [CLS] q [SEP] d [SEP] with token types 0/1,
 * BERT forward, pooler (tanh) and the 1-logit classifier; raw logits (no activation). */
int frs_encoder_score_pairs(frs_encoder* enc, const int32_t* dev_ids, const int32_t* dev_type_ids,
                            const int32_t* host_cu_seqlens, int n_seqs, float* dev_logits, void* stream);
int frs_encoder_score_pairs_host(frs_encoder* enc, const int32_t* host_ids, const int32_t* host_type_ids,
                                 const int32_t* host_cu_seqlens, int n_seqs, float* host_logits);
/* last_hidden_state of the most recent forward pass, widened to fp32: [n_tokens, 384] (test aid) */
int frs_encoder_last_hidden(frs_encoder* enc, float* dev_out, int n_tokens, void* stream);
/* diagnostics: the first n_elems bf16 values of a workspace buffer of the most recent pass, widened
 * to fp32.  which: 0 x0 [T,384] (layer output), 1 x1 [T,384] (after attention block), 2 qk [T,768]
 * (scaled q | k), 3 vt [384,T] (v transposed, T = max_tokens), 4 ctx [T,384], 5 h [T,1536].
 * With a 1-layer model this exposes every stage of a layer.  Rows are in the library's internal
 * layout: sequence s starts at row sum_{j<s} roundup8(len_j).  Test aid, not a product path. */
int frs_encoder_debug_read(frs_encoder* enc, int which, float* dev_out, int64_t n_elems, void* stream);
/* Profiling (off by default): CUDA events around every kernel of a forward pass.  read_profile
 * synchronises and returns the milliseconds of the LAST pass per kernel class:
 * [0] embeddings+LN, [1] QKV GEMM, [2] attention, [3] out-proj GEMM+LN, [4] FFN-up GEMM+GELU,
 * [5] FFN-down GEMM+LN, [6] pooling / head, [7] kernels launched */
int frs_encoder_set_profiling(frs_encoder* enc, int on);
int frs_encoder_read_profile(frs_encoder* enc, double* host_out8);

#ifdef __cplusplus
}
#endif
#endif /* FRS_B200_H_ */

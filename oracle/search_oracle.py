"""CPU oracle for the retrieval step — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product path (financial_rag_system_b200) never does.

PARITY UNPINNED.  The reference (pythonmailer/financial-rag-system) performs this step through
`QdrantClient.query_points(collection_name, query=vec, limit=15, query_filter=Filter(must=[...]))`
(main.py:215-239, main2.py:160-163) on a collection created with
`VectorParams(size=384, distance=Distance.COSINE)` (ingest.py:86-96, database.py:111-143).
qdrant-client / the Qdrant server are un-vendored and un-pinned dependencies (requirements.txt:21,
docker-compose.yml:22 `qdrant/qdrant:latest`), are not installed in this image, and the reference's
own tests stub the call (`TESTING` → `points=[]`, main.py:216), so there is no golden vector or
known-answer test to pin against.  What is restated here is the published behaviour of Qdrant's
exact cosine search as recalled from qdrant-client's local mode:

  * on upsert a COSINE collection L2-normalises every vector in float32;
  * on query the query vector is L2-normalised, the score is the dot product with every stored
    vector, points whose payload does not satisfy the `must` keyword conditions are masked out,
    and the `limit` best scores are returned in descending order.

The reference's order among exactly tied scores is implementation defined (`np.argsort()[::-1]` /
a binary heap), so the oracle fixes it: (score descending, row id ascending).  Scores are the
float64 dot products of the *stored* (possibly bf16-rounded) rows with the *prepared* query, which
is what the CUDA path reports as well.
"""
from __future__ import annotations

import numpy as np

DIM = 384
CODE_TOMBSTONE = 0x80000000


# ----------------------------------------------------------------------------------------------
# storage conventions
# ----------------------------------------------------------------------------------------------
def l2_normalize_f32(x: np.ndarray) -> np.ndarray:
    """Cosine collection: `v / ||v||` in float32 (zero vectors stay zero)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = np.sqrt(np.sum(x * x, axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    out = np.divide(x, n, out=np.zeros_like(x), where=n > 0)
    return out.astype(np.float32)


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 (round to nearest even) -> float32, with numpy bit arithmetic."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    # NaN stays NaN
    is_nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = np.where(is_nan, u | 0x00400000, r)
    return (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32).reshape(np.shape(x))


def store_rows(x: np.ndarray, dtype: str) -> np.ndarray:
    """What the chunk store keeps for `x`: normalised, rounded to the storage dtype, as float32."""
    y = l2_normalize_f32(x)
    return round_to_bf16(y) if dtype == "bf16" else y


def prepare_queries(q: np.ndarray, dtype: str) -> np.ndarray:
    """What the scores are dot products with: normalised query, rounded like the stored rows."""
    return store_rows(q, dtype)


# ----------------------------------------------------------------------------------------------
# exact search
# ----------------------------------------------------------------------------------------------
def scores_f64(rows_f32: np.ndarray, queries_f32: np.ndarray, chunk: int = 262144) -> np.ndarray:
    """[nq, n] float64 dot products of stored rows with prepared queries."""
    q64 = np.ascontiguousarray(queries_f32, dtype=np.float64)
    n = rows_f32.shape[0]
    out = np.empty((q64.shape[0], n), dtype=np.float64)
    for s in range(0, n, chunk):
        out[:, s : s + chunk] = (rows_f32[s : s + chunk].astype(np.float64) @ q64.T).T
    return out


def payload_mask(codes: np.ndarray, q_code: int, q_mask: int) -> np.ndarray:
    """Keyword `must` conditions folded into integers: row matches iff ((code ^ q_code) & q_mask) == 0."""
    c = np.asarray(codes, dtype=np.uint32)
    return ((c ^ np.uint32(q_code)) & np.uint32(q_mask)) == 0


def exact_topk(rows_f32, queries_f32, codes, q_codes, q_masks, k: int, base: int = 0):
    """Exact top-k per query: ids int64 [nq,k] (-1 padded), scores float64 [nq,k] (-inf padded).

    rows_f32 / queries_f32 are the *stored* rows and *prepared* queries (see store_rows /
    prepare_queries, or read them back from the CUDA index).  Order: score desc, id asc.
    """
    rows_f32 = np.asarray(rows_f32, dtype=np.float32)
    queries_f32 = np.asarray(queries_f32, dtype=np.float32)
    nq = queries_f32.shape[0]
    n = rows_f32.shape[0]
    ids = np.full((nq, k), -1, dtype=np.int64)
    sc = np.full((nq, k), -np.inf, dtype=np.float64)
    if n == 0:
        return ids, sc
    s = scores_f64(rows_f32, queries_f32)
    row_ids = np.arange(n, dtype=np.int64)
    for qi in range(nq):
        m = payload_mask(codes, int(q_codes[qi]), int(q_masks[qi]))
        cand = row_ids[m]
        if cand.size == 0:
            continue
        cs = s[qi, m]
        if cand.size > 4 * k:
            # cut to everything >= the k-th best score (keeps all ties), then order exactly
            kth = np.partition(cs, cand.size - k)[cand.size - k]
            keep = cs >= kth
            cand, cs = cand[keep], cs[keep]
        order = np.lexsort((cand, -cs))[:k]
        ids[qi, : order.size] = cand[order] + base
        sc[qi, : order.size] = cs[order]
    return ids, sc


def merge_shards(ids_list, scores_list, k: int):
    """Final merge of per-shard exact top-k lists ([nq,k] each) by (score desc, id asc)."""
    ids = np.concatenate(ids_list, axis=1)
    sc = np.concatenate(scores_list, axis=1)
    nq = ids.shape[0]
    out_i = np.full((nq, k), -1, dtype=np.int64)
    out_s = np.full((nq, k), -np.inf, dtype=np.float64)
    for qi in range(nq):
        v = ids[qi] >= 0
        ci, cs = ids[qi][v], sc[qi][v]
        o = np.lexsort((ci, -cs))[:k]
        out_i[qi, : o.size] = ci[o]
        out_s[qi, : o.size] = cs[o]
    return out_i, out_s


# ----------------------------------------------------------------------------------------------
# "as shipped" restatement used for CPU timing (float32 dot + full argsort, like qdrant-client
# local mode): this is the arithmetic whose cost bench.py reports as the CPU baseline.
# ----------------------------------------------------------------------------------------------
def as_shipped_search(vectors_f32_normalised: np.ndarray, query: np.ndarray, match: np.ndarray | None, limit: int = 15):
    q = np.asarray(query, dtype=np.float32)
    nrm = np.linalg.norm(q)
    if nrm > 0:
        q = q / nrm
    scores = vectors_f32_normalised @ q  # float32 sgemv
    order = np.argsort(scores)[::-1]
    if match is None:
        top = order[:limit]
    else:
        top = order[match[order]][:limit]
    return top.astype(np.int64), scores[top]

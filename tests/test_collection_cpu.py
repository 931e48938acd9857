"""Host-side logic of the chunk store and the reference-facing surface, on CPU: payload codes and
predicates, idempotent upsert, the Qdrant look-alike, rerank ordering, the dynamic batcher.  The GPU
index / encoders are replaced by doubles built on the oracle (tests may use it; the product never)."""
import asyncio

import numpy as np
import pytest

from financial_rag_system_b200 import synth
from financial_rag_system_b200.api import DynamicBatcher, Retriever
from financial_rag_system_b200.collection import Collection, QdrantCompat, models
from oracle import search_oracle as so


class OracleIndex:
    """VectorIndex double: float32 rows on the host, exact search by the oracle."""

    device = "cpu"

    def __init__(self, capacity):
        self.rows = np.zeros((0, 384), dtype=np.float32)
        self.codes = np.zeros(0, dtype=np.uint32)

    def add(self, vecs, codes=None):
        self.rows = np.concatenate([self.rows, so.store_rows(vecs, "f32")])
        self.codes = np.concatenate([self.codes, np.asarray(codes, dtype=np.uint32)])

    def search(self, q, code, mask, k):
        ids, sc = so.exact_topk(self.rows, so.prepare_queries(q, "f32"), self.codes, code, mask, k)
        return ids, sc.astype(np.float32)

    dtype = "f32"
    tiles_seen = None

    def search_tiles(self, q, code, mask, k, tile_ids):
        """Restricted scan double: ONLY rows of the listed tiles exist for this search."""
        import torch

        tiles = np.asarray(tile_ids, dtype=np.int64)
        OracleIndex.tiles_seen = tiles
        keep = np.zeros(len(self.rows), dtype=bool)
        for t in tiles:
            keep[t * 128:(t + 1) * 128] = True
        sel = np.flatnonzero(keep)
        ids, sc = so.exact_topk(self.rows[sel], so.prepare_queries(q, "f32"), self.codes[sel], code, mask, k)
        ids = np.where(ids >= 0, sel[np.maximum(ids, 0)] if len(sel) else -1, -1)
        return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(sc.astype(np.float32))

    def set_rows(self, row0, vecs, codes=None):
        v = so.store_rows(np.asarray(vecs, dtype=np.float32), "f32")
        self.rows[row0:row0 + len(v)] = v
        if codes is not None:
            self.codes[row0:row0 + len(v)] = np.asarray(codes).astype(np.int64).astype(np.uint32)

    def export_raw(self, row0=0, n=None):
        n = len(self.rows) - row0 if n is None else n
        return self.rows[row0:row0 + n].copy(), self.codes[row0:row0 + n].copy()

    def import_raw(self, rows, codes):
        self.rows = np.concatenate([self.rows, np.asarray(rows, dtype=np.float32)])
        self.codes = np.concatenate([self.codes, np.asarray(codes, dtype=np.uint32)])


def _collection(n=400, seed=0):
    rng = np.random.default_rng(seed)
    ids, texts, payloads = synth.make_chunks(n, n_tickers=6, seed=5)
    vecs = rng.standard_normal((n, 384)).astype(np.float32)
    c = Collection(n + 50, index=OracleIndex(n + 50))
    c.upsert(ids, vecs, payloads)
    return c, ids, vecs, payloads


def test_filter_is_the_and_of_keyword_equalities():
    c, ids, vecs, payloads = _collection()
    t = payloads[0]["ticker"]
    got, scores = c.search(vecs[:3], t, limit=15)
    for row in got.ravel():
        assert row < 0 or payloads[row]["ticker"] == t
    got2, _ = c.search(vecs[:3], t, limit=15, document_type="10-K")
    for row in got2.ravel():
        assert row < 0 or (payloads[row]["ticker"] == t and payloads[row]["document_type"] == "10-K")
    # the exact result, by brute force over the matching rows
    x = so.l2_normalize_f32(vecs)
    m = np.array([p["ticker"] == t for p in payloads])
    s = (x.astype(np.float64) @ x[0].astype(np.float64))
    s[~m] = -np.inf
    assert got[0].tolist() == np.lexsort((np.arange(len(s)), -s))[:15].tolist()
    assert np.all(np.diff(scores[0][np.isfinite(scores[0])]) <= 0)


def test_unknown_ticker_matches_nothing_and_per_query_tickers():
    c, ids, vecs, payloads = _collection()
    got, scores = c.search(vecs[:2], "NOPE")
    assert (got == -1).all() and np.isneginf(scores).all()
    ts = [payloads[0]["ticker"], payloads[7]["ticker"]]
    got, _ = c.search(vecs[[0, 7]], ts, limit=5)
    assert got[0, 0] == 0 and got[1, 0] == 7
    assert all(payloads[r]["ticker"] == ts[1] for r in got[1] if r >= 0)
    with pytest.raises(ValueError):
        c.search(vecs[:2], ["A"], limit=5)


def test_more_than_32_queries_are_searched_in_slices():
    c, ids, vecs, payloads = _collection()
    got, _ = c.search(vecs[:70], None, limit=3)
    assert got.shape == (70, 3) and got[:, 0].tolist() == list(range(70))


def test_upsert_appends_new_ids_and_is_idempotent_within_a_call():
    c, ids, vecs, payloads = _collection(50)
    n = len(c)
    c.upsert(["x", "x"], np.stack([vecs[0], vecs[1]]), [payloads[0], payloads[1]])
    assert len(c) == n + 1 and c.payloads[-1] == payloads[1]
    with pytest.raises(ValueError):
        c.upsert(["a"], vecs[:2], [payloads[0]])


def test_qdrant_lookalike_follows_the_reference_call_sites():
    """ingest.py:86-96,148-175  main.py:215-239 — built exactly as the reference builds them."""
    q = QdrantCompat(capacity=500, index_factory=OracleIndex)
    name = "financial_documents"
    assert not q.collection_exists(name)
    q.create_collection(collection_name=name, vectors_config=models.VectorParams(size=384, distance=models.Distance.COSINE))
    assert q.collection_exists(name) and [c.name for c in q.get_collections().collections] == [name]
    with pytest.raises(ValueError):
        q.create_collection(collection_name="bad", vectors_config=models.VectorParams(size=128, distance=models.Distance.COSINE))
    ids, texts, payloads = synth.make_chunks(300, n_tickers=5, seed=2)
    vecs = np.random.default_rng(1).standard_normal((300, 384)).astype(np.float32)
    pts = [models.PointStruct(id=i, vector=v.tolist(), payload=p) for i, v, p in zip(ids, vecs, payloads)]
    for s in range(0, 300, 256):  # UPSERT_BATCH = 256
        q.upsert(collection_name=name, points=pts[s:s + 256])
    t = payloads[3]["ticker"]
    must = [models.FieldCondition(key="ticker", match=models.MatchValue(value=t.upper()))]
    res = q.query_points(collection_name=name, query=vecs[3].tolist(), limit=15, query_filter=models.Filter(must=must))
    assert 1 <= len(res.points) <= 15 and res.points[0].id == ids[3]
    assert abs(res.points[0].score - 1.0) < 1e-5
    assert all(p.payload["ticker"] == t and "text" in p.payload for p in res.points)
    assert [p.score for p in res.points] == sorted((p.score for p in res.points), reverse=True)
    none = q.query_points(collection_name=name, query=vecs[3].tolist(), limit=15,
                          query_filter=models.Filter(must=[models.FieldCondition(key="ticker", match=models.MatchValue(value="ZZZZ"))]))
    assert none.points == []


class _FakeEmbedder:
    def __init__(self, table):
        self.table, self.calls = table, 0

    def encode(self, texts):
        self.calls += 1
        if isinstance(texts, str):
            return self.table[texts]
        return np.stack([self.table[t] for t in texts])

    def close(self):
        pass


class _FakeReranker:
    def __init__(self):
        self.calls = 0

    def predict(self, pairs):
        self.calls += 1
        return np.array([float(len(set(q.split()) & set(t.split()))) + 0.001 * (len(t) % 7) for q, t in pairs], dtype=np.float32)

    def close(self):
        pass


def _retriever():
    c, ids, vecs, payloads = _collection(300)
    qs, ts = synth.make_queries(40, n_tickers=6)
    ts = [payloads[i]["ticker"] for i in range(len(qs))]
    table = {q: vecs[i] + 0.01 for i, q in enumerate(qs)}
    return Retriever(c, _FakeEmbedder(table), _FakeReranker()), qs, ts, payloads


def test_rerank_contract_of_the_reference():
    """rerank_documents, main.py:241-247: (argsort(scores)[::-1][:top_k], scores); empty -> ([], zeros(0))."""
    r, qs, ts, payloads = _retriever()
    idx, scores = r.rerank("net sales increased", ["net sales fell", "unrelated text", "net sales increased again"], 2)
    assert list(idx) == list(np.argsort(scores)[::-1][:2]) and scores.shape == (3,)
    idx, scores = r.rerank("q", [], 5)
    assert list(idx) == [] and scores.shape == (0,)


def test_retrieve_batch_equals_the_per_query_pipeline_and_uses_three_passes():
    r, qs, ts, payloads = _retriever()
    batch = r.retrieve_batch(qs[:32], ts[:32], top_k=5)
    assert r.embedder.calls == 1 and r.reranker.calls == 1
    for i in (0, 5, 31):
        vec = r.embed(qs[i])
        ids, scores = r.search(vec, ts[i], 15)
        rows = [int(x) for x in ids[0] if x >= 0]
        idx, logits = r.rerank(qs[i], [payloads[x]["text"] for x in rows], 5)
        assert [h.row for h in batch[i]] == [rows[j] for j in idx]
        assert all(h.payload["ticker"] == ts[i] for h in batch[i])


def test_dynamic_batcher_batches_search_and_rerank_too():
    """main2.py:281-295 policy (first request + 50 ms window, <= 32 per batch)."""
    r, qs, ts, payloads = _retriever()

    async def run():
        b = DynamicBatcher(r, window_s=0.02)
        b.start()
        res = await asyncio.gather(*[b.submit(qs[i], ts[i], 3) for i in range(40)])
        await b.stop()
        return b, res

    b, res = asyncio.run(run())
    assert len(res) == 40 and all(len(h) <= 3 for h in res)
    assert b.batches == 2  # 32 + 8
    assert r.embedder.calls == 2 and r.reranker.calls == 2
    direct = r.retrieve_batch(qs[:1], ts[:1], top_k=3)
    assert [h.row for h in res[0]] == [h.row for h in direct[0]]


def test_save_and_load_round_trip(tmp_path):
    c, ids, vecs, payloads = _collection(120)
    t = payloads[5]["ticker"]
    before = c.search(vecs[:8], t, limit=10)
    c.save(str(tmp_path / "col"))
    d = Collection.load(str(tmp_path / "col"), index=OracleIndex(0))
    assert len(d) == len(c) and d.ids == c.ids and d.payloads == c.payloads
    after = d.search(vecs[:8], t, limit=10)
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    # the dictionaries survive: a new upsert with a known ticker gets the same code, ids stay idempotent
    d.upsert([ids[0]], vecs[3:4], [payloads[0]])  # overwrite in place: row 0 now holds vector 3
    assert len(d) == len(c)
    got, sc = d.search(vecs[3:4], payloads[0]["ticker"], limit=2)
    assert got[0, 0] == 0 and abs(sc[0, 0] - 1.0) < 1e-6


def test_ticker_segmented_search_reads_only_the_tickers_tiles_and_changes_nothing():
    """rows ingested ticker by ticker (ingest.py:109-177) -> a filtered batch is answered from
    the tiles of its tickers only, with exactly the ids / scores of the full scan."""
    rng = np.random.default_rng(4)
    n_t, per = 12, 700
    vecs = rng.standard_normal((n_t * per, 384)).astype(np.float32)
    payloads = [{"ticker": f"T{t}", "document_type": ("10-K", "10-Q")[i % 2], "text": ""} for t in range(n_t) for i in range(per)]
    ids = list(range(n_t * per))
    seg = Collection(n_t * per + 10, index=OracleIndex(0))
    assert seg.segmented
    for s in range(0, len(ids), 256):
        seg.upsert(ids[s:s + 256], vecs[s:s + 256], payloads[s:s + 256])
    full = Collection(n_t * per + 10, index=OracleIndex(0))
    full.segmented = False
    full.upsert(ids, vecs, payloads)
    q = vecs[[5, 800, 4000, 8399]] + 0.1
    ts = ["T0", "T1", "T5", "T11"]
    a = seg.search(q, ts, 15, ["10-K", None, None, "10-Q"])
    b = full.search(q, ts, 15, ["10-K", None, None, "10-Q"])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    scanned, total = seg.last_scan_tiles
    assert total == (n_t * per + 127) // 128 and scanned <= 4 * (per // 128 + 2) < total / 2
    # a query without ticker condition needs every row: full scan
    seg.search(q[:1], None, 5)
    assert seg.last_scan_tiles == (total, total)
    # unknown ticker: nothing to read, nothing returned
    got, sc = seg.search(q[:1], "NOPE", 5)
    assert (got == -1).all() and seg.last_scan_tiles[0] == 0
    # a row re-upserted under another ticker is found under the new one
    seg.upsert([ids[3]], vecs[3:4], [{"ticker": "T7", "document_type": "10-K", "text": ""}])
    got, sc = seg.search(vecs[3:4], "T7", 1)
    assert got[0, 0] == 3

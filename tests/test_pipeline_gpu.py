"""End-to-end retrieval through the reference-facing surface on the GPU: embed -> upsert -> search with
ticker filter -> rerank, checked stage by stage against the oracles (config 1 / config 5 of
BASELINE.json at a size the CPU oracle finishes in seconds)."""
import numpy as np
import pytest

from financial_rag_system_b200 import synth
from financial_rag_system_b200.checkpoint import BGE_SMALL, MINILM_L6_CE, synthetic_checkpoint

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack():
    from financial_rag_system_b200.api import Retriever
    from financial_rag_system_b200.collection import Collection
    from financial_rag_system_b200.encoder import Embedder, Reranker
    from financial_rag_system_b200.tokenizer import WordPiece

    tok = WordPiece.synthetic()
    emb = Embedder(device=0, max_tokens=32768, tokenizer=tok)
    rr = Reranker(device=0, max_tokens=32768, tokenizer=tok)
    ids, texts, payloads = synth.make_chunks(600, n_tickers=8, seed=3)
    col = Collection(1000, dtype="bf16", device=0)
    vecs = emb.encode(texts)
    for s in range(0, len(ids), 256):  # UPSERT_BATCH, ingest.py:28
        col.upsert(ids[s:s + 256], vecs[s:s + 256], payloads[s:s + 256])
    r = Retriever(col, emb, rr)
    yield r, tok, ids, texts, payloads, vecs
    r.close()


def test_ingest_embeddings_match_the_oracle(stack):
    from oracle import encoder_oracle as eo

    r, tok, ids, texts, payloads, vecs = stack
    sel = list(range(0, 600, 25))
    pi, pc = tok.pack_texts([texts[i] for i in sel])
    ref = eo.embed(BGE_SMALL, synthetic_checkpoint(BGE_SMALL, 1234), pi, pc)
    assert np.abs(vecs[sel] - ref).max() <= 1e-2
    assert ((vecs[sel] * ref).sum(1)).min() >= 0.999


def test_batched_retrieval_matches_oracle_search_and_rerank(stack):
    from oracle import encoder_oracle as eo
    from oracle import search_oracle as so

    r, tok, ids, texts, payloads, vecs = stack
    qs, _ = synth.make_queries(32, n_tickers=8, seed=4)
    ts = [payloads[(37 * i) % 600]["ticker"] for i in range(32)]
    hits = r.retrieve_batch(qs, ts, top_k=5)
    assert len(hits) == 32 and all(1 <= len(h) <= 5 for h in hits)

    # search stage: exact ids of the oracle on the SAME stored rows / prepared queries
    qv = r.embed(qs)
    got_ids, got_scores = r.search(qv, ts, 15)
    rows = r.collection.index.read_rows().cpu().numpy()
    qp = so.prepare_queries(qv, "bf16")
    codes = r.collection._codes
    pred = [r.collection.predicate(t) for t in ts]
    oi, osc = so.exact_topk(rows, qp, codes, [p[0] for p in pred], [p[1] for p in pred], 15)
    assert np.array_equal(got_ids, oi)
    assert np.allclose(got_scores, osc, atol=1e-6)
    for i in range(32):
        assert all(payloads[x]["ticker"] == ts[i] for x in got_ids[i] if x >= 0)

    # rerank stage: logits of the oracle cross-encoder on the same pairs (first 6 queries)
    wc = synthetic_checkpoint(MINILM_L6_CE, 4321)
    for i in range(6):
        cand = [int(x) for x in got_ids[i] if x >= 0]
        pairs = [[qs[i], payloads[x]["text"]] for x in cand]
        pi, pt, pc = tok.pack_pairs(pairs)
        ref = eo.score_pairs(MINILM_L6_CE, wc, pi, pt, pc)
        idx, logits = r.rerank(qs[i], [p[1] for p in pairs], 5)
        assert np.abs(logits - ref).max() <= 0.12
        order = np.argsort(ref)[::-1]
        gaps = ref[order][:5] - ref[order][1:6] if len(ref) > 5 else np.array([1.0])
        if gaps.min() > 0.25:
            assert list(idx) == list(order[:5])
            assert [h.row for h in hits[i]] == [cand[j] for j in order[:5]]


def test_reference_shaped_calls(stack):
    """embed_query (main.py:211-213), retrieve_from_qdrant-style single query, rerank_documents."""
    r, tok, ids, texts, payloads, vecs = stack
    v = r.embed("What was the effective tax rate?")
    assert v.shape == (384,) and abs(np.linalg.norm(v) - 1.0) < 1e-5 and len(v.tolist()) == 384
    t = payloads[0]["ticker"]
    got, sc = r.search(v, t.upper(), limit=15)
    assert got.shape == (1, 15)
    idx, scores = r.rerank("effective tax rate", [payloads[x]["text"] for x in got[0] if x >= 0], 5)
    assert len(idx) <= 5 and scores.dtype == np.float32


def test_identical_text_is_its_own_nearest_neighbour(stack):
    r, tok, ids, texts, payloads, vecs = stack
    sel = [3, 77, 401]
    got, sc = r.search(r.embed([texts[i] for i in sel]), [payloads[i]["ticker"] for i in sel], 3)
    # duplicates of a chunk text can exist (same sentence windows); the row itself must be among the ties
    for k, i in enumerate(sel):
        assert abs(sc[k, 0] - 1.0) < 2e-3
        assert i in got[k][np.abs(sc[k] - sc[k, 0]) < 1e-6]


def test_evaluation_harness_scores_self_queries_perfectly():
    """evaluate.py:59-128 on the drop-in surface with the GPU embedder for corpus AND queries: a chunk asked for with
    its own text must come back first (Hit@5 = 100 %, MRR = 1.0), whatever the weights."""
    from financial_rag_system_b200 import synth
    from financial_rag_system_b200.collection import QdrantCompat, models
    from financial_rag_system_b200.encoder import Embedder
    from financial_rag_system_b200.evaluate import COLLECTION_NAME, run_evaluation, synthetic_eval_set

    n = 600
    ids, texts, payloads = synth.make_chunks(n, n_tickers=8, seed=5)
    emb = Embedder(device=0, max_tokens=32768)
    q = QdrantCompat(capacity=n)
    q.create_collection(COLLECTION_NAME, models.VectorParams(size=384, distance=models.Distance.COSINE))
    for s in range(0, n, 256):
        vecs = emb.encode(texts[s:s + 256])
        q.upsert(COLLECTION_NAME, [models.PointStruct(id=ids[i], vector=vecs[i - s].tolist(), payload=payloads[i])
                                   for i in range(s, min(n, s + 256))])
    r = run_evaluation(q, emb, synthetic_eval_set(texts, payloads, 40), k=5)
    assert r["hit_rate"] == 100.0 and r["mrr"] == 1.0
    emb.close()

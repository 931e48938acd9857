"""The Hit@k / MRR harness (reference evaluate.py:59-128) on the drop-in surface, with an injected CPU index."""
import numpy as np
import pytest

from financial_rag_system_b200 import synth
from financial_rag_system_b200.collection import QdrantCompat, models
from financial_rag_system_b200.evaluate import COLLECTION_NAME, run_evaluation, synthetic_eval_set
from test_collection_cpu import OracleIndex


class HashEmbedder:
    """Deterministic text -> unit vector (same text, same vector): stands in for the GPU embedder on CPU."""

    def encode(self, texts):
        one = isinstance(texts, str)
        out = []
        for t in ([texts] if one else texts):
            rng = np.random.default_rng(abs(hash(t)) % (1 << 32))
            v = rng.standard_normal(384).astype(np.float32)
            out.append(v / np.linalg.norm(v))
        return out[0] if one else np.stack(out)


def _client(n=300):
    ids, texts, payloads = synth.make_chunks(n, n_tickers=6, seed=5)
    emb = HashEmbedder()
    q = QdrantCompat(capacity=n, index_factory=OracleIndex)
    q.create_collection(COLLECTION_NAME, models.VectorParams(size=384, distance=models.Distance.COSINE))
    vecs = emb.encode(texts)
    q.upsert(COLLECTION_NAME, [models.PointStruct(id=ids[i], vector=vecs[i].tolist(), payload=payloads[i]) for i in range(n)])
    return q, emb, texts, payloads


def test_self_queries_score_perfectly():
    q, emb, texts, payloads = _client()
    r = run_evaluation(q, emb, synthetic_eval_set(texts, payloads, 40), k=5)
    assert r["hit_rate"] == 100.0 and r["mrr"] == 1.0 and r["queries"] == 40 and all(x == 1 for x in r["ranks"])


def test_metric_definitions_follow_the_reference():
    """evaluate.py:97-118: first rank whose text contains any keyword; Hit@k in percent; MRR = mean(1/rank or 0)."""
    q, emb, texts, payloads = _client()
    items = synthetic_eval_set(texts, payloads, 4)
    # item 0: hit at rank 1; item 1: keyword of ANOTHER chunk of the same ticker -> found at some rank r > 1 or missed;
    # item 2: impossible keyword -> miss
    items[2] = dict(items[2], expected_keywords=["\x00 never occurs"])
    r = run_evaluation(q, emb, items[:3], k=5)
    want_rr = [1.0 / x if x else 0.0 for x in r["ranks"]]
    assert r["ranks"][0] == 1 and r["ranks"][2] == 0
    assert r["mrr"] == pytest.approx(sum(want_rr) / 3)
    assert r["hit_rate"] == pytest.approx(100.0 * sum(x > 0 for x in r["ranks"]) / 3)
    # the ticker filter is honoured: a wrong ticker can never hit
    wrong = [dict(items[0], ticker="ZZZZ")]
    assert run_evaluation(q, emb, wrong, k=5)["hit_rate"] == 0.0


def test_missing_collection_is_an_error():
    q = QdrantCompat(capacity=10, index_factory=OracleIndex)
    with pytest.raises(RuntimeError):
        run_evaluation(q, HashEmbedder(), [], k=5)

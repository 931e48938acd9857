"""Retrieval evaluation — the Hit@k / MRR harness of the reference (evaluate.py:59-128) against the drop-in
surface: queries are embedded with `embedder.encode`, searched with `qdrant.query_points(..., limit=k,
query_filter=Filter(must=[ticker == T]))`, and a query counts as a hit at the first rank whose payload text
contains any of its expected keywords (case-insensitive), exactly as evaluate.py:97-118 scores it:

    Hit@k = 100 * (queries with a hit in the top k) / (queries)
    MRR   = mean over queries of 1 / (rank of the first hit), 0 when there is none

One difference, on purpose: the reference embeds the evaluation queries with `all-MiniLM-L6-v2` (evaluate.py:22)
while the collection was built with `BAAI/bge-small-en-v1.5` (main.py:84) — two unrelated vector spaces.  Here the
caller passes the embedder, and `scripts/evaluate_synth.py` uses the SAME one for corpus and queries.

`synthetic_eval_set` builds a ground truth without the network: every item asks about one synthetic chunk and
expects a phrase that occurs in that chunk's text.
"""
from __future__ import annotations

import time
from typing import Sequence

from .collection import models

COLLECTION_NAME = "financial_documents"  # evaluate.py:64, main.py:43


def run_evaluation(qdrant, embedder, dataset: Sequence[dict], k: int = 5, collection_name: str = COLLECTION_NAME,
                   verbose: bool = False) -> dict:
    """dataset items: {"query": str, "ticker": str, "expected_keywords": [str, ...]} (evaluate.py:28-54).
    Returns {"hit_rate": percent, "mrr": float, "avg_latency_ms": float, "ranks": [first-hit rank or 0]}."""
    if not qdrant.collection_exists(collection_name):
        raise RuntimeError(f"collection {collection_name!r} not found")  # evaluate.py:69-72
    hits, rr, ranks, total_latency = 0, [], [], 0.0
    for item in dataset:
        t0 = time.time()
        query_vector = embedder.encode(item["query"]).tolist()                      # evaluate.py:80
        res = qdrant.query_points(                                                  # evaluate.py:83-90
            collection_name=collection_name, query=query_vector, limit=k,
            query_filter=models.Filter(must=[models.FieldCondition(key="ticker", match=models.MatchValue(value=item["ticker"]))]))
        total_latency += (time.time() - t0) * 1000
        found = 0
        for rank, hit in enumerate(res.points, start=1):                            # evaluate.py:97-102
            text = hit.payload.get("text", "").lower()
            if any(kw.lower() in text for kw in item["expected_keywords"]):
                found = rank
                break
        ranks.append(found)
        hits += found > 0
        rr.append(1.0 / found if found else 0.0)                                    # evaluate.py:105-112
        if verbose:
            print(("[HIT]  rank %d" % found if found else "[MISS]        ") + " | " + item["query"][:60])
    n = max(len(dataset), 1)
    return {"k": k, "hit_rate": 100.0 * hits / n, "mrr": sum(rr) / n, "avg_latency_ms": total_latency / n, "ranks": ranks,
            "queries": len(dataset)}


def synthetic_eval_set(texts: Sequence[str], payloads: Sequence[dict], n: int = 50, seed: int = 3, phrase_chars: int = 48,
                       self_queries: bool = True) -> list[dict]:
    """Ground truth over synthetic chunks (synth.make_chunks): item i is about chunk c_i; its expected keyword is a
    phrase cut from the middle of that chunk, its ticker the chunk's.  self_queries=True asks with the chunk's own
    text (an embedder worth its name must then retrieve the chunk first: Hit@k = 100, MRR = 1); False asks with the
    phrase's sentence only, a harder, realistic query."""
    import numpy as np

    rng = np.random.default_rng(seed)
    picks = rng.choice(len(texts), size=min(n, len(texts)), replace=False)
    out = []
    for c in picks:
        text = texts[int(c)]
        mid = len(text) // 2
        phrase = text[mid:mid + phrase_chars]
        if self_queries:
            query = text
        else:
            lo, hi = text.rfind(". ", 0, mid), text.find(". ", mid)
            query = text[lo + 2 if lo >= 0 else 0:hi + 1 if hi >= 0 else len(text)]
        out.append({"query": query, "ticker": payloads[int(c)]["ticker"], "expected_keywords": [phrase], "chunk": int(c)})
    return out

"""The call surface of the reference's hot path (north_star): `embed(texts)`,
`search(query_vecs, ticker, limit=15)`, `rerank(query, chunks, top_k)` — plus the batched forms the
32-query dynamic batch of main2.py:281-295 needs so that search and rerank see the whole batch (in the
reference only the embedding does; retrieval and rerank run one query per thread, main2.py:228,242).

    reference                                              here
    embed_query / embed / embed_query_batch                Retriever.embed(texts)
        main.py:144-149, 211-213  main2.py:170-171
    retrieve_from_qdrant(vec, ticker, doc_type, limit)     Retriever.search(vecs, ticker, limit, document_type)
        main.py:215-239  main2.py:160-163
    rerank_documents(query, texts, top_k)                  Retriever.rerank(query, chunks, top_k)
        main.py:241-247  main2.py:165-168
    batch_processor + process_independently                Retriever.retrieve_batch / DynamicBatcher
        main2.py:207-295

Every arithmetic step is a call into libfrs_b200.so; there is no CPU path.
"""
from __future__ import annotations

import asyncio
import threading
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

MAX_BATCH_SIZE = 32     # main2.py:51
BATCH_WINDOW_S = 0.05   # main2.py:286
SEARCH_LIMIT = 15       # main.py:215


@dataclass
class Hit:
    row: int
    score: float
    rerank_score: float
    payload: dict


class Retriever:
    def __init__(self, collection, embedder=None, reranker=None, device: int = 0):
        """collection: a `Collection`; embedder / reranker default to the synthetic-weight
        `Embedder()` / `Reranker()` (pass `Embedder(model_dir)` for real checkpoints)."""
        from .encoder import Embedder, Reranker

        self.collection = collection
        self.embedder = embedder if embedder is not None else Embedder(device=device)
        self.reranker = reranker if reranker is not None else Reranker(device=device)

    # -- the three calls of the reference -----------------------------------------------------------
    def embed(self, texts) -> np.ndarray:
        """list[str] -> float32 [n,384] L2-normalised (str -> [384]); `.tolist()` gives main.py:149's JSON."""
        return self.embedder.encode(texts)

    def search(self, query_vecs, ticker, limit: int = SEARCH_LIMIT, document_type=None):
        """-> (ids int64 [B,limit], scores float32 [B,limit]); -1 / -inf where fewer rows match."""
        return self.collection.search(query_vecs, ticker, limit, document_type)

    def rerank(self, query: str, chunks: Sequence[str], top_k: int):
        """rerank_documents: (indices of the top_k chunks by descending logit, all logits)."""
        if not chunks:
            return [], np.zeros(0, dtype=np.float32)
        scores = self.reranker.predict([[query, t] for t in chunks])
        return np.argsort(scores)[::-1][:top_k], scores

    # -- batched forms ------------------------------------------------------------------------------
    def rerank_batch(self, queries: Sequence[str], chunk_lists: Sequence[Sequence[str]], top_k: int):
        """One cross-encoder pass over every (query, chunk) pair of the batch (32 x 15 = 480 pairs)."""
        pairs = [[q, t] for q, chunks in zip(queries, chunk_lists) for t in chunks]
        flat = self.reranker.predict(pairs) if pairs else np.zeros(0, dtype=np.float32)
        out, o = [], 0
        for chunks in chunk_lists:
            s = flat[o:o + len(chunks)]
            o += len(chunks)
            out.append((np.argsort(s)[::-1][:top_k], s))
        return out

    def retrieve_batch(self, queries: Sequence[str], tickers: Sequence[str], top_k: int = 5,
                       document_types: Optional[Sequence[Optional[str]]] = None, limit: int = SEARCH_LIMIT):
        """embed -> search -> gather texts -> rerank for a whole dynamic batch: three GPU passes in
        total instead of 1 + 2 x len(batch).  Returns, per query, the top_k `Hit`s in rerank order
        (what process_independently builds as `sources`, main2.py:230-246)."""
        if len(queries) != len(tickers):
            raise ValueError("one ticker per query")
        vecs = self.embed(list(queries))
        ids, scores = self.search(vecs, list(tickers), limit, document_types)
        rows = [[int(r) for r in ids[i] if r >= 0] for i in range(len(queries))]
        texts = [[self.collection.payloads[r].get("text", "") for r in rr] for rr in rows]
        ranked = self.rerank_batch(queries, texts, top_k)
        out = []
        for i, (idx, logits) in enumerate(ranked):
            out.append([Hit(rows[i][j], float(scores[i][j]), float(logits[j]), self.collection.payloads[rows[i][j]])
                        for j in idx])
        return out

    def close(self) -> None:
        self.embedder.close()
        self.reranker.close()
        self.collection.close()


class DynamicBatcher:
    """main2.py's batch engine (request_queue + batch_processor, main2.py:50-53, 281-295) with the same
    policy — take one request, wait 50 ms, drain up to 32 — but the WHOLE retrieval step runs on the
    batch, not only the embedding.  `await batcher.submit(query, ticker, top_k)` resolves to that
    request's hits."""

    def __init__(self, retriever: Retriever, max_batch: int = MAX_BATCH_SIZE, window_s: float = BATCH_WINDOW_S):
        self.retriever, self.max_batch, self.window_s = retriever, max_batch, window_s
        self.queue: asyncio.Queue = asyncio.Queue()
        self.batches = 0
        self._task: Optional[asyncio.Task] = None

    async def submit(self, query: str, ticker: str, top_k: int = 5, document_type: Optional[str] = None):
        fut = asyncio.get_running_loop().create_future()
        await self.queue.put((fut, query, ticker, top_k, document_type))
        return await fut

    def start(self) -> None:
        self._task = asyncio.get_running_loop().create_task(self._run())

    async def stop(self) -> None:
        if self._task:
            self._task.cancel()
            try:
                await self._task
            except asyncio.CancelledError:
                pass

    async def _run(self) -> None:
        while True:
            batch = [await self.queue.get()]
            await asyncio.sleep(self.window_s)
            while not self.queue.empty() and len(batch) < self.max_batch:
                batch.append(self.queue.get_nowait())
            self.batches += 1
            top_k = max(b[3] for b in batch)
            try:
                res = await asyncio.to_thread(self.retriever.retrieve_batch, [b[1] for b in batch], [b[2] for b in batch],
                                              top_k, [b[4] for b in batch])
                for b, hits in zip(batch, res):
                    if not b[0].done():
                        b[0].set_result(hits[:b[3]])
            except Exception as e:  # main2.py:274-276 propagates retrieval errors to the request
                for b in batch:
                    if not b[0].done():
                        b[0].set_exception(e)


# -- module-level convenience: lazy singleton, like get_embedder()/get_reranker()/get_qdrant() ------
_default: Optional[Retriever] = None
_default_lock = threading.Lock()


def configure(retriever: Retriever) -> None:
    global _default
    _default = retriever


def _get() -> Retriever:
    if _default is None:
        raise RuntimeError("call financial_rag_system_b200.api.configure(Retriever(collection, ...)) first")
    return _default


def embed(texts):
    return _get().embed(texts)


def search(query_vecs, ticker, limit: int = SEARCH_LIMIT, document_type=None):
    return _get().search(query_vecs, ticker, limit, document_type)


def rerank(query, chunks, top_k):
    return _get().rerank(query, chunks, top_k)

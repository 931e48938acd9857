#!/usr/bin/env python
"""Hit@k / MRR of the retrieval path on a synthetic SEC-style corpus (the reference's evaluate.py against the
drop-in surface; see financial_rag_system_b200/evaluate.py).  Needs a B200:  python scripts/evaluate_synth.py
[--chunks 10000] [--queries 100] [--k 5] [--model DIR]   (DIR = a local bge-small-en-v1.5 checkpoint; default:
seeded synthetic weights of its shape — random weights still map identical text to identical vectors, so the
self-query set must score Hit@k = 100 %, MRR = 1.0; the sentence-query numbers only mean something with real weights)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=10_000)
    ap.add_argument("--queries", type=int, default=100)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--model", default=None)
    ap.add_argument("--devices", default=None, help="comma-separated CUDA devices to shard the collection over")
    args = ap.parse_args()
    from financial_rag_system_b200 import synth
    from financial_rag_system_b200.collection import QdrantCompat, models
    from financial_rag_system_b200.encoder import Embedder
    from financial_rag_system_b200.evaluate import COLLECTION_NAME, run_evaluation, synthetic_eval_set

    ids, texts, payloads = synth.make_chunks(args.chunks)
    emb = Embedder(args.model)
    devices = [int(d) for d in args.devices.split(",")] if args.devices else None
    qdrant = QdrantCompat(capacity=args.chunks, devices=devices)
    qdrant.create_collection(COLLECTION_NAME, models.VectorParams(size=384, distance=models.Distance.COSINE))
    for s in range(0, args.chunks, 256):                       # ingest.py:27,148-175: embed 64 at a time, upsert 256
        vecs = emb.encode(texts[s:s + 256])
        qdrant.upsert(COLLECTION_NAME, [models.PointStruct(id=ids[i], vector=vecs[i - s].tolist(), payload=payloads[i])
                                        for i in range(s, min(args.chunks, s + 256))])
    out = {}
    for name, self_q in (("self_queries", True), ("sentence_queries", False)):
        r = run_evaluation(qdrant, emb, synthetic_eval_set(texts, payloads, args.queries, self_queries=self_q), k=args.k)
        r.pop("ranks")
        out[name] = r
    print(json.dumps(out, indent=1))
    emb.close()


if __name__ == "__main__":
    main()

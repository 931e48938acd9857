"""INTEGRATION.md is executable documentation: the ctypes stubs it prints are extracted and run against the real
library, so the binding a maintainer would copy can never drift from include/frs_b200.h again (round 1's stub was
missing frs_bert_cfg.precision)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from financial_rag_system_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _python_blocks():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.findall(r"```python\n(.*?)```", text, flags=re.S)


def _stub_namespace():
    """Executes the library-binding blocks (the ones that are complete modules or build on them), in document order."""
    ns = {"__name__": "integration_md"}
    ran = 0
    for block in _python_blocks():
        if "lru_cache" in block or "batcher = DynamicBatcher" in block:
            continue  # application-side glue of main.py (needs the FastAPI app around it)
        code = block.replace('C.CDLL("libfrs_b200.so")', f'C.CDLL({_lib.LIB_PATH!r})')
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
        ran += 1
    assert ran >= 3
    return ns


def test_printed_stubs_execute_and_match_the_header():
    ns = _stub_namespace()
    from financial_rag_system_b200.encoder import _BertCfg

    # struct frs_bert_cfg: same size, same fields in the same order as the shipped binding (and the header)
    assert C.sizeof(ns["_Cfg"]) == C.sizeof(_BertCfg) == 40
    assert [f[0] for f in ns["_Cfg"]._fields_] == [f[0] for f in _BertCfg._fields_]
    header = open(os.path.join(ROOT, "include", "frs_b200.h")).read()
    struct = header[header.index("typedef struct frs_bert_cfg {"):header.index("} frs_bert_cfg;")]
    assert re.findall(r"^\s*(?:int32_t|float)\s+(\w+);", struct, flags=re.M) == [f[0] for f in ns["_Cfg"]._fields_]
    # every entry point the document names exists in the library
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(set(re.findall(r"\b(frs_[a-z0-9_]+)\b", text))):
        if name in ("frs_qdrant", "frs_models", "frs_bert_cfg", "frs_exchange", "frs_sharded", "frs_index", "frs_b200", "frs_encoder"):
            continue
        if name.endswith("_"):   # prefixes such as frs_exchange_* / frs_sharded_*
            continue
        assert hasattr(lib, name), f"INTEGRATION.md names {name}, which the library does not export"
    for cls in ("FrsQdrant", "FrsQdrantMultiGpu", "FrsBert", "ScoredPoint", "Response"):
        assert cls in ns


@pytest.mark.gpu
def test_printed_stubs_work_end_to_end(tmp_path):
    """The stub classes, exactly as printed, against a GPU: upsert / query_points (one GPU and sharded) and the two
    encoders from a Hugging Face style directory."""
    from financial_rag_system_b200 import synth
    from financial_rag_system_b200.checkpoint import BGE_SMALL, MINILM_L6_CE, save_hf_directory, synthetic_checkpoint
    from financial_rag_system_b200.collection import Collection, models
    from financial_rag_system_b200.encoder import Embedder, Reranker
    from financial_rag_system_b200.tokenizer import synthetic_vocab

    ns = _stub_namespace()
    n = 3000
    ids, texts, payloads = synth.make_chunks(n, n_tickers=5, seed=5)
    vecs = np.random.default_rng(0).standard_normal((n, 384)).astype(np.float32)
    pts = [models.PointStruct(id=ids[i], vector=vecs[i].tolist(), payload=payloads[i]) for i in range(n)]
    ours = Collection(n)
    ours.upsert(ids, vecs, payloads)
    flt = models.Filter(must=[models.FieldCondition(key="ticker", match=models.MatchValue(value=payloads[7]["ticker"])),
                              models.FieldCondition(key="document_type", match=models.MatchValue(value=payloads[7]["document_type"]))])
    want_ids, want_scores = ours.search(vecs[7], payloads[7]["ticker"], 15, payloads[7]["document_type"])
    for client in (ns["FrsQdrant"](capacity=n), ns["FrsQdrantMultiGpu"](capacity=n, devices=(0, 0))):
        for s in range(0, n, 256):
            client.upsert("financial_documents", pts[s:s + 256])
        res = client.query_points("financial_documents", query=vecs[7].tolist(), limit=15, query_filter=flt)
        assert [p.id for p in res.points] == [ids[r] for r in want_ids[0] if r >= 0]
        assert np.allclose([p.score for p in res.points], [s for s in want_scores[0] if np.isfinite(s)], atol=1e-6)
        assert res.points[0].id == ids[7] and all("text" in p.payload for p in res.points)

    vocab = synthetic_vocab()
    for shape, seed, has_head, name in ((BGE_SMALL, 1234, False, "bge"), (MINILM_L6_CE, 4321, True, "ce")):
        d = str(tmp_path / name)
        save_hf_directory(d, shape, synthetic_checkpoint(shape, seed), vocab)
        stub = ns["FrsBert"](d, has_head=has_head, max_tokens=8192)
        if has_head:
            pairs = [["What drove iPhone revenue?", t] for t in texts[:5]]
            mine = Reranker(d, device=0, max_tokens=8192)
            assert np.allclose(stub.predict(pairs), mine.predict(pairs), atol=1e-6)
            mine.close()
        else:
            mine = Embedder(d, device=0, max_tokens=8192)
            a, b = stub.encode(texts[:6]), mine.encode(texts[:6])
            assert a.shape == (6, 384) and np.allclose(a, b, atol=1e-6)
            assert stub.encode(texts[0]).shape == (384,)
            mine.close()

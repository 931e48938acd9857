"""Generates tests/golden/search_golden.npz from oracle/search_oracle.py.

The reference cannot run here (qdrant-client is not installed and its tests stub the call, see
oracle/search_oracle.py), so these vectors pin the ORACLE, not the reference: they freeze today's
oracle behaviour so that later edits to it (or to numpy) cannot silently move the target the CUDA
path is compared with.   Run:  python tests/golden/make_search_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import search_oracle as so  # noqa: E402

rng = np.random.default_rng(20261018)
n, nq, k = 512, 8, 15
cent = rng.standard_normal((16, 384)).astype(np.float32)
rows = (cent[rng.integers(0, 16, n)] + 0.6 * rng.standard_normal((n, 384))).astype(np.float32)
rows[100] = rows[7]          # exact duplicates: tie broken by the lower id
rows[300] = rows[7]
ticker = rng.integers(0, 5, n).astype(np.uint32)
doctype = rng.integers(0, 2, n).astype(np.uint32)
codes = (ticker | (doctype << 24)).astype(np.uint32)
codes[100] = codes[7]
codes[300] = codes[7]
codes[11] |= np.uint32(0x80000000)  # a tombstoned row
queries = (rows[[7, 50, 90, 200, 310, 400, 450, 500]] + 0.05 * rng.standard_normal((nq, 384))).astype(np.float32)
q_code = codes[[7, 50, 90, 200, 310, 400, 450, 500]] & np.uint32(0x7FFFFFFF)
q_mask = np.full(nq, 0x80FFFFFF, dtype=np.uint32)   # ticker must match, tombstones never match
q_mask[1] = 0xFFFFFFFF                               # ticker AND document_type
q_mask[2] = 0x80000000                               # no payload filter
out = {"rows": rows, "codes": codes, "queries": queries, "q_code": q_code, "q_mask": q_mask, "k": np.int64(k)}
for dtype in ("bf16", "f32"):
    stored = so.store_rows(rows, dtype)
    qp = so.prepare_queries(queries, dtype)
    ids, sc = so.exact_topk(stored, qp, codes, q_code, q_mask, k)
    out[f"ids_{dtype}"] = ids
    out[f"scores_{dtype}"] = sc
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "search_golden.npz"), **out)
print({k_: (v.shape if hasattr(v, "shape") else v) for k_, v in out.items()})

"""Collection — the chunk store behind the reference's Qdrant calls, with the payload tables the C ABI
keeps out of the GPU library (strings never cross it):

    create_collection(VectorParams(size=384, distance=COSINE))      ingest.py:86-96, database.py:111-143
    upsert(points=[PointStruct(id, vector, payload)])               ingest.py:148-175 (idempotent on id)
    query_points(query=vec, limit=15, query_filter=Filter(must))    main.py:215-239, main2.py:160-163

`Collection` is the native surface (numpy in / numpy out, batched, per-query tickers);
`QdrantCompat` + `models` reproduce the handful of qdrant-client names main.py / ingest.py /
evaluate.py touch, so `get_qdrant()` can return it unchanged.  Vectors live in a `VectorIndex`
(financial_rag_system_b200/index.py -> libfrs_b200.so); this module only maps
ticker / document_type strings to the uint32 payload codes of include/frs_b200.h and row numbers to
point ids / payload dicts.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Any, Optional, Sequence

import numpy as np

from ._lib import CODE_DOCTYPE_SHIFT, CODE_TICKER_MASK, CODE_TOMBSTONE, FRS_DIM, FRS_MAX_BATCH

_DOCTYPE_LIMIT = 127


class Collection:
    def __init__(self, capacity: int, dtype: str = "bf16", device: int = 0, index=None, devices=None):
        """index: an object with VectorIndex's add/set_rows/set_codes/search (tests inject a CPU double).
        devices: a list of CUDA devices -> the rows are sharded over them behind one handle
        (multigpu.MultiGpuIndex, one process driving all GPUs); default: one GPU (`device`)."""
        if index is None and devices is not None and len(devices) > 1:
            from .multigpu import MultiGpuIndex

            index = MultiGpuIndex(capacity, dtype=dtype, devices=devices)
        if index is None:
            from .index import VectorIndex

            index = VectorIndex(capacity, dtype=dtype, device=devices[0] if devices else device)
        self.index = index
        self.capacity = int(capacity)
        self._tickers: dict[str, int] = {}
        self._doctypes: dict[str, int] = {}
        self._row_of_id: dict[Any, int] = {}
        self.ids: list[Any] = []
        self.payloads: list[dict] = []
        self._codes = np.zeros(0, dtype=np.uint32)
        self._lock = threading.Lock()
        # ticker-segmented search: the 128-row tiles that hold rows of each ticker.  Ingest is
        # per ticker (ingest.py:109-177), so these sets are small and a filtered batch scans only them.
        self.segmented = hasattr(index, "search_tiles")
        self._ticker_tiles: dict[int, np.ndarray] = {}
        self.last_scan_tiles = None  # (tiles scanned, tiles in the index) of the most recent search

    def __len__(self) -> int:
        return len(self.ids)

    # -- payload codes ----------------------------------------------------------------------------
    def _ticker_code(self, t: str, create: bool) -> Optional[int]:
        t = str(t)
        c = self._tickers.get(t)
        if c is None and create:
            c = len(self._tickers) + 1  # 0 = "no ticker"
            if c > CODE_TICKER_MASK:
                raise ValueError("too many distinct tickers")
            self._tickers[t] = c
        return c

    def _doctype_code(self, d: str, create: bool) -> Optional[int]:
        d = str(d)
        c = self._doctypes.get(d)
        if c is None and create:
            c = len(self._doctypes) + 1
            if c > _DOCTYPE_LIMIT:
                raise ValueError("too many distinct document types")
            self._doctypes[d] = c
        return c

    def _row_code(self, payload: dict) -> int:
        code = 0
        if payload.get("ticker") is not None:
            code |= self._ticker_code(payload["ticker"], True)
        if payload.get("document_type") is not None:
            code |= self._doctype_code(payload["document_type"], True) << CODE_DOCTYPE_SHIFT
        return code

    def predicate(self, ticker: Optional[str], document_type: Optional[str] = None) -> tuple[int, int]:
        """(code, mask) of `Filter(must=[ticker == T, document_type == D])`: a row matches iff
        ((row_code ^ code) & mask) == 0.  An unknown keyword can match nothing: the tombstone bit is
        demanded set, which no live row has."""
        code, mask = 0, CODE_TOMBSTONE
        if ticker is not None:
            c = self._ticker_code(ticker, False)
            if c is None:
                return CODE_TOMBSTONE, CODE_TOMBSTONE
            code |= c
            mask |= CODE_TICKER_MASK
        if document_type is not None:
            c = self._doctype_code(document_type, False)
            if c is None:
                return CODE_TOMBSTONE, CODE_TOMBSTONE
            code |= c << CODE_DOCTYPE_SHIFT
            mask |= _DOCTYPE_LIMIT << CODE_DOCTYPE_SHIFT
        return code, mask

    # -- write path -------------------------------------------------------------------------------
    def upsert(self, ids: Sequence[Any], vectors, payloads: Sequence[dict]) -> None:
        """qdrant.upsert: a point whose id already exists is overwritten in place (same row), new ids
        are appended.  Within one call the last occurrence of a repeated id wins."""
        vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        if vectors.ndim != 2 or vectors.shape[1] != FRS_DIM or len(ids) != len(vectors) or len(ids) != len(payloads):
            raise ValueError("ids, vectors [n,384] and payloads must have the same length")
        with self._lock:
            last = {pid: i for i, pid in enumerate(ids)}
            new_i, upd = [], []
            for i, pid in enumerate(ids):
                if last[pid] != i:
                    continue
                row = self._row_of_id.get(pid)
                (upd if row is not None else new_i).append((i, row))
            if len(self.ids) + len(new_i) > self.capacity:
                raise ValueError(f"collection full: {len(self.ids)} + {len(new_i)} > {self.capacity}")
            if new_i:
                sel = [i for i, _ in new_i]
                codes = np.array([self._row_code(payloads[i]) for i in sel], dtype=np.uint32)
                self.index.add(vectors[sel], codes)
                for i in sel:
                    self._row_of_id[ids[i]] = len(self.ids)
                    self.ids.append(ids[i])
                    self.payloads.append(dict(payloads[i]))
                self._note_tiles(len(self._codes), codes)
                self._codes = np.concatenate([self._codes, codes])
            for i, row in upd:
                code = np.array([self._row_code(payloads[i])], dtype=np.uint32)
                self._set_row(row, vectors[i:i + 1], code)
                self.payloads[row] = dict(payloads[i])
                self._codes[row] = code[0]
                self._note_tiles(row, code)

    def _note_tiles(self, row0: int, codes: np.ndarray) -> None:
        """Record which 128-row tiles now hold rows of which ticker (sets only ever grow: a stale entry
        costs a little bandwidth, never correctness — rows are still filtered by the predicate)."""
        tick = (codes & CODE_TICKER_MASK).astype(np.int64)
        tile = (row0 + np.arange(len(codes), dtype=np.int64)) // 128
        pairs = np.unique((tick << 32) | tile)
        for t in np.unique(pairs >> 32):
            new = (pairs[(pairs >> 32) == t] & 0xFFFFFFFF).astype(np.int64)
            old = self._ticker_tiles.get(int(t))
            self._ticker_tiles[int(t)] = new if old is None else np.union1d(old, new)

    def _set_row(self, row: int, vec: np.ndarray, code: np.ndarray) -> None:
        self.index.set_rows(row, vec, code)

    def delete(self, ids: Sequence[Any]) -> None:
        """Tombstone the rows of these ids (they stop matching any query)."""
        with self._lock:
            for pid in ids:
                row = self._row_of_id.pop(pid, None)
                if row is None:
                    continue
                self._codes[row] |= CODE_TOMBSTONE
                self.index.set_codes(row, self._codes[row:row + 1])

    # -- read path --------------------------------------------------------------------------------
    def search(self, query_vecs, ticker, limit: int = 15, document_type=None):
        """Exact cosine top-`limit` per query among rows whose payload matches.  query_vecs [B,384]
        (or [384]); ticker / document_type: one value for all queries or one per query.  Returns
        (ids int64 [B,limit] row numbers, -1 padded; scores float32 [B,limit], -inf padded)."""
        q = np.ascontiguousarray(query_vecs, dtype=np.float32)
        if q.ndim == 1:
            q = q[None]
        B = q.shape[0]
        tick = [ticker] * B if (ticker is None or isinstance(ticker, str)) else list(ticker)
        doc = [document_type] * B if (document_type is None or isinstance(document_type, str)) else list(document_type)
        if len(tick) != B or len(doc) != B:
            raise ValueError("one ticker / document_type per query")
        pred = [self.predicate(t, d) for t, d in zip(tick, doc)]
        code = np.array([p[0] for p in pred], dtype=np.uint32)
        mask = np.array([p[1] for p in pred], dtype=np.uint32)
        ids = np.empty((B, limit), dtype=np.int64)
        scores = np.empty((B, limit), dtype=np.float32)
        total_tiles = (len(self._codes) + 127) // 128
        for s in range(0, B, FRS_MAX_BATCH):
            e = min(B, s + FRS_MAX_BATCH)
            tiles = self._batch_tiles(code[s:e], mask[s:e], total_tiles) if self.segmented else None
            if tiles is None:
                self.last_scan_tiles = (total_tiles, total_tiles)
                ids[s:e], scores[s:e] = self.index.search(q[s:e], code[s:e], mask[s:e], limit)
            else:
                self.last_scan_tiles = (len(tiles), total_tiles)
                i_, s_ = self.index.search_tiles(q[s:e], code[s:e], mask[s:e], limit, tiles)
                ids[s:e], scores[s:e] = i_.cpu().numpy(), s_.cpu().numpy()
        return ids, scores

    def _batch_tiles(self, code: np.ndarray, mask: np.ndarray, total_tiles: int):
        """Tiles a filtered batch has to read: the union of its tickers' tile sets; None = full scan (a
        query without ticker condition, or a union that is most of the index anyway)."""
        sets = []
        for c, m in zip(code, mask):
            if int(c) & CODE_TOMBSTONE:      # unknown keyword: matches nothing, needs no tile
                continue
            if (int(m) & CODE_TICKER_MASK) != CODE_TICKER_MASK:
                return None
            t = self._ticker_tiles.get(int(c) & CODE_TICKER_MASK)
            if t is not None:
                sets.append(t)
        tiles = np.unique(np.concatenate(sets)) if sets else np.zeros(0, dtype=np.int64)
        return None if len(tiles) > 0.75 * total_tiles else tiles

    # -- persistence (the Qdrant volume of docker-compose.yml:26-27) ----------------------------------
    def save(self, path: str, chunk_rows: int = 1 << 20) -> None:
        """Write the collection to a directory: `rows.bin` (storage-dtype rows as they sit in HBM),
        `codes.npy`, `points.jsonl` (id + payload per row) and `meta.json` (dtype, dictionaries)."""
        import json
        import os

        os.makedirs(path, exist_ok=True)
        with self._lock:
            n = len(self.ids)
            codes = np.empty(n, dtype=np.uint32)
            with open(os.path.join(path, "rows.bin"), "wb") as f:
                for s in range(0, n, chunk_rows):
                    rows, c = self.index.export_raw(s, min(chunk_rows, n - s))
                    rows.tofile(f)
                    codes[s:s + len(c)] = c
            np.save(os.path.join(path, "codes.npy"), codes)
            with open(os.path.join(path, "points.jsonl"), "w") as f:
                for pid, payload in zip(self.ids, self.payloads):
                    f.write(json.dumps({"id": pid, "payload": payload}) + "\n")
            with open(os.path.join(path, "meta.json"), "w") as f:
                json.dump({"format": 1, "dtype": getattr(self.index, "dtype", "bf16"), "rows": n, "dim": FRS_DIM,
                           "tickers": self._tickers, "doctypes": self._doctypes}, f)

    @classmethod
    def load(cls, path: str, capacity: Optional[int] = None, device: int = 0, index=None, chunk_rows: int = 1 << 20,
             devices=None):
        """Rebuild a collection saved by `save`; every query then returns bit-identical ids and scores."""
        import json
        import os

        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        n = int(meta["rows"])
        c = cls(max(int(capacity or 0), n, 1), dtype=meta["dtype"], device=device, index=index, devices=devices)
        codes = np.load(os.path.join(path, "codes.npy"))
        esz, dt = (4, np.float32) if meta["dtype"] == "f32" else (2, np.uint16)
        with open(os.path.join(path, "rows.bin"), "rb") as f:
            for s in range(0, n, chunk_rows):
                m = min(chunk_rows, n - s)
                rows = np.frombuffer(f.read(m * FRS_DIM * esz), dtype=dt).reshape(m, FRS_DIM)
                c.index.import_raw(rows, codes[s:s + m])
        with open(os.path.join(path, "points.jsonl")) as f:
            for row, line in enumerate(f):
                rec = json.loads(line)
                pid = rec["id"]
                c.ids.append(pid)
                c.payloads.append(rec["payload"])
                if not int(codes[row]) & CODE_TOMBSTONE:
                    c._row_of_id[pid] = row
        c._codes = codes.astype(np.uint32)
        for s in range(0, n, 1 << 20):
            c._note_tiles(s, c._codes[s:s + (1 << 20)])
        c._tickers = {k: int(v) for k, v in meta["tickers"].items()}
        c._doctypes = {k: int(v) for k, v in meta["doctypes"].items()}
        return c

    def close(self) -> None:
        if hasattr(self.index, "close"):
            self.index.close()


# ------------------------------------------------------------------------------------------------
# qdrant-client look-alikes: only what main.py / main2.py / ingest.py / database.py / evaluate.py use
# ------------------------------------------------------------------------------------------------
class models:  # noqa: N801 - mirrors `from qdrant_client.http import models`
    @dataclass
    class MatchValue:
        value: Any

    @dataclass
    class FieldCondition:
        key: str
        match: "models.MatchValue"

    @dataclass
    class Filter:
        must: list = field(default_factory=list)

    @dataclass
    class PointStruct:
        id: Any
        vector: Any
        payload: dict = field(default_factory=dict)

    class Distance:
        COSINE = "Cosine"

    @dataclass
    class VectorParams:
        size: int
        distance: str = "Cosine"


@dataclass
class ScoredPoint:
    id: Any
    score: float
    payload: dict


@dataclass
class QueryResponse:
    points: list


@dataclass
class _CollectionDescription:
    name: str


@dataclass
class _CollectionsResponse:
    collections: list


class QdrantCompat:
    """Duck-typed stand-in for `QdrantClient` (main.py:92-95 get_qdrant)."""

    def __init__(self, capacity: int = 1_000_000, dtype: str = "bf16", device: int = 0, index_factory=None, devices=None):
        """devices=[0, 1, ...]: every collection is sharded over these GPUs behind the one client object, the
        way the reference's single server process holds one QdrantClient (main2.py:104-108)."""
        self._capacity, self._dtype, self._device, self._factory = capacity, dtype, device, index_factory
        self._devices = list(devices) if devices is not None else None
        self._collections: dict[str, Collection] = {}

    def collection_exists(self, collection_name: str) -> bool:
        return collection_name in self._collections

    def get_collections(self):
        return _CollectionsResponse([_CollectionDescription(n) for n in self._collections])

    def create_collection(self, collection_name: str, vectors_config=None, **_):
        size = getattr(vectors_config, "size", FRS_DIM)
        dist = getattr(vectors_config, "distance", models.Distance.COSINE)
        if size != FRS_DIM or str(dist).lower() not in ("cosine", "distance.cosine"):
            raise ValueError("only VectorParams(size=384, distance=COSINE) collections are supported")
        idx = self._factory(self._capacity) if self._factory else None
        self._collections[collection_name] = Collection(self._capacity, self._dtype, self._device, index=idx,
                                                        devices=self._devices)
        return True

    def collection(self, collection_name: str) -> Collection:
        return self._collections[collection_name]

    def upsert(self, collection_name: str, points, **_):
        c = self._collections[collection_name]
        c.upsert([p.id for p in points], np.asarray([p.vector for p in points], dtype=np.float32),
                 [p.payload or {} for p in points])

    def query_points(self, collection_name: str, query, limit: int = 10, query_filter=None, **_):
        c = self._collections[collection_name]
        ticker = doc = None
        for cond in (getattr(query_filter, "must", None) or []):
            if cond.key == "ticker":
                ticker = cond.match.value
            elif cond.key == "document_type":
                doc = cond.match.value
            else:
                raise ValueError(f"unsupported filter key {cond.key!r} (the reference filters on ticker / document_type)")
        ids, scores = c.search(np.asarray(query, dtype=np.float32), ticker, limit, doc)
        pts = [ScoredPoint(c.ids[r], float(s), c.payloads[r]) for r, s in zip(ids[0], scores[0]) if r >= 0]
        return QueryResponse(pts)

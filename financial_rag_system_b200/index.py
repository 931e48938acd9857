"""VectorIndex — thin Python handle on one GPU shard of the chunk store (C ABI: include/frs_b200.h).

PyTorch is used for device memory and streams only; every kernel launched here is from
libfrs_b200.so.  Replaces the Qdrant collection of the reference (create_collection
ingest.py:86-96, upsert ingest.py:171-175, query_points main.py:232-237).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FRS_DIM, FRS_DTYPE_BF16, FRS_DTYPE_F32, FRS_MAX_BATCH, FRS_MAX_K, check

_DTYPES = {"bf16": FRS_DTYPE_BF16, "f32": FRS_DTYPE_F32, "fp32": FRS_DTYPE_F32}


def _ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class PendingSearch:
    """Result of a pipelined search: `ids` / `scores` are valid once the ticket has been waited for."""

    def __init__(self, index, ticket: int, ids, scores, keep=()):
        self._index, self.ticket, self.ids, self.scores, self._keep = index, ticket, ids, scores, keep

    def wait(self):
        """Block the host until the results are complete; returns (ids, scores)."""
        if self._index is not None:
            self._index.sync(self.ticket)
            self._index = None
        self._keep = ()
        return self.ids, self.scores

    def wait_stream(self):
        """Make the current CUDA stream wait for the results (no host block); returns (ids, scores)."""
        if self._index is not None:
            self._index.wait(self.ticket)
        return self.ids, self.scores


class VectorIndex:
    """One shard: `capacity` rows x 384, stored L2-normalised in bf16 or fp32 on one GPU."""

    def __init__(self, capacity: int, dtype: str = "bf16", device: int = 0, base: int = 0):
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self._lib = _lib.lib()
        self.dtype = "f32" if _DTYPES[dtype] == FRS_DTYPE_F32 else "bf16"
        self.device = torch.device("cuda", device)
        self.capacity = int(capacity)
        self.base = int(base)
        h = C.c_void_p()
        check(self._lib.frs_index_create(device, FRS_DIM, self.capacity, _DTYPES[dtype], C.byref(h)))
        self._h = h
        check(self._lib.frs_index_set_base(self._h, self.base))

    # -- lifetime -------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.frs_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._lib.frs_index_size(self._h))

    # -- write path -----------------------------------------------------------------------------
    def add(self, vecs, codes=None) -> None:
        """Append rows.  vecs [n,384] float32 (torch cuda tensor or numpy), codes [n] uint32/int32."""
        if isinstance(vecs, torch.Tensor):
            v = vecs.to(device=self.device, dtype=torch.float32).contiguous()
            if v.dim() != 2 or v.shape[1] != FRS_DIM:
                raise ValueError("vecs must be [n, 384]")
            c = None
            if codes is not None:
                c = torch.as_tensor(codes).to(device=self.device).contiguous()
                if c.dtype not in (torch.int32, torch.uint32):
                    c = c.to(torch.int64).to(torch.int32)
                if c.numel() != v.shape[0]:
                    raise ValueError("codes must be [n]")
            check(self._lib.frs_index_add(self._h, _ptr(v), _ptr(c), v.shape[0], _stream_ptr(self.device)))
            # keep v/c alive until the kernel has consumed them
            torch.cuda.current_stream(self.device).synchronize()
        else:
            v = np.ascontiguousarray(vecs, dtype=np.float32)
            if v.ndim != 2 or v.shape[1] != FRS_DIM:
                raise ValueError("vecs must be [n, 384]")
            c = None if codes is None else np.ascontiguousarray(codes).astype(np.uint32)
            check(self._lib.frs_index_add_host(self._h, _ptr(v), _ptr(c), v.shape[0]))

    def set_rows(self, row0: int, vecs, codes=None) -> None:
        """Overwrite rows in place (torch tensors or numpy); codes=None keeps the stored payload codes."""
        v = torch.as_tensor(vecs).to(device=self.device, dtype=torch.float32).contiguous()
        c = None if codes is None else self._as_code_tensor(codes)
        check(self._lib.frs_index_set_rows(self._h, int(row0), _ptr(v), _ptr(c), v.shape[0], _stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()

    def set_codes(self, row0: int, codes) -> None:
        c = self._as_code_tensor(codes)
        check(self._lib.frs_index_set_codes(self._h, int(row0), _ptr(c), c.numel(), _stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()

    def read_rows(self, row0: int = 0, n: int | None = None) -> torch.Tensor:
        n = len(self) - row0 if n is None else n
        out = torch.empty((n, FRS_DIM), dtype=torch.float32, device=self.device)
        check(self._lib.frs_index_read_rows(self._h, int(row0), int(n), _ptr(out), _stream_ptr(self.device)))
        return out

    # -- persistence ----------------------------------------------------------------------------
    def export_raw(self, row0: int = 0, n: int | None = None):
        """Stored rows exactly as they sit in HBM (uint16 bf16 bit patterns or float32) + payload codes."""
        n = len(self) - row0 if n is None else n
        rows = np.empty((n, FRS_DIM), dtype=np.float32 if self.dtype == "f32" else np.uint16)
        codes = np.empty((n,), dtype=np.uint32)
        check(self._lib.frs_index_export_raw(self._h, int(row0), int(n), _ptr(rows), _ptr(codes)))
        return rows, codes

    def import_raw(self, rows: np.ndarray, codes: np.ndarray) -> None:
        """Append rows previously produced by export_raw (no normalisation, no rounding)."""
        want = np.float32 if self.dtype == "f32" else np.uint16
        rows = np.ascontiguousarray(rows)
        if rows.dtype != want or rows.ndim != 2 or rows.shape[1] != FRS_DIM:
            raise ValueError(f"raw rows must be [n, 384] {np.dtype(want).name} for a {self.dtype} index")
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        check(self._lib.frs_index_import_raw(self._h, _ptr(rows), _ptr(codes), rows.shape[0]))

    def set_scan_grid(self, grid: int) -> None:
        check(self._lib.frs_index_set_scan_grid(self._h, int(grid)))

    def set_pipeline_reserve(self, sms: int) -> None:
        """SMs the persistent scan kernel leaves to the prep / merge / exchange kernels in the pipelined forms."""
        check(self._lib.frs_index_set_pipeline_reserve(self._h, int(sms)))

    def set_scan_streams(self, n: int) -> None:
        check(self._lib.frs_index_set_scan_streams(self._h, int(n)))

    # -- search ---------------------------------------------------------------------------------
    @staticmethod
    def _check_batch(nq: int, k: int) -> None:
        if not 1 <= nq <= FRS_MAX_BATCH:
            raise ValueError(f"1..{FRS_MAX_BATCH} queries per call (got {nq})")
        if not 1 <= k <= FRS_MAX_K:
            raise ValueError(f"k must be in 1..{FRS_MAX_K} (got {k})")

    def search(self, queries, q_code, q_mask, k: int = 15):
        """Exact cosine top-k.  Device tensors in -> device tensors out (async on the current
        stream); numpy in -> numpy out through the host entry point (copies inside the call).
        Returns (ids int64 [nq,k], scores float32 [nq,k])."""
        if isinstance(queries, torch.Tensor):
            q = queries.to(device=self.device, dtype=torch.float32).contiguous()
            nq = q.shape[0]
            self._check_batch(nq, k)
            qc = torch.as_tensor(q_code).to(device=self.device, dtype=torch.int32).contiguous()
            qm = torch.as_tensor(q_mask).to(device=self.device, dtype=torch.int32).contiguous()
            scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            check(self._lib.frs_index_search(self._h, _ptr(q), _ptr(qc), _ptr(qm), nq, k, _ptr(scores), _ptr(ids),
                                             _stream_ptr(self.device)))
            # q/qc/qm are consumed by kernels already enqueued on this stream; record them so the
            # caching allocator does not hand the memory to another stream early
            for t in (q, qc, qm):
                t.record_stream(torch.cuda.current_stream(self.device))
            return ids, scores
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        self._check_batch(nq, k)
        qc = np.ascontiguousarray(np.asarray(q_code, dtype=np.int64).astype(np.uint32))
        qm = np.ascontiguousarray(np.asarray(q_mask, dtype=np.int64).astype(np.uint32))
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        check(self._lib.frs_index_search_host(self._h, _ptr(q), _ptr(qc), _ptr(qm), nq, k, _ptr(scores), _ptr(ids)))
        return ids, scores

    # -- pipelined forms (frs_index_search_async / _host_submit / _host_collect) ------------------
    def search_async(self, queries: torch.Tensor, q_code: torch.Tensor, q_mask: torch.Tensor, k: int = 15,
                     exchange=None, out=None) -> PendingSearch:
        """Pipelined search of device tensors: prep / scan / merge run on the index's internal streams, so
        consecutive calls overlap (the scans run back to back).  Inputs are read in current-stream order.
        `exchange`: a sharded.PeerExchange — the result is then the global top-k over all ranks.
        `out`: optional (ids int64 [nq,k], scores float32 [nq,k]) device tensors to write into (a serving loop that
        recycles its result buffers keeps the allocator out of the request path)."""
        q = queries
        if q.dtype != torch.float32 or q.device != self.device or not q.is_contiguous():
            q = q.to(device=self.device, dtype=torch.float32).contiguous()
        nq = q.shape[0]
        self._check_batch(nq, k)
        qc = q_code if (q_code.dtype == torch.int32 and q_code.device == self.device) else self._as_code_tensor(q_code)
        qm = q_mask if (q_mask.dtype == torch.int32 and q_mask.device == self.device) else self._as_code_tensor(q_mask)
        if out is not None:
            ids, scores = out
            assert ids.dtype == torch.int64 and scores.dtype == torch.float32 and tuple(ids.shape) == tuple(scores.shape) == (nq, k)
            assert ids.is_contiguous() and scores.is_contiguous() and ids.device == scores.device == self.device
        else:
            # one allocation for both outputs: [nq*k] int64 ids followed by [nq*k] float32 scores
            buf = torch.empty(nq * k * 12, dtype=torch.uint8, device=self.device)
            ids = buf[:nq * k * 8].view(torch.int64).view(nq, k)
            scores = buf[nq * k * 8:].view(torch.float32).view(nq, k)
        ticket = C.c_int(-1)
        check(self._lib.frs_index_search_async(self._h, exchange._h if exchange is not None else None, _ptr(q), _ptr(qc),
                                               _ptr(qm), nq, k, _ptr(scores), _ptr(ids), _stream_ptr(self.device),
                                               C.byref(ticket)))
        return PendingSearch(self, ticket.value, ids, scores, keep=(q, qc, qm))

    def wait(self, ticket: int = -1) -> None:
        """The current stream waits for a pipelined search (-1: everything submitted so far)."""
        check(self._lib.frs_index_wait(self._h, int(ticket), _stream_ptr(self.device)))

    def sync(self, ticket: int = -1) -> None:
        check(self._lib.frs_index_sync(self._h, int(ticket)))

    def submit_host(self, queries: np.ndarray, q_code: np.ndarray, q_mask: np.ndarray, k: int = 15, exchange=None) -> int:
        """Host buffers in; returns a ticket for collect_host.  Up to 4 batches in flight per index."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        self._check_batch(nq, k)
        qc = np.ascontiguousarray(np.asarray(q_code, dtype=np.int64).astype(np.uint32))
        qm = np.ascontiguousarray(np.asarray(q_mask, dtype=np.int64).astype(np.uint32))
        ticket = C.c_int(-1)
        check(self._lib.frs_index_search_host_submit(self._h, exchange._h if exchange is not None else None, _ptr(q),
                                                     _ptr(qc), _ptr(qm), nq, k, C.byref(ticket)))
        self._pending_shape = getattr(self, "_pending_shape", {})
        self._pending_shape[ticket.value] = (nq, k)
        return ticket.value

    def collect_host(self, ticket: int, exchange=None):
        nq, k = self._pending_shape.pop(ticket)
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        check(self._lib.frs_index_search_host_collect(self._h, exchange._h if exchange is not None else None, int(ticket),
                                                      _ptr(scores), _ptr(ids)))
        return ids, scores

    def search_tiles(self, queries, q_code, q_mask, k: int, tile_ids):
        """Exact search restricted to the listed 128-row tiles (ascending unique int32/uint32 array;
        numpy or torch).  Queries may be numpy (copied) or device tensors; returns device tensors."""
        q = torch.as_tensor(queries).to(device=self.device, dtype=torch.float32).contiguous()
        nq = q.shape[0]
        self._check_batch(nq, k)
        qc = self._as_code_tensor(q_code)
        qm = self._as_code_tensor(q_mask)
        if isinstance(tile_ids, torch.Tensor):
            t = tile_ids.to(device=self.device, dtype=torch.int32).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(tile_ids, dtype=np.int64).astype(np.int32)).to(self.device)
        scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        check(self._lib.frs_index_search_tiles(self._h, _ptr(q), _ptr(qc), _ptr(qm), nq, k, _ptr(t), int(t.numel()),
                                               _ptr(scores), _ptr(ids), _stream_ptr(self.device)))
        for x in (q, qc, qm, t):
            x.record_stream(torch.cuda.current_stream(self.device))
        return ids, scores

    def _as_code_tensor(self, c) -> torch.Tensor:
        """uint32 payload words (numpy or torch, any integer dtype) as the int32 bit pattern the ABI reads."""
        if isinstance(c, torch.Tensor) and c.dtype == torch.int32:
            return c.to(self.device).contiguous()
        c = torch.as_tensor(np.asarray(c.cpu() if isinstance(c, torch.Tensor) else c, dtype=np.int64) & 0xFFFFFFFF)
        c = torch.where(c >= (1 << 31), c - (1 << 32), c)
        return c.to(torch.int32).to(self.device).contiguous()

    def search_local(self, q: torch.Tensor, qc: torch.Tensor, qm: torch.Tensor, k: int,
                     out_scores64: torch.Tensor, out_ids: torch.Tensor) -> None:
        """Shard-local pass of a sharded search: exact local top-k as (float64, int64 global id)."""
        self._check_batch(q.shape[0], k)
        check(self._lib.frs_index_search_local(self._h, _ptr(q), _ptr(qc), _ptr(qm), q.shape[0], k,
                                               _ptr(out_scores64), _ptr(out_ids), _stream_ptr(self.device)))

    def search_push(self, q: torch.Tensor, qc: torch.Tensor, qm: torch.Tensor, k: int, exchange) -> None:
        """Shard-local pass whose merge kernel also pushes the result into every peer's gather buffer
        (frs_index_search_push); `exchange` is a sharded.PeerExchange created for (len(q), k)."""
        self._check_batch(q.shape[0], k)
        check(self._lib.frs_index_search_push(self._h, _ptr(q), _ptr(qc), _ptr(qm), q.shape[0], k, exchange._h,
                                              _stream_ptr(self.device)))

    def last_queries(self) -> torch.Tensor:
        out = torch.empty((FRS_MAX_BATCH, FRS_DIM), dtype=torch.float32, device=self.device)
        check(self._lib.frs_index_last_queries(self._h, _ptr(out), _stream_ptr(self.device)))
        return out

    def debug_scores(self, queries: torch.Tensor) -> torch.Tensor:
        """Raw tensor-core pre-filter scores [32, n] (diagnostics)."""
        q = queries.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.zeros((FRS_MAX_BATCH, len(self)), dtype=torch.float32, device=self.device)
        check(self._lib.frs_index_debug_scores(self._h, _ptr(q), q.shape[0], _ptr(out), _stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()
        return out

    def set_profiling(self, mode: int) -> None:
        check(self._lib.frs_index_set_profiling(self._h, int(mode)))

    def read_profile(self) -> dict:
        """{'n', 'prep_ms', 'scan_ms', 'merge_ms'}: per-kernel CUDA-event time summed over the
        searches recorded since the last call (profiling mode >= 1)."""
        buf = (C.c_double * 4)()
        check(self._lib.frs_index_read_profile(self._h, buf))
        return {"n": int(buf[0]), "prep_ms": buf[1], "scan_ms": buf[2], "merge_ms": buf[3]}

    def read_profile_ex(self) -> dict:
        """read_profile plus 'exchange_ms' (cross-shard wait + merge), 'scan_gap_ms' (idle time of the scan stream
        between consecutive scan kernels) and 'span_ms' (first prep start to last search end)."""
        buf = (C.c_double * 8)()
        check(self._lib.frs_index_read_profile_ex(self._h, buf))
        keys = ("n", "prep_ms", "scan_ms", "merge_ms", "exchange_ms", "scan_gap_ms", "span_ms")
        d = dict(zip(keys, [float(x) for x in buf]))
        d["n"] = int(d["n"])
        return d

    def read_profile_bracket_rel(self, ev_before, ev_after) -> tuple[float, float]:
        """Bracket mode (3): (ms from `ev_before` to the first scan kernel's start, ms from the last scan kernel's end to
        `ev_after`) for two recorded torch.cuda.Event objects — fill and drain of a pipelined run.  Call before
        read_profile_ex (which resets the bracket)."""
        buf = (C.c_double * 2)()
        check(self._lib.frs_index_read_profile_bracket_rel(self._h, C.c_void_p(ev_before.cuda_event), C.c_void_p(ev_after.cuda_event), buf))
        return float(buf[0]), float(buf[1])

    def read_profile_raw(self, max_searches: int = 256) -> np.ndarray:
        """[n, 7] ms: prep start / end, scan start / end, merge start / end, exchange end of the recorded searches
        (oldest first), relative to the first one's prep start.  Call before read_profile[_ex] (which resets)."""
        buf = (C.c_double * (7 * max_searches))()
        n = self._lib.frs_index_read_profile_raw(self._h, buf, int(max_searches))
        if n < 0:
            check(n)
        return np.array(buf[:7 * n], dtype=np.float64).reshape(n, 7)

    def read_timeline(self, n_ctas: int) -> np.ndarray:
        out = np.zeros((n_ctas, 16), dtype=np.uint64)
        check(self._lib.frs_index_read_timeline(self._h, _ptr(out), n_ctas))
        return out

    def last_stats(self) -> dict:
        buf = (C.c_int64 * 6)()
        check(self._lib.frs_index_last_stats(self._h, buf))
        keys = ("appended", "compactions", "resolutions", "rescored", "grid", "launches")
        return dict(zip(keys, [int(x) for x in buf]))


def merge_shards(scores64: torch.Tensor, ids: torch.Tensor, k: int):
    """[n_shards, nq, k] exact candidates -> ([nq,k] ids, [nq,k] float32 scores), on device."""
    n_shards, nq, kk = scores64.shape
    assert kk == k and ids.shape == scores64.shape
    dev = scores64.device
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    check(_lib.lib().frs_merge_shards(dev.index or 0, _ptr(scores64.contiguous()), _ptr(ids.contiguous()), n_shards, nq, k,
                                      _ptr(out_s), _ptr(out_i), _stream_ptr(dev)))
    return out_i, out_s


def merge_shards_packed(packed: torch.Tensor, k: int):
    """[n_shards, 2, nq, k] int64 exchange buffer (plane 0 fp64 score bits, plane 1 ids) -> merged
    ([nq,k] ids, [nq,k] float32 scores), on device."""
    n_shards, two, nq, kk = packed.shape
    assert two == 2 and kk == k and packed.dtype == torch.int64 and packed.is_contiguous()
    dev = packed.device
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    check(_lib.lib().frs_merge_shards_packed(dev.index or 0, _ptr(packed), n_shards, nq, k, _ptr(out_s), _ptr(out_i),
                                             _stream_ptr(dev)))
    return out_i, out_s

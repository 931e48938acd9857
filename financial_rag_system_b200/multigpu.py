"""MultiGpuIndex — ONE process, several GPUs: Python handle on `frs_sharded` (include/frs_b200.h,
csrc/sharded.cu).  This is the form the reference's server process would hold: `get_qdrant()` returns one
client object (main.py:92-95, main2.py:104-108) that request threads call concurrently
(main.py:215-239, main2.py:160-163).  Same surface as `VectorIndex` where `Collection` needs it
(add / set_rows / set_codes / search / export_raw / import_raw), ids are global row numbers in insertion
order, results are identical to a single `VectorIndex` over all rows.

(`sharded.ShardedIndex` is the other multi-GPU form: one process per GPU under torch.distributed.)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FRS_DIM, FRS_DTYPE_BF16, FRS_DTYPE_F32, FRS_MAX_BATCH, FRS_MAX_K, check

_DTYPES = {"bf16": FRS_DTYPE_BF16, "f32": FRS_DTYPE_F32, "fp32": FRS_DTYPE_F32}


def _np_ptr(a: Optional[np.ndarray]) -> C.c_void_p:
    return C.c_void_p(0) if a is None else C.c_void_p(a.ctypes.data)


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray((np.asarray(a, dtype=np.int64) & 0xFFFFFFFF).astype(np.uint32))


class MultiGpuIndex:
    def __init__(self, capacity: int, dtype: str = "bf16", devices: Optional[Sequence[int]] = None):
        """devices: CUDA device numbers, one shard each (default: every visible device)."""
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self._lib = _lib.lib()
        if devices is None:
            n = self._lib.frs_device_count()
            if n < 1:
                check(n if n < 0 else -2)
            devices = list(range(n))
        self.devices = [int(d) for d in devices]
        self.dtype = "f32" if _DTYPES[dtype] == FRS_DTYPE_F32 else "bf16"
        self.capacity = int(capacity)
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        check(self._lib.frs_sharded_create(len(self.devices), arr, FRS_DIM, self.capacity, _DTYPES[dtype], C.byref(h)))
        self._h = h
        self.block_rows = int(self._lib.frs_sharded_block_rows(self._h))
        self._shape = {}

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.frs_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._lib.frs_sharded_size(self._h))

    @property
    def n_shards(self) -> int:
        return len(self.devices)

    def placement(self, row: int) -> tuple[int, int]:
        """(shard, local row) of a global row: block-cyclic, see csrc/sharded.cu."""
        blk, r = divmod(int(row), self.block_rows)
        return blk % self.n_shards, (blk // self.n_shards) * self.block_rows + r

    # -- write path -------------------------------------------------------------------------------
    def add(self, vecs, codes=None) -> None:
        v = np.ascontiguousarray(_to_numpy(vecs), dtype=np.float32)
        if v.ndim != 2 or v.shape[1] != FRS_DIM:
            raise ValueError("vecs must be [n, 384]")
        c = None if codes is None else _u32(_to_numpy(codes))
        if c is not None and c.shape != (v.shape[0],):
            raise ValueError("codes must be [n]")
        check(self._lib.frs_sharded_add_host(self._h, _np_ptr(v), _np_ptr(c), v.shape[0]))

    def set_rows(self, row0: int, vecs, codes=None) -> None:
        v = np.ascontiguousarray(_to_numpy(vecs), dtype=np.float32)
        c = None if codes is None else _u32(_to_numpy(codes))
        check(self._lib.frs_sharded_set_rows_host(self._h, int(row0), _np_ptr(v), _np_ptr(c), v.shape[0]))

    def set_codes(self, row0: int, codes) -> None:
        c = _u32(_to_numpy(codes))
        check(self._lib.frs_sharded_set_rows_host(self._h, int(row0), None, _np_ptr(c), c.shape[0]))

    def read_rows(self, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = len(self) - row0 if n is None else n
        out = np.empty((n, FRS_DIM), dtype=np.float32)
        check(self._lib.frs_sharded_read_rows_host(self._h, int(row0), int(n), _np_ptr(out)))
        return out

    def export_raw(self, row0: int = 0, n: Optional[int] = None):
        n = len(self) - row0 if n is None else n
        rows = np.empty((n, FRS_DIM), dtype=np.float32 if self.dtype == "f32" else np.uint16)
        codes = np.empty((n,), dtype=np.uint32)
        check(self._lib.frs_sharded_export_raw(self._h, int(row0), int(n), _np_ptr(rows), _np_ptr(codes)))
        return rows, codes

    def import_raw(self, rows: np.ndarray, codes: np.ndarray) -> None:
        want = np.float32 if self.dtype == "f32" else np.uint16
        rows = np.ascontiguousarray(rows)
        if rows.dtype != want or rows.ndim != 2 or rows.shape[1] != FRS_DIM:
            raise ValueError(f"raw rows must be [n, 384] {np.dtype(want).name} for a {self.dtype} index")
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        check(self._lib.frs_sharded_import_raw(self._h, _np_ptr(rows), _np_ptr(codes), rows.shape[0]))

    # -- search -----------------------------------------------------------------------------------
    @staticmethod
    def _check_batch(nq: int, k: int) -> None:
        if not 1 <= nq <= FRS_MAX_BATCH:
            raise ValueError(f"1..{FRS_MAX_BATCH} queries per call (got {nq})")
        if not 1 <= k <= FRS_MAX_K:
            raise ValueError(f"k must be in 1..{FRS_MAX_K} (got {k})")

    def search(self, queries, q_code, q_mask, k: int = 15):
        """numpy in -> (ids int64 [nq,k], scores float32 [nq,k]) numpy; thread-safe (each call takes a slot)."""
        return self.collect(self.submit(queries, q_code, q_mask, k))

    def submit(self, queries, q_code, q_mask, k: int = 15) -> int:
        q = np.ascontiguousarray(_to_numpy(queries), dtype=np.float32)
        nq = q.shape[0]
        self._check_batch(nq, k)
        qc, qm = _u32(_to_numpy(q_code)), _u32(_to_numpy(q_mask))
        t = C.c_int(-1)
        check(self._lib.frs_sharded_search_host_submit(self._h, _np_ptr(q), _np_ptr(qc), _np_ptr(qm), nq, k, C.byref(t)))
        self._shape[t.value] = (nq, k)
        return t.value

    def collect(self, ticket: int):
        nq, k = self._shape.pop(ticket)
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        check(self._lib.frs_sharded_search_host_collect(self._h, int(ticket), _np_ptr(scores), _np_ptr(ids)))
        return ids, scores

    # -- shard access (device-side fill, diagnostics) ------------------------------------------------
    def shard_handle(self, s: int) -> C.c_void_p:
        return C.c_void_p(self._lib.frs_sharded_shard(self._h, int(s)))

    def add_device(self, vecs, codes) -> None:
        """Append rows that already live on a GPU (torch CUDA tensors): each placement block is copied to its
        shard's GPU and stored there (frs_index_add on the shard), no host round trip."""
        import torch

        n = int(vecs.shape[0])
        size = len(self)
        if size + n > self.capacity:
            raise ValueError("index full")
        o = 0
        while o < n:
            g = size + o
            m = min(n - o, self.block_rows - g % self.block_rows)
            s, _ = self.placement(g)
            dev = torch.device("cuda", self.devices[s])
            v = vecs[o:o + m].to(device=dev, dtype=torch.float32).contiguous()
            c = codes[o:o + m].to(device=dev, dtype=torch.int32).contiguous()
            st = torch.cuda.current_stream(dev)
            check(self._lib.frs_index_add(self.shard_handle(s), C.c_void_p(v.data_ptr()), C.c_void_p(c.data_ptr()), m,
                                          C.c_void_p(st.cuda_stream)))
            st.synchronize()
            o += m
        check(self._lib.frs_sharded_set_size(self._h, size + n))

    def last_queries(self) -> np.ndarray:
        """The prepared queries of shard 0's last search, [32, 384] float32 (every shard prepares them identically)."""
        import torch

        dev = torch.device("cuda", self.devices[0])
        out = torch.empty((FRS_MAX_BATCH, FRS_DIM), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev)
        check(self._lib.frs_index_last_queries(self.shard_handle(0), C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream)))
        st.synchronize()
        return out.cpu().numpy()


def _to_numpy(x):
    if hasattr(x, "detach"):  # torch tensor
        return x.detach().cpu().numpy()
    return np.asarray(x)

"""ShardedIndex — corpus (row) sharding of the chunk store over the GPUs of one box.

One process per GPU (torch.distributed).  Rank r owns the contiguous global row range
[base_r, base_r + len_r): brute force reads every row whatever the filter, so contiguous placement
is balanced.  A search is

    local pass    every rank scans its shard for all queries -> exact local top-k as
                  (float64 score, int64 global id)                       [CUDA: prep/scan/merge]
    exchange      every rank WRITES its nq * k * 16 bytes into every peer's gather buffer over NVLink
                  peer memory and publishes a sequence flag (csrc/exchange.cu, `PeerExchange`); or, with
                  exchange="nccl", ONE all-gather of world * nq * k * 16 bytes           [CUDA | NCCL]
    final merge   [world, nq, k] -> [nq, k] by (score desc, id asc) on every rank      [CUDA]

The per-shard lists are exact, so the result is identical to a single-shard search of the whole
corpus whatever the number of shards.  `search_async` is the pipelined form: ONE call into the library
per batch (frs_index_search_async with the exchange) enqueues prep / scan / merge+push / wait+merge on the
index's internal streams, so the scans of consecutive batches run back to back and everything else —
including the exchange — hides behind them; no NCCL kernel is in the data path.

The reference has no counterpart (one Qdrant server, main.py:215-239); this is north-star item (3).
`local_search` / `merge` are injectable so that the host-side logic (partitioning, gather layout,
ordering) is exercised on CPU under the gloo backend in tests/test_sharded_cpu.py.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_range(total_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous partition: the first (total % world) shards get one extra row."""
    q, r = divmod(int(total_rows), int(world))
    start = rank * q + min(rank, r)
    return start, q + (1 if rank < r else 0)


class PendingSearch:
    """Result of search_async (all-gather form): tensors are valid once `ready` has been waited on."""

    def __init__(self, ids: torch.Tensor, scores: torch.Tensor, ready):
        self.ids, self.scores, self._ready = ids, scores, ready

    def wait(self):
        if self._ready is not None:
            self._ready.synchronize()
        return self.ids, self.scores

    def wait_stream(self):
        if self._ready is not None:
            torch.cuda.current_stream(self.ids.device).wait_event(self._ready)
        return self.ids, self.scores


class PeerExchange:
    """The exchange step over NVLink peer memory (include/frs_b200.h frs_exchange_*): push = this rank's
    block into every peer's gather buffer + a sequence flag; wait_merge = wait for all ranks' flags, then the
    cross-shard merge.  ONE object serves every batch size <= nq and limit <= k (default 32 / 32 = the ABI's
    maxima), so nothing is created in the request path.  Every rank pushes / merges once per batch, same order."""

    def __init__(self, device: torch.device, world: int, rank: int, nq: int = 32, k: int = 32, group=None,
                 connect: bool = True, timeout_ms: Optional[int] = None):
        import ctypes as C

        from . import _lib

        self._lib, self._C = _lib, C
        self.device, self.world, self.rank, self.nq, self.k = device, int(world), int(rank), int(nq), int(k)
        h = C.c_void_p()
        if not connect:
            _lib.check(_lib.lib().frs_exchange_create(device.index or 0, self.world, self.rank, self.nq, self.k, C.byref(h)))
            self._h = h
            if timeout_ms:
                self.set_timeout_ms(timeout_ms)
            return
        # Collective set-up: every rank takes part in the handle all-gather and in the final agreement even if one
        # of its own steps failed (CUDA IPC can be unavailable, e.g. in a restricted container), so that no rank is
        # left waiting in a collective and all ranks reach the same verdict.
        self._h, err = None, None
        buf = (C.c_uint8 * 128)()
        try:
            _lib.check(_lib.lib().frs_exchange_create(device.index or 0, self.world, self.rank, self.nq, self.k, C.byref(h)))
            self._h = h
            _lib.check(_lib.lib().frs_exchange_handle(self._h, buf))
        except Exception as e:  # noqa: BLE001
            err = e
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=device)
        every = torch.empty(self.world * 128, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(every, mine, group=group)
        if err is None:
            try:
                raw = bytes(every.cpu().numpy().tobytes())
                _lib.check(_lib.lib().frs_exchange_connect(self._h, C.cast(C.c_char_p(raw), C.c_void_p)))
            except Exception as e:  # noqa: BLE001
                err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"peer-memory exchange unavailable on at least one rank (this rank: {err})")
        if timeout_ms:
            self.set_timeout_ms(timeout_ms)

    @staticmethod
    def link(exchanges) -> None:
        """In-process form: several shards of one process (tests); exchanges[r] is rank r."""
        import ctypes as C

        from . import _lib

        arr = (C.c_void_p * len(exchanges))(*[e._h for e in exchanges])
        for e in exchanges:
            _lib.check(_lib.lib().frs_exchange_connect_local(e._h, arr))

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_timeout_ms(self, ms: int) -> None:
        self._lib.check(self._lib.lib().frs_exchange_set_timeout_ms(self._h, int(ms)))

    def status(self) -> None:
        """Raises FrsError (FRS_E_TIMEOUT) once a peer failed to publish a batch in time."""
        self._lib.check(self._lib.lib().frs_exchange_status(self._h))

    def push(self, local_packed: torch.Tensor) -> None:
        """Stand-alone push of a block in the exchange's layout ([2, nq, k] dense for the shapes it was created with)."""
        assert local_packed.dtype == torch.int64 and tuple(local_packed.shape) == (2, self.nq, self.k) and local_packed.is_contiguous()
        self._lib.check(self._lib.lib().frs_exchange_push(self._h, self._C.c_void_p(local_packed.data_ptr()), self._stream()))

    def wait_merge(self, nq: Optional[int] = None, k: Optional[int] = None):
        nq, k = int(nq or self.nq), int(k or self.k)
        out_s = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        self._lib.check(self._lib.lib().frs_exchange_wait_merge_n(self._h, nq, k, self._C.c_void_p(out_s.data_ptr()),
                                                                  self._C.c_void_p(out_i.data_ptr()), self._stream()))
        return out_i, out_s

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._lib.lib().frs_exchange_destroy(self._h)
            self._h = None


class ShardedIndex:
    def __init__(self, local_index, rank: int, world: int, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 device: Optional[torch.device] = None, reserve_sms: int = 8, exchange: Optional[str] = None,
                 timeout_ms: Optional[int] = None):
        """local_index: a VectorIndex whose base is this shard's first global row (or any object
        when local_search/merge are injected).

        exchange: "p2p" (default on CUDA) = stores into the peers' buffers over NVLink peer memory, fused into the
        local merge kernel, flags + a bounded wait (csrc/exchange.cu); ONE exchange object (32 queries x 32) is
        created here, collectively, and serves every batch.  "nccl" = one all-gather of the packed
        (score, id) lists per batch (the form north_star names; also what the injectable CPU path under gloo
        uses).  "auto" = p2p, falling back to nccl on every rank if CUDA IPC is unavailable on any.

        reserve_sms (nccl form only): the scan kernel is persistent with one CTA per SM and all of shared memory,
        so a collective kernel still resident when the next scan starts keeps scan CTAs waiting for an SM; the
        scan then leaves `reserve_sms` SMs free.  The p2p pipelined form reserves SMs inside the library
        (frs_index_set_pipeline_reserve)."""
        self.local = local_index
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = device if device is not None else getattr(local_index, "device", torch.device("cpu"))
        self._local_search = local_search or self._cuda_local_search
        self._merge = merge or self._cuda_merge
        self._side = None
        self._slot = 0
        self._slot_free = [None, None]  # event: the side stream is done with this slot's buffers
        self._bufs = {}
        import os

        cuda_path = self.world > 1 and merge is None and local_search is None and self.device.type == "cuda"
        self.exchange = (exchange or os.environ.get("FRS_EXCHANGE") or "auto").lower()
        if self.exchange not in ("auto", "p2p", "nccl"):
            raise ValueError(f"exchange must be 'auto', 'p2p' or 'nccl', got {self.exchange!r}")
        if not cuda_path:
            self.exchange = "nccl"
        self._peer: Optional[PeerExchange] = None
        if self.exchange in ("auto", "p2p"):
            try:
                self._peer = PeerExchange(self.device, self.world, self.rank, group=self.group, timeout_ms=timeout_ms)
                self.exchange = "p2p"
            except RuntimeError as e:
                if self.exchange != "auto":
                    raise
                import warnings

                warnings.warn(f"{e}; using the NCCL all-gather exchange")
                self.exchange = "nccl"
        if self.exchange == "nccl" and cuda_path and reserve_sms > 0:
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            if sms > 2 * reserve_sms and hasattr(local_index, "set_scan_grid"):
                local_index.set_scan_grid(sms - reserve_sms)

    # -- default (CUDA) implementations -----------------------------------------------------------
    def _cuda_local_search(self, q, qc, qm, k, out_scores64, out_ids):
        self.local.search_local(q, qc, qm, k, out_scores64, out_ids)

    @staticmethod
    def _cuda_merge(packed, k):
        from .index import merge_shards_packed

        return merge_shards_packed(packed, k)

    # -- buffers ------------------------------------------------------------------------------------
    def _buffers(self, nq: int, k: int, slot: int):
        key = (nq, k, slot)
        if key not in self._bufs:
            # one int64 tensor carries both halves of the candidates: plane 0 = fp64 score bits,
            # plane 1 = global ids; the local pass writes straight into the planes
            loc = torch.empty((2, nq, k), dtype=torch.int64, device=self.device)
            gat = torch.empty((self.world, 2, nq, k), dtype=torch.int64, device=self.device)
            self._bufs[key] = (loc, gat)
        return self._bufs[key]

    def _exchange_and_merge(self, loc: torch.Tensor, gat: torch.Tensor, k: int):
        if self.world > 1:
            # output viewed as the concatenation of the per-rank inputs along dim 0 (what gloo expects;
            # NCCL accepts both forms)
            dist.all_gather_into_tensor(gat.view(-1, gat.shape[2], gat.shape[3]), loc, group=self.group)
        else:
            gat.copy_(loc.unsqueeze(0))
        return self._merge(gat, k)

    def _local_pass(self, q, qc, qm, k, loc):
        self._local_search(q, qc, qm, k, loc[0].view(torch.float64), loc[1])

    # -- public -------------------------------------------------------------------------------------
    def search(self, queries, q_code, q_mask, k: int = 15):
        """Synchronous-in-stream sharded search; every rank must call it with the same queries.
        Returns (ids int64 [nq,k] global, scores float32 [nq,k]) on every rank."""
        q, qc, qm = self._prep(queries, q_code, q_mask)
        if self._peer is not None:
            # after pipelined calls: their exchanges come first (sequence numbers, gather slots)
            self.local.wait(-1)
            # the local merge kernel writes the shard's top-k into every peer's gather buffer itself
            self.local.search_push(q, qc, qm, k, self._peer)
            return self._peer.wait_merge(q.shape[0], k)
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        loc, gat = self._buffers(q.shape[0], k, 0)
        self._local_pass(q, qc, qm, k, loc)
        return self._exchange_and_merge(loc, gat, k)

    def search_async(self, queries, q_code, q_mask, k: int = 15):
        """Pipelined variant (CUDA only).  p2p: one library call; the prep of the next batch and the merge + push +
        wait + cross-shard merge of the previous one overlap this batch's scan on the index's internal streams.
        nccl: the all-gather and final merge run on a side stream.  Returns an object with .wait() -> (ids, scores)."""
        assert self.device.type == "cuda", "search_async needs CUDA streams"
        q, qc, qm = self._prep(queries, q_code, q_mask)
        if self._peer is not None:
            return self.local.search_async(q, qc, qm, k, exchange=self._peer)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        slot = self._slot
        self._slot ^= 1
        loc, gat = self._buffers(q.shape[0], k, slot)
        # the buffers of this slot were last used two calls ago on the side stream
        if self._slot_free[slot] is not None:
            main.wait_event(self._slot_free[slot])
        self._local_pass(q, qc, qm, k, loc)
        done_local = torch.cuda.Event()
        done_local.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(done_local)
            ids, scores = self._exchange_and_merge(loc, gat, k)
            ready = torch.cuda.Event()
            ready.record(self._side)
        self._slot_free[slot] = ready
        return PendingSearch(ids, scores, ready)

    def drain(self) -> None:
        """The current stream waits for every pipelined search issued so far."""
        if self._peer is not None:
            self.local.wait(-1)
        elif self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)

    # -- host buffers in / out (what a request handler calls) -------------------------------------------
    def submit_host(self, queries, q_code, q_mask, k: int = 15) -> int:
        """numpy in; returns a ticket for collect_host.  One pinned H2D + one D2H copy per batch inside the library,
        overlapping the neighbouring batches (up to 4 in flight).  Every rank submits the same batches in order."""
        if self._peer is None:
            raise RuntimeError("submit_host needs the peer-memory exchange (exchange='p2p')")
        return self.local.submit_host(queries, q_code, q_mask, k, exchange=self._peer)

    def collect_host(self, ticket: int):
        return self.local.collect_host(ticket, exchange=self._peer)

    def close(self) -> None:
        """Releases the peer-memory exchange (IPC mappings, gather buffers).  The local index stays open."""
        if self._side is not None:
            self._side.synchronize()
        if self._peer is not None:
            if getattr(self.local, "_h", None):   # (a local index that was closed first has nothing left to drain)
                self.local.sync(-1)
            self._peer.close()
            self._peer = None

    def _prep(self, queries, q_code, q_mask):
        q = torch.as_tensor(queries).to(device=self.device, dtype=torch.float32).contiguous()
        qc = torch.as_tensor(q_code).to(device=self.device)
        qm = torch.as_tensor(q_mask).to(device=self.device)
        if qc.dtype != torch.int32:
            qc = qc.to(torch.int64).to(torch.int32)
        if qm.dtype != torch.int32:
            qm = qm.to(torch.int64).to(torch.int32)
        return q, qc.contiguous(), qm.contiguous()

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
corpus whatever the number of shards.  The exchange + final merge of batch i run on a side stream
and overlap the local pass of batch i+1 (`search_async`).

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
    """Result of search_async: tensors are valid once `ready` has been waited on."""

    def __init__(self, ids: torch.Tensor, scores: torch.Tensor, ready):
        self.ids, self.scores, self._ready = ids, scores, ready

    def wait(self):
        if self._ready is not None:
            self._ready.synchronize()
        return self.ids, self.scores


class PeerExchange:
    """The exchange step over NVLink peer memory (include/frs_b200.h frs_exchange_*): push = this rank's
    [2, nq, k] block into every peer's gather buffer + a sequence flag; wait_merge = wait for all ranks' flags,
    then the cross-shard merge.  One object per (nq, k).  Every rank calls push / wait_merge once per batch."""

    def __init__(self, device: torch.device, world: int, rank: int, nq: int, k: int, group=None, connect: bool = True):
        import ctypes as C

        from . import _lib

        self._lib, self._C = _lib, C
        self.device, self.world, self.rank, self.nq, self.k = device, int(world), int(rank), int(nq), int(k)
        h = C.c_void_p()
        if not connect:
            _lib.check(_lib.lib().frs_exchange_create(device.index or 0, self.world, self.rank, self.nq, self.k, C.byref(h)))
            self._h = h
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

    def push(self, local_packed: torch.Tensor) -> None:
        assert local_packed.dtype == torch.int64 and tuple(local_packed.shape) == (2, self.nq, self.k) and local_packed.is_contiguous()
        self._lib.check(self._lib.lib().frs_exchange_push(self._h, self._C.c_void_p(local_packed.data_ptr()), self._stream()))

    def wait_merge(self):
        out_s = torch.empty((self.nq, self.k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((self.nq, self.k), dtype=torch.int64, device=self.device)
        self._lib.check(self._lib.lib().frs_exchange_wait_merge(self._h, self._C.c_void_p(out_s.data_ptr()),
                                                                self._C.c_void_p(out_i.data_ptr()), self._stream()))
        return out_i, out_s

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._lib.lib().frs_exchange_destroy(self._h)
            self._h = None


class ShardedIndex:
    def __init__(self, local_index, rank: int, world: int, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 device: Optional[torch.device] = None, reserve_sms: int = 8, exchange: Optional[str] = None):
        """local_index: a VectorIndex whose base is this shard's first global row (or any object
        when local_search/merge are injected).

        reserve_sms: the scan kernel is persistent with one CTA per SM and all of shared memory, so a
        collective kernel that is still resident when the next scan starts keeps scan CTAs waiting for an
        SM — a static tile split then ends a whole "wave" late (measured on 8 B200, 1.25M rows per GPU:
        scan 255 us instead of 151 us, 110k instead of 181k QPS).  With world > 1 the scan therefore
        leaves `reserve_sms` SMs to the exchange (NCCL_MAX_NCHANNELS=1..2 keeps NCCL inside them)."""
        self.local = local_index
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = device if device is not None else getattr(local_index, "device", torch.device("cpu"))
        self._local_search = local_search or self._cuda_local_search
        self._merge = merge or self._cuda_merge
        self._side = None
        self._slot = 0
        self._slot_free = [None, None]  # event: the side stream is done with this slot's buffers
        self._bufs = {}
        # exchange step: "p2p" = writes into the peers' buffers over NVLink (PeerExchange), "nccl" = one all-gather.
        # The injectable CPU path (gloo tests) and world 1 use the collective form.
        import os

        # "auto" (default): the synchronous search() — the latency path a request sees — uses the push fused into
        # the local merge kernel (8 GPUs, 10M rows: 109 k vs 86 k QPS host to host), the pipelined search_async()
        # keeps the all-gather on its side stream, where it is hidden completely behind the next local pass
        # (181 k vs 137 k QPS device-timed: the fused push sits on the main stream and lengthens the merge kernel).
        cuda_path = self.world > 1 and merge is None and self.device.type == "cuda"
        self.exchange = (exchange or os.environ.get("FRS_EXCHANGE") or "auto").lower()
        if self.exchange not in ("auto", "p2p", "nccl"):
            raise ValueError(f"exchange must be 'auto', 'p2p' or 'nccl', got {self.exchange!r}")
        if not cuda_path:
            self.exchange = "nccl"
        self._peer = {}
        if self.world > 1 and local_search is None and self.device.type == "cuda" and reserve_sms > 0:
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

    def _peer_exchange(self, nq: int, k: int) -> Optional["PeerExchange"]:
        """Collective on first use: every rank creates and connects it.  In "auto" mode a set-up that fails on any
        rank (no CUDA IPC) switches every rank to the all-gather form — both are GPU paths; "p2p" raises."""
        key = (nq, k)
        if key not in self._peer:
            try:
                self._peer[key] = PeerExchange(self.device, self.world, self.rank, nq, k, group=self.group)
            except RuntimeError as e:
                if self.exchange != "auto":
                    raise
                import warnings

                warnings.warn(f"{e}; using the NCCL all-gather exchange")
                self.exchange = "nccl"
                return None
        return self._peer[key]

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
        if self._side is not None:  # after pipelined calls: their exchanges come first (sequence numbers, slots)
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        ex = self._peer_exchange(q.shape[0], k) if self.exchange in ("p2p", "auto") else None
        if ex is not None:
            # the local merge kernel writes the shard's top-k into every peer's gather buffer itself
            self.local.search_push(q, qc, qm, k, ex)
            return ex.wait_merge()
        loc, gat = self._buffers(q.shape[0], k, 0)
        self._local_pass(q, qc, qm, k, loc)
        return self._exchange_and_merge(loc, gat, k)

    def search_async(self, queries, q_code, q_mask, k: int = 15) -> PendingSearch:
        """Pipelined variant (CUDA only): the exchange and final merge run on a side stream so the
        next call's local pass overlaps them.  Alternates between two buffer sets."""
        assert self.device.type == "cuda", "search_async needs CUDA streams"
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        q, qc, qm = self._prep(queries, q_code, q_mask)
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
        # "p2p": the stand-alone push of the finished block + wait + merge, all on the side stream (the push of
        # batch t + 1 is stream-ordered behind this rank's merge of batch t, the d = 1 case of the slot rule in
        # csrc/scan.cuh).  The push FUSED into the merge kernel would sit on the caller's stream in front of the
        # next scan (8 GPUs: 137 k against 181 k QPS), so the pipelined form does not use it.
        ex = self._peer_exchange(q.shape[0], k) if self.exchange == "p2p" else None
        with torch.cuda.stream(self._side):
            self._side.wait_event(done_local)
            if ex is not None:
                ex.push(loc)
                ids, scores = ex.wait_merge()
            else:
                ids, scores = self._exchange_and_merge(loc, gat, k)
            ready = torch.cuda.Event()
            ready.record(self._side)
        self._slot_free[slot] = ready
        return PendingSearch(ids, scores, ready)

    def close(self) -> None:
        """Releases the peer-memory exchanges (IPC mappings, gather buffers).  The local index stays open."""
        if self._side is not None:
            self._side.synchronize()
        for ex in self._peer.values():
            ex.close()
        self._peer = {}

    def _prep(self, queries, q_code, q_mask):
        q = torch.as_tensor(queries).to(device=self.device, dtype=torch.float32).contiguous()
        qc = torch.as_tensor(q_code).to(device=self.device)
        qm = torch.as_tensor(q_mask).to(device=self.device)
        if qc.dtype != torch.int32:
            qc = qc.to(torch.int64).to(torch.int32)
        if qm.dtype != torch.int32:
            qm = qm.to(torch.int64).to(torch.int32)
        return q, qc.contiguous(), qm.contiguous()

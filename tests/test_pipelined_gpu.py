"""Pipelined search, host submit/collect, the robust exchange and the one-process multi-GPU store
(frs_index_search_async / _host_submit / _host_collect, frs_exchange_*, frs_sharded_*), against the
synchronous search — itself checked against the oracle in test_search_gpu.py — and against the oracle."""
import threading

import numpy as np
import pytest
import torch

from oracle import search_oracle as so

pytestmark = pytest.mark.gpu

ANY = 0x80000000
TICKER = 0x80FFFFFF


def _i32(x):
    return torch.as_tensor(np.asarray(x, dtype=np.uint32).astype(np.int64)).to(torch.int32).cuda()


def _data(n, seed, tickers=5):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((n, 384), generator=g, device="cuda")
    t = torch.randint(0, tickers, (n,), generator=g, device="cuda", dtype=torch.int32)
    return x, t, g


def _index(n, dtype="bf16", **kw):
    from financial_rag_system_b200.index import VectorIndex

    return VectorIndex(max(n, 1), dtype=dtype, device=0, **kw)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("reserve", [0, 4, 16])
def test_pipelined_search_equals_synchronous_search(dtype, reserve):
    """12 batches of varying size / limit in flight back to back == the same batches searched one by one."""
    n = 60_000
    x, codes, g = _data(n, 3)
    ix = _index(n, dtype)
    ix.add(x, codes)
    ix.set_pipeline_reserve(reserve)
    batches = []
    for b in range(12):
        nq, k = [32, 7, 1, 19][b % 4], [15, 16, 3, 15][b % 4]
        q = x[b * 40:b * 40 + nq] + 0.1 * torch.randn((nq, 384), generator=g, device="cuda")
        qc = _i32(codes[b * 40:b * 40 + nq].cpu().numpy())
        qm = _i32(np.full(nq, TICKER if b % 3 else ANY, np.uint32))
        batches.append((q, qc, qm, k))
    want = [ix.search(*b) for b in batches]
    torch.cuda.synchronize()
    pend = [ix.search_async(*b) for b in batches]
    for (wi, ws), p in zip(want, pend):
        gi, gs = p.wait()
        assert torch.equal(gi, wi) and torch.equal(gs, ws)
    # stream-ordered wait instead of a host wait
    pend = [ix.search_async(*b) for b in batches[:5]]
    ix.wait(-1)
    torch.cuda.current_stream().synchronize()
    for (wi, ws), p in zip(want, pend):
        assert torch.equal(p.ids, wi) and torch.equal(p.scores, ws)
    ix.close()


def test_profiling_modes_leave_results_alone_and_account_for_the_run():
    """The three event-profiling modes (per-kernel events, bracket around the scan launches, bracket relative to the
    caller's own events = fill and drain of a pipelined run) return the same results as an unprofiled run and numbers
    that add up: fill + n x scan + drain <= the caller's region, every part positive."""
    n = 200_000
    x, codes, g = _data(n, 17)
    ix = _index(n)
    ix.add(x, codes)
    q, qc, qm = x[:32] + 0.05 * torch.randn((32, 384), generator=g, device="cuda"), _i32(codes[:32].cpu().numpy()), _i32(np.full(32, TICKER, np.uint32))
    wi, ws = ix.search(q, qc, qm, 15)
    torch.cuda.synchronize()
    # per-kernel events + the raw time line
    ix.set_profiling(1)
    pend = [ix.search_async(q, qc, qm, 15) for _ in range(6)]
    ix.wait(-1)
    torch.cuda.synchronize()
    tl = ix.read_profile_raw(16)
    assert tl.shape == (6, 7) and np.all(np.diff(tl, axis=1)[:, [0, 2, 4]] > 0)       # every kernel has a duration
    assert np.all(tl[:, 2] >= tl[:, 1]) and np.all(tl[:, 4] >= tl[:, 3])               # prep -> scan -> merge
    prof = ix.read_profile_ex()
    assert prof["n"] == 6 and prof["scan_ms"] > 0 and prof["span_ms"] >= tl[-1, 5] - 1e-3
    # bracket, relative to the caller's events
    ix.set_profiling(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pend += [ix.search_async(q, qc, qm, 15) for _ in range(6)]
    ix.wait(-1)
    e1.record()
    torch.cuda.synchronize()
    fill, drain = ix.read_profile_bracket_rel(e0, e1)
    br = ix.read_profile_ex()
    assert br["n"] == 6 and fill > 0 and drain > 0 and br["scan_ms"] > 0
    assert fill + br["scan_ms"] + drain <= e0.elapsed_time(e1) * 1.001 + 1e-3
    ix.set_profiling(0)
    for p in pend:
        assert torch.equal(p.ids, wi) and torch.equal(p.scores, ws)
    ix.close()


def test_host_submit_collect_keeps_batches_in_flight():
    n = 40_000
    x, codes, g = _data(n, 5)
    ix = _index(n)
    ix.add(x, codes)
    qs = [(x[i * 32:(i + 1) * 32] + 0.05 * torch.randn((32, 384), generator=g, device="cuda")).cpu().numpy() for i in range(9)]
    qc = [codes[i * 32:(i + 1) * 32].cpu().numpy().astype(np.uint32) for i in range(9)]
    qm = np.full(32, TICKER, np.uint32)
    want = [ix.search(q, c, qm, 15) for q, c in zip(qs, qc)]
    got, inflight = [], []
    for q, c in zip(qs, qc):
        inflight.append(ix.submit_host(q, c, qm, 15))
        if len(inflight) == 4:  # the library holds 4 slots
            got.append(ix.collect_host(inflight.pop(0)))
    while inflight:
        got.append(ix.collect_host(inflight.pop(0)))
    for (wi, ws), (gi, gs) in zip(want, got):
        assert np.array_equal(gi, wi) and np.array_equal(gs, ws)
    ix.close()


def test_many_threads_call_the_host_entry_point_at_once():
    """25 request threads (main2.py:52-53) on one index: every call gets its own staging slot, results are the
    single-threaded ones."""
    n = 30_000
    x, codes, g = _data(n, 7)
    ix = _index(n)
    ix.add(x, codes)
    qs = [torch.randn((1 + i % 32, 384), generator=g, device="cuda").cpu().numpy() for i in range(25)]
    want = [ix.search(q, np.zeros(len(q)), np.full(len(q), ANY, np.uint32), 15) for q in qs]
    got, errs = [None] * len(qs), []

    def work(i):
        try:
            for _ in range(8):
                got[i] = ix.search(qs[i], np.zeros(len(qs[i])), np.full(len(qs[i]), ANY, np.uint32), 15)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(qs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for w, g_ in zip(want, got):
        assert np.array_equal(w[0], g_[0]) and np.array_equal(w[1], g_[1])
    ix.close()


def test_searches_during_ingest_never_see_half_written_rows():
    """ingest.py upserts while main2.py serves queries: a search that sees the new size also sees the new rows."""
    n, step = 64_000, 4_000
    x, codes, g = _data(n, 9)
    ix = _index(n)
    stop, errs = threading.Event(), []
    q = x[:16].cpu().numpy()  # each query is a stored row: once row i is visible it must be its own best hit, score 1

    def reader():
        try:
            while not stop.is_set():
                size = len(ix)
                if size == 0:
                    continue
                ids, sc = ix.search(q, np.zeros(16), np.full(16, ANY, np.uint32), 1)
                ok = ids[:, 0] >= 0
                assert np.all(np.isfinite(sc[ok, 0])) and np.all(sc[ok, 0] <= 1.0 + 1e-3)
                if size >= 16:
                    assert np.array_equal(ids[:, 0], np.arange(16)), ids[:, 0]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=reader) for _ in range(3)]
    [t.start() for t in th]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for s in range(0, n, step):
            ix.add(x[s:s + step], codes[s:s + step])
    stop.set()
    [t.join() for t in th]
    assert not errs, errs
    assert len(ix) == n
    ix.close()


def _three_shards(n, cuts, x, codes, dtype="bf16"):
    from financial_rag_system_b200.sharded import PeerExchange

    shards = []
    for r in range(len(cuts) - 1):
        ix = _index(cuts[r + 1] - cuts[r], dtype, base=cuts[r])
        ix.add(x[cuts[r]:cuts[r + 1]], codes[cuts[r]:cuts[r + 1]])
        shards.append(ix)
    dev = torch.device("cuda", 0)
    exs = [PeerExchange(dev, len(shards), r, connect=False) for r in range(len(shards))]  # 32 x 16: serves every batch
    PeerExchange.link(exs)
    return shards, exs


def test_pipelined_sharded_search_with_one_exchange_for_every_batch_shape():
    """Three shards (one process, linked by pointer), ONE exchange object of 32 x 16 per shard: batches of
    different nq and k go through frs_index_search_async with the exchange and equal the one-index search."""
    n, cuts = 45_001, [0, 9_000, 21_345, 45_001]
    x, codes, g = _data(n, 11)
    x[30_000] = x[3].clone()   # an exact duplicate across shards: ties break by global id
    codes[30_000] = codes[3].clone()
    whole = _index(n)
    whole.add(x, codes)
    shards, exs = _three_shards(n, cuts, x, codes)
    batches = []
    for b in range(10):
        nq, k = [32, 5, 17, 1][b % 4], [15, 16, 4, 15][b % 4]
        q = x[b * 32:b * 32 + nq] + 0.1 * torch.randn((nq, 384), generator=g, device="cuda")
        batches.append((q, _i32(codes[b * 32:b * 32 + nq].cpu().numpy()), _i32(np.full(nq, TICKER if b % 2 else ANY, np.uint32)), k))
    want = [whole.search(*b) for b in batches]
    torch.cuda.synchronize()
    pend = [[shards[r].search_async(*b, exchange=exs[r]) for r in range(3)] for b in batches]
    for (wi, ws), ps in zip(want, pend):
        for r, p in enumerate(ps):
            gi, gs = p.wait()
            assert torch.equal(gi, wi), r
            assert torch.equal(gs, ws), r
    # host buffers through the same exchange
    tick = [[shards[r].submit_host(b[0].cpu().numpy(), b[1].cpu().numpy(), b[2].cpu().numpy(), b[3], exchange=exs[r]) for r in range(3)]
            for b in batches[:3]]
    for (wi, ws), ts in zip(want, tick):
        for r, t in enumerate(ts):
            gi, gs = shards[r].collect_host(t, exchange=exs[r])
            assert np.array_equal(gi, wi.cpu().numpy()) and np.array_equal(gs, ws.cpu().numpy())
    for e in exs:
        e.status()
        e.close()
    for ix in [whole] + shards:
        ix.close()


def test_a_lost_peer_times_out_without_destroying_the_context():
    """Only rank 0 of a two-rank exchange ever pushes.  The wait gives up after the (shortened) time-out, the
    batch comes back empty, the status is FRS_E_TIMEOUT — and the CUDA context, with the resident shard, lives."""
    from financial_rag_system_b200._lib import FRS_E_TIMEOUT, FrsError
    from financial_rag_system_b200.sharded import PeerExchange

    n = 5_000
    x, codes, g = _data(n, 13)
    ix = _index(n)
    ix.add(x, codes)
    dev = torch.device("cuda", 0)
    exs = [PeerExchange(dev, 2, r, connect=False, timeout_ms=100) for r in range(2)]
    PeerExchange.link(exs)
    q, qc, qm = x[:8], _i32(np.zeros(8)), _i32(np.full(8, ANY, np.uint32))
    p = ix.search_async(q, qc, qm, 15, exchange=exs[0])
    ids, scores = p.wait()
    assert torch.all(ids == -1) and torch.all(torch.isinf(scores))
    with pytest.raises(FrsError) as ei:
        exs[0].status()
    assert ei.value.code == FRS_E_TIMEOUT
    with pytest.raises(FrsError):  # a poisoned exchange refuses further batches
        ix.search_async(q, qc, qm, 15, exchange=exs[0])
    wi, ws = ix.search(q, qc, qm, 15)   # the shard itself is untouched
    torch.cuda.synchronize()
    assert int(wi[0, 0]) == 0 and float(ws[0, 0]) > 0.99
    for e in exs:
        e.close()
    ix.close()


def test_a_late_peer_is_waited_for():
    """Rank 1 pushes 300 ms after rank 0 started waiting (a host stall): the result is complete and correct."""
    import time

    n, cut = 20_000, 8_000
    x, codes, g = _data(n, 15)
    whole = _index(n)
    whole.add(x, codes)
    shards, exs = _three_shards(n, [0, cut, n], x, codes)
    q, qc, qm = x[:32] + 0.1 * torch.randn((32, 384), generator=g, device="cuda"), _i32(np.zeros(32)), _i32(np.full(32, ANY, np.uint32))
    wi, ws = whole.search(q, qc, qm, 15)
    p0 = shards[0].search_async(q, qc, qm, 15, exchange=exs[0])
    time.sleep(0.3)
    p1 = shards[1].search_async(q, qc, qm, 15, exchange=exs[1])
    for p in (p0, p1):
        gi, gs = p.wait()
        assert torch.equal(gi, wi) and torch.equal(gs, ws)
    for e in exs:
        e.close()
    for ix in [whole] + shards:
        ix.close()


# ---- one process, several GPUs (frs_sharded) ----------------------------------------------------
def _devices(n):
    have = torch.cuda.device_count()
    return list(range(n)) if have >= n else [i % have for i in range(n)]  # several shards per GPU on a small box


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("n_shards", [1, 2, 3])
def test_one_process_multi_gpu_store_matches_the_oracle(dtype, n_shards):
    from financial_rag_system_b200.multigpu import MultiGpuIndex

    n = 3 * 4096 * n_shards + 1234   # several placement blocks per shard + a ragged tail
    rng = np.random.default_rng(17)
    x = rng.standard_normal((n, 384)).astype(np.float32)
    codes = rng.integers(0, 5, n).astype(np.uint32)
    x[9000] = x[3]   # duplicate in another block / shard: tie broken by global id
    codes[9000] = codes[3]
    mg = MultiGpuIndex(n + 100, dtype=dtype, devices=_devices(n_shards))
    mg.add(x[:5000], codes[:5000])   # ragged appends across block boundaries
    mg.add(x[5000:5001], codes[5000:5001])
    mg.add(x[5001:], codes[5001:])
    assert len(mg) == n
    rows = mg.read_rows()
    assert np.allclose(rows, so.store_rows(x, dtype), rtol=1e-2, atol=1e-4)   # (the norm's summation order differs)
    for nq, k, mask in ((32, 15, TICKER), (5, 16, ANY), (1, 1, TICKER)):
        q = x[:nq] + 0.1 * rng.standard_normal((nq, 384)).astype(np.float32)
        qc, qm = codes[:nq], np.full(nq, mask, np.uint32)
        ids, sc = mg.search(q, qc, qm, k)
        qp = mg.last_queries()[:nq]
        oi, os_ = so.exact_topk(rows, qp, codes, qc, qm, k)
        assert np.array_equal(ids, oi), (nq, k)
        assert np.allclose(sc, os_, atol=1e-6)
    # in-place overwrite + tombstone through global row numbers
    mg.set_rows(7000, x[11:12], codes[11:12])
    mg.set_codes(11, np.array([codes[11] | 0x80000000], dtype=np.uint32))
    ids, _ = mg.search(x[11:12], codes[11:12], np.array([TICKER], np.uint32), 3)
    assert ids[0, 0] == 7000 and 11 not in ids[0]
    # raw export / import round trip into a store with a different shard count
    raw_rows, raw_codes = mg.export_raw()
    other = MultiGpuIndex(n + 100, dtype=dtype, devices=_devices(2 if n_shards != 2 else 3))
    other.import_raw(raw_rows, raw_codes)
    q = x[100:132]
    a = mg.search(q, codes[100:132], np.full(32, TICKER, np.uint32), 15)
    b = other.search(q, codes[100:132], np.full(32, TICKER, np.uint32), 15)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    other.close()
    mg.close()


def test_one_process_multi_gpu_store_pipelines_and_serves_threads():
    from financial_rag_system_b200.multigpu import MultiGpuIndex

    n = 50_000
    x, codes, g = _data(n, 19)
    mg = MultiGpuIndex(n, devices=_devices(2))
    mg.add_device(x, codes)   # rows already on a GPU: block by block to their shards, no host round trip
    one = _index(n)
    one.add(x, codes)
    qs = [torch.randn((32, 384), generator=g, device="cuda").cpu().numpy() for _ in range(8)]
    z, m = np.zeros(32, np.uint32), np.full(32, ANY, np.uint32)
    want = [one.search(q, z, m, 15) for q in qs]
    tickets = [mg.submit(q, z, m, 15) for q in qs[:4]]
    for t, w in zip(tickets, want):
        gi, gs = mg.collect(t)
        assert np.array_equal(gi, w[0]) and np.array_equal(gs, w[1])
    got, errs = [None] * len(qs), []

    def work(i):
        try:
            for _ in range(4):
                got[i] = mg.search(qs[i], z, m, 15)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(qs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for w, g_ in zip(want, got):
        assert np.array_equal(w[0], g_[0]) and np.array_equal(w[1], g_[1])
    one.close()
    mg.close()


def test_collection_and_qdrant_client_over_several_gpus():
    """QdrantCompat(devices=[...]): the reference's call sites, rows sharded behind the one client object."""
    from financial_rag_system_b200 import synth
    from financial_rag_system_b200.collection import QdrantCompat, models

    n = 9000
    ids, texts, payloads = synth.make_chunks(n, n_tickers=6, seed=5)
    rng = np.random.default_rng(1)
    vecs = rng.standard_normal((n, 384)).astype(np.float32)
    multi, single = QdrantCompat(capacity=n + 10, devices=_devices(2)), QdrantCompat(capacity=n + 10)
    for c in (multi, single):
        c.create_collection("sec_filings", models.VectorParams(size=384, distance=models.Distance.COSINE))
        for s in range(0, n, 256):   # ingest.py:148-175 batches of 256
            c.upsert("sec_filings", [models.PointStruct(id=ids[i], vector=vecs[i].tolist(), payload=payloads[i])
                                     for i in range(s, min(n, s + 256))])
    flt = models.Filter(must=[models.FieldCondition(key="ticker", match=models.MatchValue(value=payloads[77]["ticker"]))])
    a = multi.query_points("sec_filings", query=vecs[77].tolist(), limit=15, query_filter=flt).points
    b = single.query_points("sec_filings", query=vecs[77].tolist(), limit=15, query_filter=flt).points
    assert [p.id for p in a] == [p.id for p in b] and a[0].id == ids[77]
    assert np.allclose([p.score for p in a], [p.score for p in b], atol=1e-6)
    assert all(p.payload["ticker"] == payloads[77]["ticker"] for p in a)

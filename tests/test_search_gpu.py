"""Parity tests of the CUDA retrieval path against the oracle (oracle/search_oracle.py).

Everything here goes through the C ABI (financial_rag_system_b200.index.VectorIndex is a ctypes
shim).  The bar: ids identical to the oracle's exact top-k — (fp64 score desc, id asc) over the
rows the index actually stores and the query it actually prepared — and scores within 1e-6 (the
GPU reports float32(fp64 dot); north_star tolerance is 1e-5 in fp32 mode / 1e-2 in bf16 mode).
"""
import os
import threading

import numpy as np
import pytest
import torch

from oracle import search_oracle as so

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "search_golden.npz")
ANY = 0x80000000          # mask: only "not a tombstone"
TICKER = 0x80FFFFFF       # mask: ticker must match
TICKER_DOC = 0xFFFFFFFF   # mask: ticker and document_type must match
SCORE_ATOL = 1e-6


def _i32(x):
    return torch.as_tensor(np.asarray(x, dtype=np.uint32).astype(np.int64)).to(torch.int32).cuda()


def _index(n, dtype, **kw):
    from financial_rag_system_b200.index import VectorIndex

    return VectorIndex(max(n, 1), dtype=dtype, device=0, **kw)


def _check(ix, q, qc, qm, k, codes, base=0):
    """Search on the GPU, recompute with the oracle on the stored rows / prepared queries."""
    ids, sc = ix.search(q, _i32(qc), _i32(qm), k)
    torch.cuda.synchronize()
    n = len(ix)
    rows = ix.read_rows().cpu().numpy() if n else np.zeros((0, 384), np.float32)
    qp = ix.last_queries().cpu().numpy()[: q.shape[0]]
    oi, os_ = so.exact_topk(rows, qp, np.asarray(codes, np.uint32)[:n], np.asarray(qc, np.uint32), np.asarray(qm, np.uint32), k, base=base)
    gi, gs = ids.cpu().numpy(), sc.cpu().numpy().astype(np.float64)
    assert np.array_equal(gi, oi), f"ids differ\n got {gi}\nwant {oi}"
    fin = np.isfinite(os_)
    assert np.array_equal(np.isfinite(gs), fin)
    assert np.allclose(gs[fin], os_[fin], atol=SCORE_ATOL)
    return gi, gs


def _data(n, seed, tickers=5, clustered=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if clustered:
        cent = torch.randn((32, 384), generator=g, device="cuda")
        x = cent[torch.randint(0, 32, (n,), generator=g, device="cuda")] + 0.3 * torch.randn((n, 384), generator=g, device="cuda")
    else:
        x = torch.randn((n, 384), generator=g, device="cuda")
    t = torch.randint(0, tickers, (n,), generator=g, device="cuda", dtype=torch.int32)
    d = torch.randint(0, 2, (n,), generator=g, device="cuda", dtype=torch.int32)
    return x, (t | (d << 24)), g


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("n,nq,k", [(1, 1, 1), (100, 7, 15), (1000, 32, 15), (4097, 32, 16), (50_000, 32, 15), (200_000, 13, 5),
                                    (4097, 32, 32), (50_000, 9, 20), (300_000, 32, 32)])
def test_parity_sizes(dtype, n, nq, k):
    x, codes, g = _data(n, 100 + n)
    ix = _index(n, dtype)
    ix.add(x, codes)
    src = torch.randint(0, n, (nq,), generator=g, device="cuda")
    q = x[src] + 0.2 * torch.randn((nq, 384), generator=g, device="cuda")
    qc = (codes[src] & 0x7FFFFFFF).cpu().numpy()
    c = codes.cpu().numpy().astype(np.uint32)
    _check(ix, q, qc, np.full(nq, TICKER, np.uint32), k, c)        # ticker filter (main.py:218-223)
    _check(ix, q, qc, np.full(nq, TICKER_DOC, np.uint32), k, c)    # + document_type (main.py:224-230)
    _check(ix, q, qc, np.full(nq, ANY, np.uint32), k, c)           # no filter
    ix.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_clustered_corpus_dense_scores(dtype):
    """Scores packed inside the pre-filter error band: exercises the exact (fp64) band resolution."""
    n = 60_000
    x, codes, g = _data(n, 7, clustered=True)
    ix = _index(n, dtype)
    ix.add(x, codes)
    q = x[:32] + 0.02 * torch.randn((32, 384), generator=g, device="cuda")
    c = codes.cpu().numpy().astype(np.uint32)
    _check(ix, q, (codes[:32] & 0x7FFFFFFF).cpu().numpy(), np.full(32, ANY, np.uint32), 15, c)
    _check(ix, q, (codes[:32] & 0x7FFFFFFF).cpu().numpy(), np.full(32, TICKER, np.uint32), 15, c)
    ix.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_exact_duplicates_tie_break_by_id(dtype):
    n = 30_000
    x, codes, g = _data(n, 9)
    dup = torch.randint(0, n, (60,), generator=g, device="cuda")
    x[dup] = x[123].clone()              # 60 scattered copies of row 123
    x[5000:5040] = x[4999].clone()       # 40 contiguous copies (one tile / one CTA)
    codes[dup] = codes[123].clone()
    codes[5000:5040] = codes[4999].clone()
    ix = _index(n, dtype)
    ix.add(x, codes)
    q = torch.stack([x[123], x[4999], x[123] + 0.01 * torch.randn(384, generator=g, device="cuda")])
    qc = torch.stack([codes[123], codes[4999], codes[123]]).cpu().numpy() & 0x7FFFFFFF
    gi, gs = _check(ix, q, qc, np.full(3, TICKER, np.uint32), 15, codes.cpu().numpy().astype(np.uint32))
    assert np.all(np.diff(gi[0]) > 0) and np.allclose(gs[0], gs[0][0], atol=1e-7)   # 15 exact ties in id order
    assert gi[1].tolist() == list(range(4999, 5014))
    assert ix.last_stats()["launches"] == 3
    ix.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_rare_ticker_padding_and_no_match(dtype):
    n = 20_000
    x, codes, g = _data(n, 11)
    codes[:] = codes & 0x00000003
    codes[[5, 777, 19_999]] = 42          # a ticker with 3 rows
    ix = _index(n, dtype)
    ix.add(x, codes)
    q = torch.randn((4, 384), generator=g, device="cuda")
    gi, gs = _check(ix, q, [42, 42, 99, 1], np.full(4, TICKER, np.uint32), 15, codes.cpu().numpy().astype(np.uint32))
    assert sorted(gi[0][:3].tolist()) == [5, 777, 19_999] and (gi[0][3:] == -1).all() and np.isneginf(gs[0][3:]).all()
    assert (gi[2] == -1).all()            # ticker 99 does not exist: empty result like main.py:238-239
    ix.close()


def test_tombstones_and_upsert_in_place():
    n = 5000
    x, codes, g = _data(n, 13)
    ix = _index(n, "bf16")
    ix.add(x, codes)
    q = x[:4].clone()
    c = codes.cpu().numpy().astype(np.uint32)
    gi, _ = _check(ix, q, np.zeros(4), np.full(4, ANY, np.uint32), 15, c)
    assert gi[:, 0].tolist() == [0, 1, 2, 3]
    # delete rows 0..3: tombstone bit
    dead = (codes[:4].to(torch.int64) | 0x80000000).to(torch.int32)
    ix.set_codes(0, dead)
    c[:4] |= np.uint32(0x80000000)
    gi, _ = _check(ix, q, np.zeros(4), np.full(4, ANY, np.uint32), 15, c)
    assert not set(gi.flatten().tolist()) & {0, 1, 2, 3}
    # idempotent upsert: overwrite row 10 with row 0's vector
    ix.set_rows(10, x[:1], codes[10:11])
    gi, _ = _check(ix, q[:1], [0], [ANY], 15, c)
    assert gi[0][0] == 10
    ix.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_result_is_independent_of_the_scan_grid(dtype):
    n = 40_000
    x, codes, g = _data(n, 17)
    ix = _index(n, dtype)
    ix.add(x, codes)
    q = x[100:132] + 0.3 * torch.randn((32, 384), generator=g, device="cuda")
    qc, qm = _i32((codes[100:132] & 0x7FFFFFFF).cpu().numpy()), _i32(np.full(32, TICKER, np.uint32))
    ref = None
    for grid in (0, 1, 2, 13, 64, 147):
        ix.set_scan_grid(grid)
        ids, sc = ix.search(q, qc, qm, 15)
        torch.cuda.synchronize()
        cur = (ids.cpu().numpy().copy(), sc.cpu().numpy().copy())
        if ref is None:
            ref = cur
        assert np.array_equal(cur[0], ref[0])
        assert np.array_equal(cur[1].view(np.uint32), ref[1].view(np.uint32))   # bit-identical scores
    ix.close()


def test_golden_fixture():
    g = np.load(GOLD)
    k = int(g["k"])
    for dtype in ("bf16", "f32"):
        ix = _index(g["rows"].shape[0], dtype)
        ix.add(g["rows"], g["codes"])                       # host entry point (frs_index_add_host)
        ids, sc = ix.search(g["queries"], g["q_code"], g["q_mask"], k)   # host entry point
        stored = ix.read_rows().cpu().numpy()
        want_stored = so.store_rows(g["rows"], dtype)
        # the fp32 norm is summed in a different order on the GPU: last-ulp differences in f32 mode,
        # a rare flipped bf16 rounding in bf16 mode
        if dtype == "bf16":
            assert np.mean(stored != want_stored) < 1e-3
        assert np.abs(stored - want_stored).max() < (1e-6 if dtype == "f32" else 4e-3)
        qp = ix.last_queries().cpu().numpy()[: g["queries"].shape[0]]
        oi, os_ = so.exact_topk(stored, qp, g["codes"], g["q_code"], g["q_mask"], k)
        assert np.array_equal(ids, oi)
        # against the committed vectors: identical wherever the golden ranking is not a near-tie
        gold_i, gold_s = g[f"ids_{dtype}"], g[f"scores_{dtype}"]
        for qi in range(ids.shape[0]):
            for r in range(k):
                if ids[qi, r] != gold_i[qi, r]:
                    gap = np.abs(np.diff(gold_s[qi][max(r - 1, 0): r + 2])).min()
                    assert gap < 1e-3, (dtype, qi, r, ids[qi], gold_i[qi])
        fin = np.isfinite(gold_s)
        assert np.allclose(sc[fin], gold_s[fin], atol=1e-2 if dtype == "bf16" else 1e-5)
        ix.close()


def test_host_and_device_entry_points_agree():
    n = 10_000
    x, codes, g = _data(n, 19)
    ix = _index(n, "bf16")
    ix.add(x, codes)
    q = torch.randn((9, 384), generator=g, device="cuda")
    qc, qm = np.arange(9) % 5, np.full(9, TICKER, np.uint32)
    di, ds = ix.search(q, _i32(qc), _i32(qm), 15)
    hi, hs = ix.search(q.cpu().numpy(), qc, qm, 15)
    assert np.array_equal(di.cpu().numpy(), hi) and np.array_equal(ds.cpu().numpy(), hs)
    ix.close()


def test_empty_index_and_argument_errors():
    from financial_rag_system_b200 import FrsError

    ix = _index(100, "bf16")
    ids, sc = ix.search(np.ones((2, 384), np.float32), [0, 0], [ANY, ANY], 15)
    assert (ids == -1).all() and np.isneginf(sc).all()
    with pytest.raises(ValueError):
        ix.search(np.ones((33, 384), np.float32), [0] * 33, [ANY] * 33, 15)
    with pytest.raises(ValueError):
        ix.search(np.ones((1, 384), np.float32), [0], [ANY], 33)
    with pytest.raises(FrsError):
        ix.add(np.ones((101, 384), np.float32))          # over capacity
    ix.close()


def test_zero_vectors_do_not_poison_results():
    n = 3000
    x, codes, g = _data(n, 23)
    x[7] = 0
    ix = _index(n, "f32")
    ix.add(x, codes)
    q = torch.stack([x[8], torch.zeros(384, device="cuda")])
    _check(ix, q, [0, 0], [ANY, ANY], 15, codes.cpu().numpy().astype(np.uint32))
    ix.close()


def test_two_shards_on_one_gpu_equal_one_index():
    """frs_index_search_local x 2 + the packed cross-shard merge == one index over all rows."""
    from financial_rag_system_b200.index import merge_shards_packed

    n, cut, k = 30_001, 12_345, 15
    x, codes, g = _data(n, 29)
    x[cut + 5] = x[3].clone()
    codes[cut + 5] = codes[3].clone()
    whole = _index(n, "bf16")
    whole.add(x, codes)
    a, b = _index(cut, "bf16", base=0), _index(n - cut, "bf16", base=cut)
    a.add(x[:cut], codes[:cut])
    b.add(x[cut:], codes[cut:])
    q = x[:32] + 0.1 * torch.randn((32, 384), generator=g, device="cuda")
    qc, qm = _i32((codes[:32] & 0x7FFFFFFF).cpu().numpy()), _i32(np.full(32, TICKER, np.uint32))
    packed = torch.empty((2, 2, 32, k), dtype=torch.int64, device="cuda")
    a.search_local(q, qc, qm, k, packed[0, 0].view(torch.float64), packed[0, 1])
    b.search_local(q, qc, qm, k, packed[1, 0].view(torch.float64), packed[1, 1])
    mi, ms = merge_shards_packed(packed, k)
    wi, ws = whole.search(q, qc, qm, k)
    torch.cuda.synchronize()
    assert torch.equal(mi, wi)
    assert torch.equal(ms, ws)
    for ix in (whole, a, b):
        ix.close()


@pytest.mark.parametrize("n_shards,nq,k", [(2, 32, 15), (8, 32, 15), (3, 5, 1), (16, 32, 32), (40, 7, 32)])
def test_cross_shard_merge_equals_the_oracle(n_shards, nq, k):
    """frs_merge_shards (one CTA per query, candidates staged in shared memory; 40 x 32 > 1024 candidates takes the
    path that ranks from global memory) == so.merge_shards: order by (score desc, id asc), duplicated scores across
    shards, empty slots (-1) anywhere, queries with fewer than k valid candidates."""
    from financial_rag_system_b200.index import merge_shards

    rng = np.random.default_rng(100 * n_shards + k)
    sc = np.round(rng.standard_normal((n_shards, nq, k)), 1)          # many equal scores: ties break by id
    ids = rng.permutation(n_shards * nq * k).reshape(n_shards, nq, k).astype(np.int64)
    ids[rng.random(ids.shape) < 0.2] = -1
    ids[:, 0, :] = -1                                                 # a query with no match anywhere
    if nq > 1:
        ids[:, 1, :] = -1
        ids[0, 1, 0] = 7                                              # ... and one with a single match
    sc[ids < 0] = -np.inf
    mi, ms = merge_shards(torch.from_numpy(sc).cuda(), torch.from_numpy(ids).cuda(), k)
    wi, ws = so.merge_shards(list(ids), list(sc), k)
    assert np.array_equal(mi.cpu().numpy(), wi)
    assert np.array_equal(ms.cpu().numpy(), ws.astype(np.float32))


def test_peer_memory_exchange_equals_one_index():
    """The exchange step over peer memory (frs_exchange_*): three shards of ONE process linked by pointer, each
    pushes its block into every shard's gather buffer; every shard's merge equals the single-index search.
    Several batches in a row exercise the double-buffered slots and the monotonic sequence flags."""
    from financial_rag_system_b200.sharded import PeerExchange

    n, cuts, k, nq = 30_001, [0, 9_000, 21_345, 30_001], 15, 32
    x, codes, g = _data(n, 37)
    whole = _index(n, "bf16")
    whole.add(x, codes)
    shards = []
    for r in range(3):
        ix = _index(cuts[r + 1] - cuts[r], "bf16", base=cuts[r])
        ix.add(x[cuts[r]:cuts[r + 1]], codes[cuts[r]:cuts[r + 1]])
        shards.append(ix)
    dev = torch.device("cuda", 0)
    exs = [PeerExchange(dev, 3, r, nq, k, connect=False) for r in range(3)]
    PeerExchange.link(exs)
    for batch in range(5):
        q = x[batch * 32:batch * 32 + nq] + 0.1 * torch.randn((nq, 384), generator=g, device="cuda")
        qc = _i32((codes[batch * 32:batch * 32 + nq] & 0x7FFFFFFF).cpu().numpy())
        qm = _i32(np.full(nq, TICKER, np.uint32))
        for r in range(3):  # all pushes first: the waits of one process would otherwise wait for each other
            if batch % 2 == 0:  # the push fused into the local merge kernel (frs_index_search_push)
                shards[r].search_push(q, qc, qm, k, exs[r])
            else:               # the stand-alone push of an existing block
                loc = torch.empty((2, nq, k), dtype=torch.int64, device="cuda")
                shards[r].search_local(q, qc, qm, k, loc[0].view(torch.float64), loc[1])
                exs[r].push(loc)
        wi, ws = whole.search(q, qc, qm, k)
        for r in range(3):
            mi, ms = exs[r].wait_merge()
            torch.cuda.synchronize()
            assert torch.equal(mi, wi), (batch, r)
            assert torch.equal(ms, ws), (batch, r)
    for e in exs:
        e.close()
    for ix in [whole] + shards:
        ix.close()


def test_concurrent_callers_on_one_index():
    """The reference calls the search from up to 25 threads (main2.py:52-53, 228)."""
    n = 20_000
    x, codes, g = _data(n, 31)
    ix = _index(n, "bf16")
    ix.add(x, codes)
    qs = [torch.randn((8, 384), generator=g, device="cuda").cpu().numpy() for _ in range(6)]
    want = [ix.search(q, np.zeros(8), np.full(8, ANY, np.uint32), 15) for q in qs]
    got = [None] * len(qs)

    def work(i):
        for _ in range(5):
            got[i] = ix.search(qs[i], np.zeros(8), np.full(8, ANY, np.uint32), 15)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(qs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for w, g_ in zip(want, got):
        assert np.array_equal(w[0], g_[0]) and np.array_equal(w[1], g_[1])
    ix.close()


def test_prefilter_error_is_inside_the_stated_bound():
    """The exactness argument needs |tensor-core score - fp64 score| <= eps (3e-5 bf16, 2e-3 tf32)."""
    n = 30_000
    x, codes, g = _data(n, 37, clustered=True)
    for dtype, eps in (("bf16", 3.0e-5), ("f32", 2.0e-3)):
        ix = _index(n, dtype)
        ix.add(x, codes)
        q = x[:32] + 0.05 * torch.randn((32, 384), generator=g, device="cuda")
        approx = ix.debug_scores(q).cpu().numpy().astype(np.float64)
        exact = so.scores_f64(ix.read_rows().cpu().numpy(), ix.last_queries().cpu().numpy())
        err = np.abs(approx - exact).max()
        assert err < eps / 4, (dtype, err)      # measured: 1.6e-7 (bf16), 1.2e-4 (tf32)
        ix.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_full_size_properties_1m(dtype):
    """BASELINE config 2 (1M x 384, 32 queries) through size-independent properties:
    self-retrieval, sortedness, k-prefix consistency, and every reported score re-derived in fp64
    from the stored row."""
    n = 1_000_000
    from financial_rag_system_b200.index import VectorIndex

    g = torch.Generator(device="cuda").manual_seed(41)
    ix = VectorIndex(n, dtype=dtype)
    keep_rows, keep_codes = [], []
    for s in range(0, n, 1 << 18):
        m = min(1 << 18, n - s)
        x = torch.randn((m, 384), generator=g, device="cuda")
        c = torch.randint(0, 500, (m,), generator=g, device="cuda", dtype=torch.int32)
        ix.add(x, c)
        if s == 0:
            keep_rows, keep_codes = x[:32].clone(), c[:32].clone()
    q = keep_rows                                   # 32 queries that ARE corpus rows 0..31
    qc, qm = _i32(keep_codes.cpu().numpy()), _i32(np.full(32, TICKER, np.uint32))
    ids15, sc15 = ix.search(q, qc, qm, 15)
    ids5, sc5 = ix.search(q, qc, qm, 5)
    torch.cuda.synchronize()
    ids15, sc15, ids5, sc5 = ids15.cpu().numpy(), sc15.cpu().numpy(), ids5.cpu().numpy(), sc5.cpu().numpy()
    assert ids15[:, 0].tolist() == list(range(32))                    # self retrieval
    assert np.allclose(sc15[:, 0], 1.0, atol=2e-2 if dtype == "bf16" else 1e-5)
    assert np.all(np.diff(sc15, axis=1) <= 0)                         # sorted
    assert np.array_equal(ids15[:, :5], ids5) and np.array_equal(sc15[:, :5], sc5)   # prefix
    qp = ix.last_queries().cpu().numpy()
    allc = None
    for qi in range(0, 32, 5):                                        # re-derive scores + filter
        rows = torch.cat([ix.read_rows(int(r), 1) for r in ids15[qi]]).cpu().numpy()
        want = rows.astype(np.float64) @ qp[qi].astype(np.float64)
        assert np.allclose(sc15[qi], want, atol=SCORE_ATOL)
    # full oracle check on this size for 4 queries (the oracle needs ~1 s per query here)
    rows_all = ix.read_rows().cpu().numpy()
    codes_all = np.empty(n, np.uint32)
    g2 = torch.Generator(device="cuda").manual_seed(41)
    off = 0
    for s in range(0, n, 1 << 18):
        m = min(1 << 18, n - s)
        torch.randn((m, 384), generator=g2, device="cuda")
        codes_all[off:off + m] = torch.randint(0, 500, (m,), generator=g2, device="cuda", dtype=torch.int32).cpu().numpy()
        off += m
    oi, os_ = so.exact_topk(rows_all, qp[:4], codes_all, keep_codes[:4].cpu().numpy().astype(np.uint32),
                            np.full(4, TICKER, np.uint32), 15)
    assert np.array_equal(ids15[:4], oi)
    assert np.allclose(sc15[:4], os_, atol=SCORE_ATOL)
    ix.close()


def test_export_import_raw_round_trip_is_bit_exact():
    """Persistence: rows exported as they sit in HBM and imported into a fresh index answer every
    query with identical ids and identical score bits (both storage dtypes)."""
    from financial_rag_system_b200.index import VectorIndex

    g = torch.Generator(device="cuda").manual_seed(5)
    n, nq, k = 30000, 32, 15
    x = torch.randn((n, 384), generator=g, device="cuda")
    codes = torch.randint(0, 9, (n,), generator=g, device="cuda", dtype=torch.int32)
    q = x[:nq] + 0.05 * torch.randn((nq, 384), generator=g, device="cuda")
    qc = codes[:nq].clone()
    qm = torch.full((nq,), 0x80FFFFFF - (1 << 32), dtype=torch.int64).to(torch.int32).cuda()
    for dtype in ("bf16", "f32"):
        a = VectorIndex(n, dtype=dtype, device=0)
        a.add(x, codes)
        ids_a, s_a = a.search(q, qc, qm, k)
        rows, c = a.export_raw()
        assert rows.shape == (n, 384) and rows.dtype == (np.uint16 if dtype == "bf16" else np.float32)
        b = VectorIndex(n, dtype=dtype, device=0)
        b.import_raw(rows[:10000], c[:10000])
        b.import_raw(rows[10000:], c[10000:])
        ids_b, s_b = b.search(q, qc, qm, k)
        torch.cuda.synchronize()
        assert torch.equal(ids_a, ids_b) and torch.equal(s_a, s_b)
        with pytest.raises(ValueError):
            b.import_raw(rows.astype(np.float64), c)
        a.close()
        b.close()


def test_restricted_scan_equals_full_scan():
    """frs_index_search_tiles: scanning only the tiles that can hold matches returns the ids and score
    bits of the full scan (ticker-grouped rows, several grid sizes, empty and single-tile lists)."""
    from financial_rag_system_b200.index import VectorIndex

    g = torch.Generator(device="cuda").manual_seed(9)
    n_t, per, nq, k = 40, 3000, 32, 15
    n = n_t * per
    x = torch.randn((n, 384), generator=g, device="cuda")
    codes = (torch.arange(n, device="cuda") // per + 1).to(torch.int32)          # ticker-grouped ingest order
    src = torch.randint(0, n, (nq,), generator=g, device="cuda")
    q = x[src] + 0.05 * torch.randn((nq, 384), generator=g, device="cuda")
    qc = codes[src].clone()
    qc[5:] = qc[5 + (torch.arange(nq - 5, device="cuda") % 3)]                   # 8 distinct tickers in the batch
    qm = torch.full((nq,), 0x80FFFFFF - (1 << 32), dtype=torch.int64).to(torch.int32).cuda()
    for dtype in ("bf16", "f32"):
        ix = VectorIndex(n, dtype=dtype, device=0)
        ix.add(x, codes)
        ids_f, s_f = ix.search(q, qc, qm, k)
        tiles = np.unique(np.concatenate([np.arange((int(c) - 1) * per // 128, ((int(c) - 1) * per + per - 1) // 128 + 1)
                                          for c in qc.cpu().numpy()]))
        assert len(tiles) < (n + 127) // 128 // 3
        for grid in (0, 1, 7, 148):
            ix.set_scan_grid(grid)
            ids_t, s_t = ix.search_tiles(q, qc, qm, k, tiles)
            torch.cuda.synchronize()
            assert torch.equal(ids_t, ids_f) and torch.equal(s_t, s_f), (dtype, grid)
        ix.set_scan_grid(0)
        one = np.array([int(src[0]) // 128])
        ids_1, s_1 = ix.search_tiles(q[:1], qc[:1], qm[:1], k, one)
        assert int(ids_1[0, 0]) == int(src[0])
        ids_0, s_0 = ix.search_tiles(q, qc, qm, k, np.zeros(0, dtype=np.int64))
        torch.cuda.synchronize()
        assert (ids_0 == -1).all() and torch.isinf(s_0).all()
        ix.close()


def test_collection_on_gpu_uses_the_restricted_scan():
    from financial_rag_system_b200.collection import Collection

    rng = np.random.default_rng(6)
    n_t, per = 10, 2000
    vecs = rng.standard_normal((n_t * per, 384)).astype(np.float32)
    payloads = [{"ticker": f"T{t}", "document_type": "10-K", "text": ""} for t in range(n_t) for _ in range(per)]
    seg = Collection(n_t * per, dtype="bf16", device=0)
    seg.upsert(list(range(n_t * per)), vecs, payloads)
    q = vecs[[10, 2500, 19990]] + 0.05
    ts = ["T0", "T1", "T9"]
    a = seg.search(q, ts, 15)
    scanned, total = seg.last_scan_tiles
    assert scanned < total / 2
    seg.segmented = False
    b = seg.search(q, ts, 15)
    assert seg.last_scan_tiles == (total, total)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    seg.close()


@pytest.mark.parametrize("dtype,eps", [("bf16", 3.0e-5), ("f32", 2.0e-3)])
def test_prefilter_bound_holds_for_same_sign_rows_with_scores_near_one(dtype, eps):
    """The adversarial case for an fp32 accumulator that might truncate: every product has the same sign and the
    partial sums grow towards 1 (worst relative rounding), on 1M rows.  The exactness argument only needs
    |tensor-core score - fp64 score| <= eps; the search on the same data (a dense band of near-equal scores:
    list compactions, exact in-scan resolutions, ties) must equal the fp64 ranking."""
    n, nq, k = 1_000_000, 32, 15
    g = torch.Generator(device="cuda").manual_seed(41)
    base = torch.rand((1, 384), generator=g, device="cuda") + 0.5          # all positive
    x = base + 0.02 * torch.rand((n, 384), generator=g, device="cuda")     # all positive, cos(row, row') ~ 0.9999
    x[5000] = x[17]                                                       # exact duplicates: ties
    x[900_000] = x[17]
    codes = torch.zeros((n,), dtype=torch.int32, device="cuda")
    ix = _index(n, dtype)
    ix.add(x, codes)
    q = x[:nq] + 0.01 * torch.rand((nq, 384), generator=g, device="cuda")
    approx = ix.debug_scores(q)[:nq].double()                              # [32, n]
    qp = ix.last_queries()[:nq].double()
    exact = torch.empty((nq, n), dtype=torch.float64, device="cuda")
    for s in range(0, n, 1 << 18):                                         # fp64 products of fp32 values are exact
        exact[:, s:s + (1 << 18)] = (ix.read_rows(s, min(1 << 18, n - s)).double() @ qp.T).T
    err = float((approx - exact).abs().max())
    print(f"same-sign {dtype}: scores in [{float(exact.min()):.5f}, {float(exact.max()):.5f}], max |pre-filter - fp64| = {err:.3e} (eps {eps:.1e})")
    assert err <= eps / 2, err
    ids, sc = ix.search(q, _i32(np.zeros(nq)), _i32(np.full(nq, ANY, np.uint32)), k)
    torch.cuda.synchronize()
    top_s, top_i = torch.topk(exact, 256, dim=1)                           # candidates; final order on the host
    top_s, top_i = top_s.cpu().numpy(), top_i.cpu().numpy()
    for qi in range(nq):
        order = np.lexsort((top_i[qi], -top_s[qi]))[:k]
        assert top_s[qi][order][-1] > top_s[qi].min(), "candidate window too small for this band"
        assert np.array_equal(ids[qi].cpu().numpy(), top_i[qi][order]), qi
        assert np.allclose(sc[qi].cpu().numpy(), top_s[qi][order], atol=SCORE_ATOL)
    st = ix.last_stats()
    print("   scan stats:", st)
    ix.close()

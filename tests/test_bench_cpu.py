"""The reference arm of bench.py (`--impl reference`) runs without a GPU: its JSON line must carry the keys the
driver reads (same metric / unit / config as the GPU arm, `impl`, `cpu_baseline`, an `e2e` object with zero copies)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, FRS_BENCH_ROWS="100000", **extra_env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *args],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = _run({})
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["metric"].startswith("exact top-15 cosine search QPS") and j["unit"] == "queries/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["steps"] == 1 and j["warmup"] == 1
    assert j["config"]["workload"].startswith("100000 x 384") and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "rows per step" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    # under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == []


def test_bench_corpus_is_the_same_on_every_path():
    """bench_data.py: the corpus is a pure function of (seed, row, column); numpy (CPU arm), torch (our arm builds its
    shards with the same integer / IEEE operations on the device) and the C + OpenMP restatement under oracle/ must
    agree bit for bit — both arms search the same rows, tickers and queries."""
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    import bench_data as bd

    x, t = bd.rows_np(123_456, 3000)
    xt, tt = bd.rows_torch(123_456, 3000, torch.device("cpu"))
    assert np.array_equal(x, xt.numpy()) and np.array_equal(t, tt.numpy().astype(np.uint32))
    assert x.dtype == np.float32 and t.dtype == np.uint32 and t.max() < bd.N_TICKERS
    # rows of a different chunking are the same rows (shards of any size see the same corpus)
    x2, t2 = bd.rows_np(124_000, 100)
    assert np.array_equal(x2, x[544:644]) and np.array_equal(t2, t[544:644])
    so = os.path.join(ROOT, "oracle", "_build", "libfrs_synth.so")
    if os.path.exists(so):
        xc, tc = bd.rows_host(123_456, 3000)
        assert np.array_equal(xc, x) and np.array_equal(tc, t)
        xn, _ = bd.rows_host(123_456, 3000, normalise=True)
        assert np.allclose(np.linalg.norm(xn, axis=1), 1.0, atol=1e-6)
    # the query sets: deterministic, the documented shapes
    for kind in bd.QUERY_KINDS:
        q, qt, qm = bd.queries_np(kind)
        assert q.shape == (32, 384) and qt.shape == (32,) and qm.shape == (32,)
        q2, _, _ = bd.queries_np(kind)
        assert np.array_equal(q, q2)
    q, qt, qm = bd.queries_np("self")
    x0, t0 = bd.rows_np(0, 32)
    assert np.array_equal(qt, t0) and (qm == bd.TICKER_MASK).all()
    cos = (q * x0).sum(1) / (np.linalg.norm(q, axis=1) * np.linalg.norm(x0, axis=1))
    assert cos.min() > 0.9   # perturbed copies of rows 0..31
    assert (bd.queries_np("any")[2] == bd.ANY_MASK).all() and (bd.queries_np("one_ticker")[1] == 0).all()

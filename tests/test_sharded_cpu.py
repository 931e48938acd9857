"""Host-side logic of the sharded search under gloo, world_size 2, on CPU: partitioning, the
all-gather layout and the final merge order.  The CUDA local pass and merge kernel are replaced
by the oracle here (that substitution is the point of the injectable hooks); the CUDA versions are
covered by tests/test_search_gpu.py::test_two_shards_on_one_gpu_equal_one_index."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from financial_rag_system_b200.sharded import ShardedIndex, shard_range
from oracle import search_oracle as so


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (a, la), (b, _) in zip(spans, spans[1:]):
                assert a + la == b
            assert spans[-1][0] + spans[-1][1] == total
            assert max(l for _, l in spans) - min(l for _, l in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _corpus():
    rng = np.random.default_rng(11)
    rows = so.store_rows(rng.standard_normal((777, 384)).astype(np.float32), "bf16")
    rows[500] = rows[3]  # duplicate across the shard boundary
    codes = rng.integers(0, 3, 777).astype(np.uint32)
    codes[500] = codes[3]
    q = so.prepare_queries(rows[[3, 400, 700]], "bf16")
    qc = codes[[3, 400, 700]]
    qm = np.array([0x80FFFFFF, 0x80000000, 0x80FFFFFF], dtype=np.uint32)
    return rows, codes, q, qc, qm


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, codes, q, qc, qm = _corpus()
    start, length = shard_range(rows.shape[0], rank, world)

    def local_search(qt, qct, qmt, k, out_s64, out_ids):
        ids, sc = so.exact_topk(rows[start:start + length], qt.numpy(), codes[start:start + length],
                                qct.numpy().astype(np.uint32), qmt.numpy().astype(np.uint32), k, base=start)
        out_s64.copy_(torch.from_numpy(sc))
        out_ids.copy_(torch.from_numpy(ids))

    def merge(packed, k):  # [world, 2, nq, k]: plane 0 = fp64 score bits, plane 1 = ids
        sc = [packed[r, 0].contiguous().view(torch.float64).numpy() for r in range(packed.shape[0])]
        ids = [packed[r, 1].numpy() for r in range(packed.shape[0])]
        mi, ms = so.merge_shards(ids, sc, k)
        return torch.from_numpy(mi), torch.from_numpy(ms.astype(np.float32))

    sh = ShardedIndex(None, rank, world, local_search=local_search, merge=merge, device=torch.device("cpu"))
    ids, scores = sh.search(torch.from_numpy(q), torch.from_numpy(qc.astype(np.int64)), torch.from_numpy(qm.astype(np.int64)), 15)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), ids=ids.numpy(), scores=scores.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_search_equals_single_shard(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rows, codes, q, qc, qm = _corpus()
    want_i, want_s = so.exact_topk(rows, q, codes, qc, qm, 15)
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"r{r}.npz"))
        assert np.array_equal(got["ids"], want_i)
        assert np.allclose(got["scores"], want_s, atol=1e-6)
    assert want_i[0][:2].tolist() == [3, 500]  # the cross-shard duplicate ties in id order


def test_exchange_selection_without_cuda():
    """The peer-memory exchange is a CUDA path: with injected hooks (the gloo test above) or one shard the index
    uses the collective form whatever was asked for; an unknown name is rejected."""
    import pytest

    noop = lambda *a: None  # noqa: E731
    for asked in (None, "auto", "p2p", "nccl"):
        sh = ShardedIndex(object(), 0, 2, local_search=noop, merge=noop, device=torch.device("cpu"), exchange=asked)
        assert sh.exchange == "nccl"
    with pytest.raises(ValueError):
        ShardedIndex(object(), 0, 2, local_search=noop, merge=noop, device=torch.device("cpu"), exchange="mpi")

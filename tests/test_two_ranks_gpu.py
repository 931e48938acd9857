"""The multi-process product path on real GPUs: two ranks (one per GPU, NCCL for set-up only), `ShardedIndex`
over the peer-memory exchange and over the all-gather, synchronous, pipelined and host-buffer forms — every
result against oracle.exact_topk of the concatenated shards.  Needs >= 2 GPUs (skipped otherwise;
`gpurun --gpus 2`).  Also: one rank deliberately late (a host stall) must be waited for, not trapped."""
import os
import socket
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TICKER = 0x80FFFFFF
ANY = 0x80000000
N, NQ, K = 70_001, 32, 15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _corpus():
    rng = np.random.default_rng(23)
    x = rng.standard_normal((N, 384)).astype(np.float32)
    codes = rng.integers(0, 4, N).astype(np.uint32)
    x[50_000] = x[5]       # duplicate across the shard boundary
    codes[50_000] = codes[5]
    q = x[:NQ] + 0.1 * rng.standard_normal((NQ, 384)).astype(np.float32)
    return x, codes, q


def _worker(rank, world, port, out_dir, late_rank):
    import torch.distributed as dist

    from financial_rag_system_b200.index import VectorIndex
    from financial_rag_system_b200.sharded import ShardedIndex, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    x, codes, q = _corpus()
    start, length = shard_range(N, rank, world)
    ix = VectorIndex(length, dtype="bf16", device=rank, base=start)
    ix.add(torch.from_numpy(x[start:start + length]).to(dev), torch.from_numpy(codes[start:start + length].astype(np.int32)).to(dev))
    res = {"rows": ix.read_rows().cpu().numpy(), "start": start}
    qd = torch.from_numpy(q).to(dev)
    qc = torch.from_numpy(codes[:NQ].astype(np.int32)).to(dev)
    qm_t = torch.full((NQ,), TICKER - (1 << 32), dtype=torch.int64).to(torch.int32).to(dev)
    qm_a = torch.full((NQ,), ANY - (1 << 32), dtype=torch.int64).to(torch.int32).to(dev)
    for form in ("p2p", "nccl"):
        sh = ShardedIndex(ix, rank, world, exchange=form)
        assert sh.exchange == form
        i1, s1 = sh.search(qd, qc, qm_t, K)
        pend = [sh.search_async(qd[:nq], qc[:nq], m[:nq], k) for nq, k, m in ((32, 15, qm_t), (7, 16, qm_a), (32, 15, qm_a), (1, 3, qm_t), (32, 15, qm_t))]
        outs = [p.wait() for p in pend]
        i2, s2 = sh.search(qd, qc, qm_a, K)   # a synchronous call after pipelined ones
        torch.cuda.synchronize(dev)
        res[form] = {"sync_t": (i1.cpu().numpy(), s1.cpu().numpy()), "sync_a": (i2.cpu().numpy(), s2.cpu().numpy()),
                     "async": [(i.cpu().numpy(), s.cpu().numpy()) for i, s in outs]}
        if form == "p2p":
            qh, ch = q, codes[:NQ]
            t = [sh.submit_host(qh, ch, np.full(NQ, TICKER, np.uint32), K), sh.submit_host(qh[:9], ch[:9], np.full(9, ANY, np.uint32), 16)]
            res["host"] = [sh.collect_host(x_) for x_ in t]
            # one rank stalls on the host for 0.4 s before it joins the batch: the others wait (bounded), nothing traps
            if rank == late_rank:
                time.sleep(0.4)
            li, ls = sh.search_async(qd, qc, qm_t, K).wait()
            res["late"] = (li.cpu().numpy(), ls.cpu().numpy())
        sh.close()
    if rank == 0:
        res["qprep"] = ix.last_queries().cpu().numpy()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    dist.barrier()
    ix.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_match_the_oracle(tmp_path):
    import torch.multiprocessing as mp

    from oracle import search_oracle as so

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), 1), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npy", allow_pickle=True)[0] for i in range(world)]
    x, codes, q = _corpus()
    rows = np.concatenate([r[0]["rows"], r[1]["rows"]])
    assert np.allclose(rows, so.store_rows(x, "bf16"), rtol=1e-2, atol=1e-4)   # (the norm's summation order differs)
    qp = r[0]["qprep"]   # the prepared queries as the GPU holds them
    qc = codes[:NQ]

    def oracle(nq, k, mask):
        return so.exact_topk(rows, qp[:nq], codes, qc[:nq], np.full(nq, mask, np.uint32), k)

    def same(got, want):
        assert np.array_equal(got[0], want[0])
        fin = np.isfinite(want[1])
        assert np.allclose(got[1][fin], want[1][fin], atol=1e-6)

    shapes = ((32, 15, TICKER), (7, 16, ANY), (32, 15, ANY), (1, 3, TICKER), (32, 15, TICKER))
    for rank in range(world):
        for form in ("p2p", "nccl"):
            same(r[rank][form]["sync_t"], oracle(NQ, K, TICKER))
            same(r[rank][form]["sync_a"], oracle(NQ, K, ANY))
            for got, (nq, k, m) in zip(r[rank][form]["async"], shapes):
                same(got, oracle(nq, k, m))
        same(r[rank]["host"][0], oracle(NQ, K, TICKER))
        same(r[rank]["host"][1], oracle(9, 16, ANY))
        same(r[rank]["late"], oracle(NQ, K, TICKER))
    # both ranks hold the same global result
    assert np.array_equal(r[0]["p2p"]["sync_t"][0], r[1]["p2p"]["sync_t"][0])

#!/usr/bin/env python
"""bench.py — headline benchmark of the retrieval hot path (BASELINE.json `metric`):
QPS of exact top-15 cosine search over 10M x 384-d bf16 chunks, 32-query batches, on 1/2/4/8 B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload search|search_segmented|embed|embed_varlen|rerank|pipeline] [--queries self|any|unrelated|one_ticker|rare]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one 32-query batch searched over the whole corpus (10M rows in total, sharded over the
N ranks: strong scaling, as BASELINE.json quotes the metric).  One JSON line is printed by rank 0.

  value      device-timed whole-job QPS through the pipelined entry point (frs_index_search_async: queries
             resident in HBM, the three kernels of consecutive batches overlap; at N > 1 the cross-shard exchange
             over NVLink peer memory is inside, no NCCL kernel in the timed region)
  e2e        the same batches through the host entry points (numpy in, numpy out; frs_index_search_host_submit /
             _collect, a few batches in flight like the reference's concurrent requests): ONE pinned H2D copy of
             the queries + predicates and ONE D2H copy of ids/scores inside every timed step
  roofline   scan kernel only: algorithmic bytes (rows*768 + rows*4 per launch) / CUDA-event time of the scan
             kernel measured live by library-side events on the stream it is launched on
  breakdown  per-step kernel milliseconds {prep, scan, merge, exchange, gap} from the same events
  per_rank   every rank's scan-kernel time and the fill (start -> first scan kernel) and drain (last scan kernel's
             end -> end) of its timed region: what a short pipelined run pays besides its scans
  parity_checked
             how many of the 32 queries of the timed batch were verified, ids and scores, against an independent
             path (every row's tensor-core score dumped, torch.topk, fp64 rescoring of the candidates, cross-rank
             merge on the host)
  secondary  the other BASELINE.json configs in the same line (1M bf16 / fp32, unfriendly query sets, encoders,
             100M rows at N = 8), each with its own roofline fraction
  cpu_baseline / --impl reference
             the reference's CPU arithmetic for this step (float32 dot + full argsort per query,
             oracle.search_oracle.as_shipped_search — a restatement of qdrant-client's local mode, which is not
             installed here) on the box's host cores: a 1M-row sample beside our line, the FULL 10M rows in
             `--impl reference`; same corpus recipe and seed as our arm (bench_data.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_data as bd  # noqa: E402

TOTAL_ROWS = int(os.environ.get("FRS_BENCH_ROWS", 10_000_000))
NQ, K, DIM = 32, 15, 384
N_TICKERS = bd.N_TICKERS
CPU_SAMPLE_ROWS = 1_000_000
E2E_DEPTH = 3  # host batches in flight in the e2e leg (the library holds 4 staging slots)
METRIC = "exact top-15 cosine search QPS (384-d, 10M chunks, 32-query batches)"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region: NVML when importable (fast),
    else the nvidia-smi query of B200_PROFILING.md."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.stop, self.gpu_index = [], threading.Event(), gpu_index
        self.nv = self.h = self.mx = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            parts = [p for p in vis.split(",") if p]
            phys = int(parts[gpu_index]) if gpu_index < len(parts) and parts[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nv = pynvml
        except Exception:
            self.nv = None
        self.thread = threading.Thread(target=self._loop, daemon=True)

    def sample(self):
        try:
            if self.nv is not None:
                nv = self.nv
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                f = lambda bit: "Active" if (r & bit) else "Not Active"  # noqa: E731
                self.samples.append([str(sm), str(self.mx), f(nv.nvmlClocksThrottleReasonHwSlowdown),
                                     f(nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                                     f(nv.nvmlClocksThrottleReasonSwThermalSlowdown), f(nv.nvmlClocksThrottleReasonSwPowerCap)])
            else:
                r = subprocess.run(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
        except Exception:
            pass

    def _loop(self):
        while not self.stop.is_set():
            self.sample()
            self.stop.wait(0.01 if self.nv is not None else 0.1)

    def start(self):
        self.thread.start()

    def finish(self):
        self.stop.set()
        return self.samples


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max(int(s[1]) for s in samples if s[1].isdigit()) if any(s[1].isdigit() for s in samples) else None
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(samples)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's arithmetic for one step
# --------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core."""
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)  # OpenBLAS / OpenMP pools behind numpy (kept for the process lifetime)
    except Exception:
        pass
    return cores


def cpu_rows(n_rows):
    """The first n_rows rows of the bench corpus, L2-normalised float32 on the host (what a COSINE collection of
    qdrant-client's local mode holds), their tickers, and the `self` query batch."""
    rows, codes = bd.rows_host(0, n_rows, normalise=True)
    q, qt, _ = bd.queries_np("self", NQ)
    return rows, codes, q, qt


def cpu_step_time(rows, codes, queries, q_ticker, steps, warmup):
    from oracle import search_oracle as so

    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for qi in range(queries.shape[0]):
            so.as_shipped_search(rows, queries[qi], codes == q_ticker[qi], K)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.median(times)), times


def cpu_baseline(steps=2, warmup=1):
    cores = use_all_host_threads()
    n = min(CPU_SAMPLE_ROWS, TOTAL_ROWS)
    rows, codes, queries, qt = cpu_rows(n)
    t, _ = cpu_step_time(rows, codes, queries, qt, steps, warmup)
    scale = TOTAL_ROWS / n
    return {
        "value": NQ / (t * scale), "unit": "queries/s", "cores": cores, "kind": "port",
        "sample": f"the first {n} of the {TOTAL_ROWS} rows per 32-query step (float32 dot + full argsort per query: a "
                  f"restatement of qdrant-client local-mode exact search, numpy/OpenBLAS, {cores} threads); step time "
                  f"scaled x{scale:g} (the scan is linear in rows); `--impl reference` runs the full {TOTAL_ROWS} rows",
        "sample_step_s": t,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()
    t0 = time.perf_counter()
    n = int(os.environ.get("FRS_REF_ROWS", TOTAL_ROWS))  # full size by default: no extrapolation
    rows, codes, queries, qt = cpu_rows(n)
    gen_s = time.perf_counter() - t0
    # bound the wall clock: one probe step, then as many warm-up / timed steps as fit ~8 minutes
    probe, _ = cpu_step_time(rows, codes, queries, qt, 1, 0)
    budget_s = float(os.environ.get("FRS_REF_BUDGET_S", 480.0))
    warmup = max(0, min(args.warmup, int(budget_s * 0.15 / probe)))
    steps = max(1, min(args.steps, int((budget_s - (warmup + 1) * probe) / probe)))
    t, times = cpu_step_time(rows, codes, queries, qt, steps, warmup)
    scale = TOTAL_ROWS / n
    qps = NQ / (t * scale)
    line = {
        "impl": "reference", "metric": METRIC,
        "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "probe_steps": 1,
        "ms_per_step": t * scale * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "self"),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{n} of {TOTAL_ROWS} rows per step" + ("" if scale == 1 else f", step time scaled x{scale:g}")
                                   + f"; {steps} timed full-size steps of 32 queries (requested {args.steps}; bounded to ~{budget_s:.0f} s of "
                                   "wall clock); restatement of qdrant-client local-mode exact search (float32 dot + full "
                                   f"argsort per query), numpy/OpenBLAS on {cores} threads; same corpus recipe, seed and queries as "
                                   "the GPU arm (bench_data.py); production Qdrant would answer from an HNSW index, not a scan"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0, "corpus_build_s": gen_s, "step_s": times,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, queries):
    return {
        "workload": f"{TOTAL_ROWS} x 384 bf16 chunk store, 32-query batches, exact cosine top-15, per-query ticker filter",
        "rows_total": TOTAL_ROWS, "rows_per_gpu": TOTAL_ROWS // n_gpus, "dim": DIM, "batch": NQ, "k": K,
        "filter": f"ticker == T, T ~ Zipf(1.1) over {N_TICKERS} tickers (reference main.py:218-223)",
        "corpus": "1024 centroids of norm ~1 + noise of norm ~0.3, L2-normalised on insert; a pure function of (seed 7, row, "
                  "column), generated on the device by our arm and on the host by the CPU arm (bench_data.py)",
        "queries": {"self": "rows 0..31 + 0.02 noise, each with its row's ticker", "any": "rows 0..31 + 0.02 noise, no filter",
                    "unrelated": "32 noise vectors (no near neighbour), tickers ~ Zipf", "one_ticker": "32 noise vectors, all on the hottest ticker",
                    "rare": "32 noise vectors on the 32 rarest tickers"}[queries],
        "parallelism": (f"rows sharded over {n_gpus} GPU(s), one process each; per batch every shard's merge kernel stores its 32x15 "
                        "(score,id) list into every peer's buffer over NVLink peer memory + a sequence flag (no NCCL in the timed region)")
        if n_gpus > 1 else "1 GPU",
        "l2": "inputs larger than L2 (7.7 GB corpus per step vs 126 MB L2)",
    }


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device plumbing shared by the measurements."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # one slice of the host cores per rank: a 150 us GPU step leaves no room for the ranks' threads to migrate
            # over each other (a 1 ms stall of ONE rank's submitting thread stalls all eight: the ranks exchange results
            # every batch)
            try:
                cores = sorted(os.sched_getaffinity(0))
                per = len(cores) // self.world
                if per >= 2:
                    os.sched_setaffinity(0, cores[self.local_rank * per:(self.local_rank + 1) * per])
            except OSError:
                pass
            os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")  # NCCL is set-up / barrier plumbing only
            dist.init_process_group("nccl", device_id=self.dev)
            self._tiny = torch.zeros(1, device=self.dev)
            self._host_group = dist.new_group(backend="gloo")  # a barrier that keeps the waiting ranks' GPUs idle

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def host_barrier(self):
        """CPU-only rendezvous (gloo): an NCCL barrier would park a spinning kernel on every waiting rank's GPU."""
        if self.world > 1:
            self.dist.barrier(group=self._host_group)

    def aligned_start(self):
        """After the host barrier: a device-side rendezvous enqueued right before the start event, so that every
        rank's timed region starts when the LAST rank arrives (host wake-up jitter stays outside)."""
        if self.world > 1:
            self.dist.all_reduce(self._tiny)

    def gather_floats(self, v):
        if self.world == 1:
            return [float(v)]
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def build_shard(cx, total_rows, dtype="bf16", grouped=False):
    """This rank's contiguous shard of the bench corpus, generated on the device chunk by chunk."""
    from financial_rag_system_b200.index import VectorIndex
    from financial_rag_system_b200.sharded import shard_range

    torch = cx.torch
    start, length = shard_range(total_rows, cx.rank, cx.world)
    ix = VectorIndex(length, dtype=dtype, device=cx.local_rank, base=start)
    cent = bd.centroids_torch(cx.dev)
    cdf = torch.from_numpy(bd.zipf_cdf()).to(cx.dev)
    codes_all = torch.empty(length, dtype=torch.int32, device=cx.dev)
    chunk = 1 << 17
    cum = torch.cumsum(torch.tensor(np.diff(np.concatenate([[0.0], bd.zipf_cdf()])) * total_rows, dtype=torch.float64, device=cx.dev), 0)
    for s in range(0, length, chunk):
        m = min(chunk, length - s)
        x, codes = bd.rows_torch(start + s, m, cx.dev, cent=cent, cdf=cdf)
        if grouped:  # rows ingested ticker by ticker (ingest.py:109-177)
            rows_g = torch.arange(start + s, start + s + m, device=cx.dev, dtype=torch.float64)
            codes = torch.searchsorted(cum, rows_g, right=True).clamp_(max=N_TICKERS - 1).to(torch.int32)
        ix.add(x, codes)
        codes_all[s:s + m] = codes
    return ix, codes_all, start, length


def device_queries(cx, kind):
    torch = cx.torch
    q, t, m = bd.queries_np(kind, NQ)
    qd = torch.from_numpy(q).to(cx.dev)
    qc = torch.from_numpy(t.astype(np.int64)).to(torch.int32).to(cx.dev)
    qm = torch.from_numpy(m.astype(np.int64)).to(torch.int32).to(cx.dev)   # wraps to the int32 bit pattern
    return (q, t, m), (qd, qc, qm)


def parity_check(cx, ix, codes_dev, dq, got_ids, got_scores, eps, cand=128):
    """Independent verification of one batch's result, all queries: (1) every row's raw tensor-core score
    (frs_index_debug_scores), ticker predicate applied with torch, torch.topk -> `cand` candidates per query and
    shard; (2) candidates re-scored in fp64 from the stored rows and the prepared queries; (3) the shards' exact
    lists merged on the host by (score desc, id asc).  The candidate set provably contains the exact top-k when the
    cand-th pre-filter score lies more than 2*eps under the k-th exact score — checked, not assumed.
    Returns (queries verified, max |pre-filter - exact| over the candidates and a random row sample)."""
    torch = cx.torch
    qd, qc, qm = dq
    n = len(ix)
    pre = ix.debug_scores(qd)[:NQ]                                 # [32, n] fp32
    qprep = ix.last_queries()[:NQ].to(torch.float64)               # prepared (normalised, rounded) queries
    ok = ((codes_dev[None, :] ^ qc[:, None]) & qm[:, None]) == 0   # tombstone bit is in the mask: live rows only
    masked = torch.where(ok, pre, torch.full_like(pre, float("-inf")))
    c = min(cand, n)
    top_pre, top_row = torch.topk(masked, c, dim=1)
    del masked, ok
    rows_u = torch.unique(top_row.flatten())
    stored = torch.cat([ix.read_rows(int(r), 1) for r in rows_u.tolist()]).to(torch.float64) if rows_u.numel() else torch.zeros((0, DIM), dtype=torch.float64, device=cx.dev)
    pos = torch.searchsorted(rows_u, top_row)
    exact = torch.einsum("qcd,qd->qc", stored[pos], qprep)         # fp64 dot products of the candidates
    exact = torch.where(torch.isinf(top_pre), torch.full_like(exact, float("-inf")), exact)
    err = float((top_pre.double() - exact)[torch.isfinite(exact)].abs().max().item()) if torch.isfinite(exact).any() else 0.0
    # pre-filter error on three random 64k-row windows of this shard as well (not only near the top)
    g = np.random.default_rng(5 + cx.rank)
    for _ in range(3):
        w = min(65536, n)
        r0 = int(g.integers(0, n - w + 1))
        ex = ix.read_rows(r0, w).to(torch.float64) @ qprep.T       # [w, 32]
        err = max(err, float((pre[:, r0:r0 + w].double().T - ex).abs().max().item()))
    gids = top_row.to(torch.int64) + ix.base
    key_s, key_i = exact.cpu().numpy(), gids.cpu().numpy()
    kth_pre = top_pre[:, -1].cpu().numpy() if c == cand else np.full(NQ, -np.inf)
    if cx.world > 1:
        parts = [None] * cx.world
        cx.dist.all_gather_object(parts, (key_s, key_i, kth_pre))
        key_s = np.concatenate([p[0] for p in parts], axis=1)
        key_i = np.concatenate([p[1] for p in parts], axis=1)
        kth_pre = np.max(np.stack([p[2] for p in parts]), axis=0)
    got_i, got_s = got_ids.cpu().numpy(), got_scores.cpu().numpy()
    checked = 0
    for qi in range(NQ):
        order = np.lexsort((key_i[qi], -key_s[qi]))
        s_sorted, i_sorted = key_s[qi][order], key_i[qi][order]
        valid = np.isfinite(s_sorted)
        want_i = np.where(valid[:K], i_sorted[:K], -1)
        want_s = np.where(valid[:K], s_sorted[:K], -np.inf)
        nvalid = int(valid[:K].sum())
        # completeness of the candidate set: everything outside it has pre-filter score <= kth_pre
        if nvalid == K and not (kth_pre[qi] + 2 * eps < want_s[K - 1] or not np.isfinite(kth_pre[qi])):
            raise AssertionError(f"parity check inconclusive for query {qi}: widen `cand` ({kth_pre[qi]} vs {want_s[K - 1]})")
        if not np.array_equal(got_i[qi], want_i):
            raise AssertionError(f"query {qi}: ids differ from the independent path\n got {got_i[qi]}\nwant {want_i}")
        f = np.isfinite(want_s)
        if not np.allclose(got_s[qi][f], want_s[f], atol=1e-6) or np.isfinite(got_s[qi][~f]).any():
            raise AssertionError(f"query {qi}: scores differ from the independent path")
        checked += 1
    return checked, cx.max_over_ranks(err)


def measure_search(cx, ix, sh, hq, dq, steps, warmup, length, row_bytes, tiles_dev=None, e2e=True):
    """The three timed regions on an already built index: value (pipelined, device-resident queries),
    per-kernel events (roofline + breakdown), e2e (host buffers)."""
    torch = cx.torch
    qd, qc, qm = dq
    peer = sh._peer if sh is not None else None

    # result buffers are recycled like a serving loop would (16 batches deep: far more than are ever in flight)
    ring = [(torch.empty((NQ, K), dtype=torch.int64, device=cx.dev), torch.empty((NQ, K), dtype=torch.float32, device=cx.dev))
            for _ in range(16)]
    turn = [0]

    def step():
        if tiles_dev is not None:
            return ix.search_tiles(qd, qc, qm, K, tiles_dev)
        turn[0] = (turn[0] + 1) & 15
        return ix.search_async(qd, qc, qm, K, exchange=peer, out=ring[turn[0]])

    def drain():
        if tiles_dev is None:
            ix.wait(-1)  # the current stream waits for every pipelined batch

    for _ in range(max(warmup, 3)):
        last = step()
    drain()
    cx.barrier()

    # ---- timed region A: value ----
    # (bracket profiling: ONE library-side event before the first scan kernel and one after the last, on the scan's
    # stream — the scan's average launch duration over this very region, without events between a step's kernels)
    if tiles_dev is None:
        ix.set_profiling(3)
    sampler = ClockSampler(cx.local_rank)
    if cx.rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc

    gc.collect()
    gc.disable()   # a collection pause of one rank's submitting thread stalls every rank (they exchange results per batch)
    cx.barrier()
    cx.aligned_start()
    e0.record()
    h0 = time.perf_counter()
    for _ in range(steps):
        last = step()
    host_us = (time.perf_counter() - h0) / steps * 1e6   # host time to enqueue one step (must stay under the GPU's)
    drain()
    e1.record()
    gc.enable()
    if cx.rank == 0:
        sampler.sample()  # the GPU is still working through the queued steps: at least one sample under load
    cx.barrier()
    samples = sampler.finish()
    ms_own = e0.elapsed_time(e1)
    ms = cx.max_over_ranks(ms_own)
    ids, scores = (last.ids, last.scores) if tiles_dev is None else last
    fill_drain = None
    if tiles_dev is None:
        torch.cuda.synchronize()
        fill_drain = ix.read_profile_bracket_rel(e0, e1)
    bracket = ix.read_profile_ex() if tiles_dev is None else None

    # ---- region B: per-kernel events of the same pipelined steps (seven events per step: the breakdown; the events
    # themselves perturb the overlap a little, which is why the roofline uses the bracket of region A) ----
    ix.set_profiling(1)
    cx.barrier()
    for _ in range(min(steps, 200)):
        step()
    drain()
    cx.barrier()
    if os.environ.get("FRS_BENCH_TIMELINE") and cx.rank == 0:
        tl = ix.read_profile_raw(min(steps, 200))
        np.savetxt(os.path.join(ROOT, "gpurun_out", f"timeline_n{cx.world}.txt"), tl, fmt="%.4f",
                   header="prep_start prep_end scan_start scan_end merge_start merge_end exchange_end (ms)")
    prof = ix.read_profile_ex()
    ix.set_profiling(0)
    n = max(prof["n"], 1)
    stats = ix.last_stats()
    launches = stats["launches"]

    out = {"host_enqueue_us_per_step": cx.max_over_ranks(host_us), "ms_per_rank": cx.gather_floats(ms_own), "ms": ms, "ids": ids, "scores": scores, "samples": samples, "launches_per_step": launches, "stats": stats,
           "breakdown": {"prep": prof["prep_ms"] / n, "scan": prof["scan_ms"] / n, "merge": prof["merge_ms"] / n,
                         "exchange": prof["exchange_ms"] / n, "gap": prof["scan_gap_ms"] / max(n - 1, 1),
                         "step": prof["span_ms"] / n}}
    scan_ms = out["breakdown"]["scan"]
    if bracket is not None and bracket["n"] > 0:
        scan_ms = bracket["scan_ms"] / bracket["n"]
        out["scan_ms_source"] = f"CUDA events on the scan stream bracketing the {bracket['n']} scan launches of the timed region (launch gaps included)"
    else:
        out["scan_ms_source"] = "per-launch CUDA events of a separate profiling pass"
    out["scan_ms"] = scan_ms
    if fill_drain is not None:
        # per rank: the scan kernels' average duration in region A, and the parts of the region before the first scan
        # kernel (prep of batch 0) and after the last one (merge + exchange of the last batches, peers that lag)
        out["per_rank"] = {"scan_ms": [round(v, 5) for v in cx.gather_floats(scan_ms)],
                           "fill_ms": [round(v, 4) for v in cx.gather_floats(fill_drain[0])],
                           "drain_ms": [round(v, 4) for v in cx.gather_floats(fill_drain[1])]}
    alg_bytes = (int(tiles_dev.numel()) * 128 if tiles_dev is not None else length) * (row_bytes + 4)
    out["alg_bytes"] = alg_bytes
    out["achieved"] = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    if not e2e:
        return out

    # ---- e2e: host buffers in, host buffers out, one H2D + one D2H per step inside the timed region ----
    q, t, m = hq
    qh, qch, qmh = q, t.astype(np.uint32), m.astype(np.uint32)

    def e2e_run(n_steps, depth):
        inflight, res = [], None
        for _ in range(n_steps):
            inflight.append(ix.submit_host(qh, qch, qmh, K, exchange=peer))
            if len(inflight) >= depth:
                res = ix.collect_host(inflight.pop(0), exchange=peer)
        while inflight:
            res = ix.collect_host(inflight.pop(0), exchange=peer)
        return res

    e2e_run(4, E2E_DEPTH)
    cx.barrier()
    t0 = time.perf_counter()
    hi, hs = e2e_run(steps, E2E_DEPTH)
    e2e_s = time.perf_counter() - t0
    cx.barrier()
    out["e2e_s"] = cx.max_over_ranks(e2e_s)
    assert np.array_equal(hi, ids.cpu().numpy()) and np.array_equal(hs, scores.cpu().numpy()), "host entry point differs from the device one"
    # one batch at a time (the latency a lone request sees)
    cx.barrier()
    t0 = time.perf_counter()
    e2e_run(min(steps, 50), 1)
    out["e2e_sync_s_per_step"] = cx.max_over_ranks((time.perf_counter() - t0) / min(steps, 50))
    cx.barrier()
    return out


def secondary_search(cx, rows, dtype, steps, peak):
    """BASELINE.json configs[1]: rows x 384 on this GPU, bf16 or fp32 storage, same measurement."""
    saved = (cx.rank, cx.world)
    cx.rank, cx.world = 0, 1  # a single-GPU measurement on rank 0's device
    try:
        ix, codes, _, length = build_shard(cx, rows, dtype=dtype)
        hq, dq = device_queries(cx, "self")
        r = measure_search(cx, ix, None, hq, dq, steps, 3, length, DIM * (4 if dtype == "f32" else 2))
        checked, err = parity_check(cx, ix, codes, dq, r["ids"], r["scores"], 2e-3 if dtype == "f32" else 3e-5)
        ix.close()
    finally:
        cx.rank, cx.world = saved
    return {"rows": rows, "dtype": dtype, "qps": NQ * steps / (r["ms"] * 1e-3), "ms_per_step": r["ms"] / steps,
            "e2e_qps": NQ * steps / r["e2e_s"], "scan_ms": r["scan_ms"], "achieved_gbs": r["achieved"],
            "frac": r["achieved"] / peak, "parity_checked": checked, "prefilter_max_err": err,
            "note": "fp32 rows: TF32 tensor-core pre-filter + fp64 rescoring; bound = HBM at 1540 B/row" if dtype == "f32" else "bf16 rows: 772 B/row"}


def single_process_multi_gpu(cx, steps, ref_ids, ref_scores):
    """frs_sharded_* (ONE process driving all N GPUs, the form the reference's single server process would hold): the
    same 10M-row corpus placed block-cyclically over the N GPUs by rank 0 alone, the same batch through the host entry
    point; ids / scores must equal the multi-process result.  The other ranks idle at a barrier meanwhile."""
    from financial_rag_system_b200.multigpu import MultiGpuIndex

    torch = cx.torch
    mg = MultiGpuIndex(TOTAL_ROWS, dtype="bf16", devices=list(range(cx.world)))
    cent = bd.centroids_torch(cx.dev)
    cdf = torch.from_numpy(bd.zipf_cdf()).to(cx.dev)
    t0 = time.perf_counter()
    chunk = 1 << 18
    for s in range(0, TOTAL_ROWS, chunk):
        m = min(chunk, TOTAL_ROWS - s)
        x, codes = bd.rows_torch(s, m, cx.dev, cent=cent, cdf=cdf)
        mg.add_device(x, codes)   # block by block to the shards' GPUs over NVLink, no host round trip
    build_s = time.perf_counter() - t0
    q, t, m = bd.queries_np("self", NQ)
    qc, qm = t.astype(np.uint32), m.astype(np.uint32)

    def run(n_steps, depth):
        inflight, res = [], None
        for _ in range(n_steps):
            inflight.append(mg.submit(q, qc, qm, K))
            if len(inflight) >= depth:
                res = mg.collect(inflight.pop(0))
        while inflight:
            res = mg.collect(inflight.pop(0))
        return res

    run(5, E2E_DEPTH)
    t0 = time.perf_counter()
    ids, scores = run(steps, E2E_DEPTH)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    run(min(steps, 50), 1)
    dt1 = (time.perf_counter() - t0) / min(steps, 50)
    same = bool(np.array_equal(ids, ref_ids) and np.array_equal(scores, ref_scores))
    mg.close()
    return {"entry_point": "frs_sharded_search_host_submit/_collect (one process, one host thread, all GPUs)", "n_gpus": cx.world,
            "rows": TOTAL_ROWS, "e2e_qps": NQ * steps / dt, "in_flight": E2E_DEPTH, "one_at_a_time_qps": NQ / dt1,
            "equals_multi_process_result": same, "build_s": build_s,
            "note": "host-timed (numpy in, numpy out); the host thread issues every GPU's copies and kernels itself"}


def run_ours(args):
    cx = Ctx(args)
    torch = cx.torch
    from financial_rag_system_b200.sharded import ShardedIndex

    grouped = args.workload == "search_segmented"
    if grouped and cx.world > 1:
        raise SystemExit("--workload search_segmented is a single-GPU measurement")
    ix, codes_dev, start, length = build_shard(cx, TOTAL_ROWS, grouped=grouped)
    sh = ShardedIndex(ix, cx.rank, cx.world, exchange="p2p") if cx.world > 1 else None
    if os.environ.get("FRS_SCAN_GRID"):  # experiment knobs
        ix.set_scan_grid(int(os.environ["FRS_SCAN_GRID"]))
    if os.environ.get("FRS_PIPE_RESERVE"):
        ix.set_pipeline_reserve(int(os.environ["FRS_PIPE_RESERVE"]))
    if os.environ.get("FRS_SCAN_STREAMS"):
        ix.set_scan_streams(int(os.environ["FRS_SCAN_STREAMS"]))
    hq, dq = device_queries(cx, args.queries)
    tiles_dev, n_tiles_total = None, (length + 127) // 128
    if grouped:
        # the tiles that hold rows of the batch's tickers (what Collection._batch_tiles computes on the host)
        cum_h = np.concatenate([[0.0], bd.zipf_cdf() * TOTAL_ROWS])
        sets = [np.arange(int(cum_h[int(c)]) // 128, min(int(np.ceil(cum_h[int(c) + 1])), length - 1) // 128 + 1)
                for c in np.unique(hq[1])]
        tiles_dev = torch.from_numpy(np.unique(np.concatenate(sets)).astype(np.int32)).to(cx.dev)
        full_ids, _ = ix.search(*dq, K)
        seg_ids, _ = ix.search_tiles(*dq, K, tiles_dev)
        assert torch.equal(full_ids, seg_ids), "restricted scan must return the ids of the full scan"

    r = measure_search(cx, ix, sh, hq, dq, args.steps, args.warmup, length, DIM * 2, tiles_dev)
    if sh is not None:
        si, ss = sh.search(*dq, K)  # the synchronous sharded form answers identically
        assert torch.equal(si, r["ids"]) and torch.equal(ss, r["scores"]), "synchronous and pipelined sharded search differ"
    checked, pre_err = parity_check(cx, ix, codes_dev, dq, r["ids"], r["scores"], 3e-5)
    peak, peak_src = peak_hbm()
    ms, steps = r["ms"], args.steps

    # ---- the other configs, same line ----
    secondary = {}
    if not args.no_secondary and not grouped and os.environ.get("FRS_BENCH_ROWS") is None:
        sec_steps = max(10, min(args.steps, 50))
        qsets = {}
        for kind in bd.QUERY_KINDS:
            if kind == args.queries:
                continue
            hq2, dq2 = device_queries(cx, kind)
            r2 = measure_search(cx, ix, sh, hq2, dq2, sec_steps, 3, length, DIM * 2, e2e=False)
            c2, _ = parity_check(cx, ix, codes_dev, dq2, r2["ids"], r2["scores"], 3e-5)
            qsets[kind] = {"qps": NQ * sec_steps / (r2["ms"] * 1e-3), "scan_ms": r2["scan_ms"],
                           "frac": r2["achieved"] / peak, "parity_checked": c2,
                           "stats": {k: r2["stats"][k] for k in ("appended", "compactions", "resolutions", "rescored")}}
        secondary["queries"] = qsets
        if cx.world == 8:
            # BASELINE.json configs[3] / the north-star target: 100M rows over 8 GPUs
            sh.close()   # (the exchange first: it drains the index's pipelined searches)
            sh = None
            ix.close()
            ix = None
            torch.cuda.empty_cache()
            big = 100_000_000
            ixb, codes_b, _, len_b = build_shard(cx, big)
            shb = ShardedIndex(ixb, cx.rank, cx.world, exchange="p2p")
            hqb, dqb = device_queries(cx, "self")
            rb = measure_search(cx, ixb, shb, hqb, dqb, sec_steps, 3, len_b, DIM * 2)
            cb_, eb = parity_check(cx, ixb, codes_b, dqb, rb["ids"], rb["scores"], 3e-5)
            agg = rb["alg_bytes"] * cx.world / (rb["ms"] / sec_steps * 1e-3) / 1e9
            secondary["rows_100m"] = {"rows": big, "qps": NQ * sec_steps / (rb["ms"] * 1e-3), "ms_per_step": rb["ms"] / sec_steps,
                                      "e2e_qps": NQ * sec_steps / rb["e2e_s"], "scan_ms": rb["scan_ms"],
                                      "scan_frac": rb["achieved"] / peak, "aggregate_gbs_whole_step": agg,
                                      "aggregate_frac_whole_step": agg / (peak * cx.world), "target_frac": 0.80,
                                      "parity_checked": cb_, "prefilter_max_err": eb, "breakdown_ms": rb["breakdown"]}
            shb.close()
            ixb.close()
        if cx.world == 1:
            ix.close()
            ix = None
            torch.cuda.empty_cache()
            secondary["search_1m_bf16"] = secondary_search(cx, 1_000_000, "bf16", sec_steps, peak)
            secondary["search_1m_f32"] = secondary_search(cx, 1_000_000, "f32", sec_steps, peak)
        if cx.world == 1 or os.environ.get("FRS_BENCH_ENCODERS_ALL_N"):
            import bench_encoders

            for w in ("embed", "embed_varlen", "rerank", "pipeline"):
                try:
                    secondary[w] = bench_encoders.measure_compact(cx, w, ClockSampler, summarize_clocks)
                except Exception as e:  # noqa: BLE001 - a secondary measurement must not cost the headline line
                    secondary[w] = {"error": f"{type(e).__name__}: {e}"}
        elif not args.no_secondary:
            import bench_encoders

            try:
                secondary["embed"] = bench_encoders.measure_compact(cx, "embed", ClockSampler, summarize_clocks)
            except Exception as e:  # noqa: BLE001
                secondary["embed"] = {"error": f"{type(e).__name__}: {e}"}
        if cx.world > 1:
            # the single-process multi-GPU entry points on the same GPUs: rank 0 drives all of them, the others wait
            if sh is not None:
                sh.close()
                sh = None
            if ix is not None:
                ix.close()
                ix = None
            torch.cuda.empty_cache()
            cx.barrier()
            cx.host_barrier()
            if cx.rank == 0:
                try:
                    secondary["single_process"] = single_process_multi_gpu(cx, sec_steps, r["ids"].cpu().numpy(), r["scores"].cpu().numpy())
                except Exception as e:  # noqa: BLE001
                    secondary["single_process"] = {"error": f"{type(e).__name__}: {e}"}
            cx.host_barrier()  # (the other ranks wait on the host: their GPUs are rank 0's to use meanwhile)

    if cx.rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if int(tj.get("rows", -1)) == length:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        cb = cpu_baseline() if cx.world == 1 and not os.environ.get("FRS_BENCH_NO_CPU") else None
        cfg = workload_config(cx.world, args.queries)
        if grouped:
            cfg["layout"] = (f"rows grouped by ticker (ingest order); the batch's {int(np.unique(hq[1]).size)} tickers occupy "
                             f"{int(tiles_dev.numel())} of {n_tiles_total} tiles; e2e is the full-scan host entry point")
        line = {
            "metric": METRIC,
            "value": NQ * steps / (ms * 1e-3), "unit": "queries/s", "n_gpus": cx.world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": NQ * steps / r["e2e_s"], "unit": "queries/s",
                    "h2d_bytes_per_step": NQ * DIM * 4 + 2 * NQ * 4, "d2h_bytes_per_step": NQ * K * (4 + 8),
                    "in_flight": E2E_DEPTH, "one_at_a_time_qps": NQ / r["e2e_sync_s_per_step"],
                    "note": f"host buffers through frs_index_search_host_submit/_collect, {E2E_DEPTH} batches in flight (the reference "
                            "serves up to 25 concurrent requests, main2.py:52-53); one_at_a_time_qps = a lone caller"},
            "gpu_launches": r["launches_per_step"] * steps,
            "roofline": {"bound": "hbm", "achieved": r["achieved"], "peak": peak, "unit": "GB/s", "frac": r["achieved"] / peak,
                         "traffic": traffic, "traffic_source": "ncu dram__bytes_read+write of one launch, builder's capture "
                         "(profiles/scan_traffic.json), not re-measured in this run" if traffic else None,
                         "kernel": "scan_kernel<bf16>", "kernel_ms": r["scan_ms"], "kernel_ms_source": r["scan_ms_source"],
                         "algorithmic_bytes_per_launch": r["alg_bytes"], "peak_source": peak_src,
                         "whole_step_frac": r["alg_bytes"] / (ms / steps * 1e-3) / 1e9 / peak,
                         "frac_of_nominal_8tbs": r["achieved"] / 8000.0},
            "breakdown_ms": {k: round(v, 5) for k, v in r["breakdown"].items()},
            "host_enqueue_us_per_step": round(r["host_enqueue_us_per_step"], 1),
            "timed_region_ms_per_rank": [round(v, 4) for v in r["ms_per_rank"]],
            "per_rank": r.get("per_rank"),
            "scan_stats_last_step": {k: r["stats"][k] for k in ("appended", "compactions", "resolutions", "rescored", "grid")},
            "parity_checked": checked, "prefilter_max_err": pre_err, "prefilter_eps": 3e-5,
            "comm": ("peer-memory stores + flags inside the merge kernel (csrc/exchange.cu); NCCL is used for process-group set-up and "
                     "barriers only, none of its kernels runs inside the timed region") if cx.world > 1 else None,
            "clocks": summarize_clocks(r["samples"]),
            "secondary": secondary,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
            if not args.no_secondary and not grouped:
                import bench_encoders

                try:
                    line["cpu_baseline"]["pipeline"] = bench_encoders.cpu_pipeline_config1()
                except Exception as e:  # noqa: BLE001
                    line["cpu_baseline"]["pipeline"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if sh is not None:
        sh.close()
    if ix is not None:
        ix.close()
    if cx.world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["search", "search_segmented", "embed", "embed_varlen", "rerank", "pipeline"], default="search",
                    help="search = the headline metric (default); the others are BASELINE.json configs[2] / configs[4], see bench_encoders.py")
    ap.add_argument("--queries", choices=list(bd.QUERY_KINDS), default="self", help="query set of the headline measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configs (quick runs, ncu captures)")
    args = ap.parse_args()
    if args.workload not in ("search", "search_segmented"):
        import bench_encoders

        if args.impl == "reference":
            bench_encoders.run_reference(args)
        else:
            bench_encoders.run_ours(args, ClockSampler, summarize_clocks, Ctx)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Encoder / pipeline workloads of bench.py (`--workload embed | rerank | pipeline`): BASELINE.json
configs[2] (bge-small ingest embedding throughput, chunks x 512 tokens, 1-8 GPUs) and configs[4]
(32-query dynamic batch: embed -> search -> 15 candidates each through the MiniLM-L-6 cross-encoder).

Same JSON contract as the headline search line (value / e2e / roofline / cpu_baseline / clocks).
  value     device-timed: token ids already in HBM, the encoder forward only
  e2e       the public API with HOST inputs: `Embedder.encode(texts)` / `Reranker.predict(pairs)` /
            `Retriever.retrieve_batch(queries, tickers)` — WordPiece tokenisation on the host cores,
            pinned H2D of ids, forward, D2H of embeddings / logits inside the timed region
  roofline  tensor-core bound: algorithmic flops (linear 2*384*(1152+384+1536+1536) per token-layer +
            attention 4*384*S per token-layer) / CUDA-event time, against the measured
            sustained bf16 peak; per-kernel-class milliseconds from library-side events
  --impl reference / cpu_baseline
            transformers' BertModel / BertForSequenceClassification (what sentence-transformers runs in
            the reference) in fp32 on all host cores over a bounded sample.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
CHUNKS_PER_STEP = 148     # x 512 tokens = 75776 packed tokens per pass and per GPU: one chunk per SM, so every kernel's tile
                          # count is a multiple of its grid (592 row tiles = 4 x 148; 128 chunks left the LayerNorm GEMMs'
                          # fourth round of 256-row pair tiles 46 % empty)
SEQ = 512
NQ, LIMIT, TOP_K = 32, 15, 5
PIPE_ROWS = int(os.environ.get("FRS_PIPE_ROWS", 100_000))
VARLEN_CHUNKS = 320  # ~73.6k tokens per step at a mean of 230: the same pass size as 148 x 512


def flops(layers, lens):
    lens = np.asarray(lens, dtype=np.float64)
    return layers * (2 * 384 * (1152 + 384 + 1536 + 1536) * lens.sum() + 4 * 384 * (lens ** 2).sum())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (MEASURED_PEAKS.json bf16_tflops_sustained, cuBLAS 8192^3 back to back)"
    return 1400.0, "fallback (B200_PROFILING.md)"


def metric_name(workload):
    return {"embed": "bge-small ingest embedding throughput (chunks x 512 tokens)",
            "embed_varlen": "bge-small ingest embedding throughput (chunks of ~230 tokens, ragged)",
            "rerank": "MiniLM-L-6 cross-encoder rerank throughput (query, chunk) pairs",
            "pipeline": "retrieve+rerank: 32-query dynamic batch, top-15 each reranked to top-5"}[workload]


def unit_name(workload):
    return {"embed": "chunks/s", "embed_varlen": "chunks/s", "rerank": "pairs/s", "pipeline": "queries/s"}[workload]


def config(workload, world):
    if workload == "embed":
        return {"workload": f"bge-small-en-v1.5 shape (12 layers, 33.4M params, seeded synthetic weights), {CHUNKS_PER_STEP} chunks x {SEQ} "
                            f"tokens per step and per GPU (one per SM), CLS pooling + L2 normalise", "chunks_per_step_per_gpu": CHUNKS_PER_STEP,
                "seq_len": SEQ, "parallelism": f"{world} GPU(s), independent chunks, weights replicated, no collective",
                "l2": "activations of one pass (~580 MB) exceed the 126 MB L2"}
    if workload == "embed_varlen":
        return {"workload": f"bge-small-en-v1.5 shape, {VARLEN_CHUNKS} chunks per step and per GPU with lengths ~ N(230, 40) clipped to "
                            "[16, 512] (SURVEY 8d config 3's realistic run: ~1000-char chunks, ingest.py:25-26), packed without padding, "
                            "CLS pooling + L2 normalise", "chunks_per_step_per_gpu": VARLEN_CHUNKS,
                "parallelism": f"{world} GPU(s), independent chunks, weights replicated, no collective",
                "l2": "activations of one pass (~570 MB) exceed the 126 MB L2"}
    if workload == "rerank":
        return {"workload": f"ms-marco-MiniLM-L-6-v2 shape (6 layers, 22.7M params, seeded synthetic weights), {NQ}x{LIMIT} = {NQ * LIMIT} "
                            "(query, ~1000-char chunk) pairs per step, raw logits", "pairs_per_step_per_gpu": NQ * LIMIT,
                "parallelism": f"{world} GPU(s), independent pairs", "l2": "activations of one pass (~1 GB) exceed the 126 MB L2"}
    return {"workload": f"{NQ} synthetic analyst questions -> embed -> exact cosine top-{LIMIT} with per-query ticker filter over "
                        f"{PIPE_ROWS} embedded synthetic SEC chunks -> {NQ * LIMIT} pairs through the cross-encoder -> top-{TOP_K}",
            "rows": PIPE_ROWS, "batch": NQ, "parallelism": f"{world} GPU(s): replicas (each rank serves its own batches)",
            "l2": "cross-encoder activations (~1 GB per pass) exceed the 126 MB L2"}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: transformers on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_encode_rate(workload, seconds_budget=20.0):
    import torch

    from financial_rag_system_b200.checkpoint import BGE_SMALL, MINILM_L6_CE, synthetic_checkpoint
    from oracle import encoder_oracle as eo

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    if workload == "embed":
        shape, seed, n, length = BGE_SMALL, 1234, 8, SEQ
    else:
        shape, seed, n, length = MINILM_L6_CE, 4321, 15, 320
    model = eo.hf_model(shape, synthetic_checkpoint(shape, seed))
    cu = (np.arange(n + 1) * length).astype(np.int32)
    ids = rng.integers(1000, 30522, size=n * length).astype(np.int32)
    tts = np.zeros_like(ids)
    run = (lambda: eo.hf_embed(model, ids, cu)) if workload == "embed" else (lambda: eo.hf_score_pairs(model, ids, tts, cu))
    run()
    t0, reps = time.perf_counter(), 0
    while True:
        run()
        reps += 1
        if time.perf_counter() - t0 > seconds_budget or reps >= 20:
            break
    dt = (time.perf_counter() - t0) / reps
    what = f"{n} sequences x {length} tokens per call, {reps} calls" + ("" if workload == "embed" else " (15 pairs = one reference rerank call, main.py:245)")
    return n / dt, {"value": n / dt, "unit": "chunks/s" if workload == "embed" else "pairs/s", "cores": cores, "kind": "reference",
                    "sample": f"transformers {shape.layers}-layer BERT fp32 on {cores} host threads, {what}; "
                              "sentence-transformers itself is not installed (it calls this module)"}


def cpu_pipeline_config1(rounds: int = 2):
    """BASELINE.json configs[0], the reference's own CPU-runnable case: 10k synthetic SEC chunks, bge-small-shaped
    query embedding, exact top-15 with ticker filter (numpy restatement of qdrant-client :memory:), MiniLM-L-6-shaped
    cross-encoder rerank to top-5 — for 10 CONCURRENT queries (load_testing.py:173-200 drives 10 users; main.py:211-247 is
    the per-request path: embed_query -> retrieve_from_qdrant -> rerank_documents), transformers fp32 on the host cores.
    The 10k stored vectors are not embedded on the CPU (that would take ~10 minutes and is ingest, not the request
    path): 64 chunks are embedded for real, the rest are seeded unit vectors; every chunk has real text for the reranker."""
    import threading
    from concurrent.futures import ThreadPoolExecutor

    import torch

    from financial_rag_system_b200 import synth
    from financial_rag_system_b200.checkpoint import BGE_SMALL, MINILM_L6_CE, synthetic_checkpoint
    from financial_rag_system_b200.tokenizer import WordPiece
    from oracle import encoder_oracle as eo
    from oracle import search_oracle as so

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    n_rows, n_users = 10_000, 10
    tok = WordPiece.synthetic()   # the `tokenizers` WordPiece the reference's models use, synthetic vocabulary
    emb = eo.hf_model(BGE_SMALL, synthetic_checkpoint(BGE_SMALL, 1234))
    ce = eo.hf_model(MINILM_L6_CE, synthetic_checkpoint(MINILM_L6_CE, 4321))
    _, texts, payloads = synth.make_chunks(n_rows, seed=1234)
    rng = np.random.default_rng(1234)
    vecs = so.l2_normalize_f32(rng.standard_normal((n_rows, 384)).astype(np.float32))
    ids64, cu64 = tok.pack_texts(texts[:64])
    vecs[:64] = eo.hf_embed(emb, ids64, cu64)
    tick = np.array([p["ticker"] for p in payloads])
    queries, _ = synth.make_queries(n_users, seed=21)
    q_tickers = [payloads[int(j)]["ticker"] for j in rng.integers(0, n_rows, n_users)]
    stage = {"embed": [], "search": [], "rerank": []}
    lock = threading.Lock()

    def one_request(i):
        t0 = time.perf_counter()
        qi, qcu = tok.pack_texts([queries[i]])
        v = eo.hf_embed(emb, qi, qcu)[0]                                   # main.py:211-213
        t1 = time.perf_counter()
        top, _ = so.as_shipped_search(vecs, v, tick == q_tickers[i], 15)   # main.py:215-239
        t2 = time.perf_counter()
        pairs = [[queries[i], texts[int(r)]] for r in top]
        pi, pt, pcu = tok.pack_pairs(pairs)
        scores = eo.hf_score_pairs(ce, pi, pt, pcu)                        # main.py:241-247
        idx = np.argsort(scores)[::-1][:5]
        t3 = time.perf_counter()
        with lock:
            stage["embed"].append(t1 - t0)
            stage["search"].append(t2 - t1)
            stage["rerank"].append(t3 - t2)
        return idx

    walls = []
    for rd in range(rounds + 1):
        for v_ in stage.values():
            v_.clear()
        t0 = time.perf_counter()
        with ThreadPoolExecutor(n_users) as ex:
            list(ex.map(one_request, range(n_users)))
        if rd > 0:
            walls.append(time.perf_counter() - t0)
    wall = float(np.median(walls))
    return {"value": n_users / wall, "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": f"{n_users} concurrent requests (threads) over {n_rows} chunks, {rounds} timed rounds after one warm-up; transformers "
                      f"BertModel-12 / BertForSequenceClassification-6 fp32 (seeded synthetic weights of the reference checkpoints' shapes) on "
                      f"{cores} host threads; numpy restatement of qdrant-client exact search; sentence-transformers / qdrant-client "
                      "themselves are not installed in this image",
            "wall_s_per_round": wall,
            "stage_ms_median_last_round": {k: float(np.median(v_) * 1e3) for k, v_ in stage.items()}}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t0 = time.perf_counter()
    w = args.workload
    if w == "pipeline":
        cb = cpu_pipeline_config1()   # the stated config: 10k chunks, 10 concurrent requests, embed -> search -> rerank
        value = cb["value"]
    else:
        value, cb = cpu_encode_rate("embed" if w == "embed_varlen" else w)
    line = {"impl": "reference", "metric": metric_name(w), "value": value, "unit": unit_name(w), "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(w, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": unit_name(w), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def measure(cx, w, steps, warmup, ClockSampler, summarize_clocks, with_cpu=True):
    """One workload on the ranks of `cx` (bench.Ctx); returns the JSON line (rank 0) or None."""
    import torch

    from financial_rag_system_b200 import synth
    from financial_rag_system_b200.encoder import Embedder, Reranker
    from financial_rag_system_b200.tokenizer import WordPiece

    rank, world, local_rank, dev = cx.rank, cx.world, cx.local_rank, cx.dev
    dist = cx.dist
    barrier = cx.barrier
    tok = WordPiece.synthetic()
    rng = np.random.default_rng(100 + rank)

    class _A:
        pass

    args = _A()
    args.steps, args.warmup = steps, warmup

    retr = None
    if w in ("embed", "embed_varlen"):
        model = Embedder(device=local_rank, max_tokens=CHUNKS_PER_STEP * SEQ, tokenizer=tok)
        bert, layers = model.bert, 12
        if w == "embed":
            lens = [SEQ] * CHUNKS_PER_STEP
        else:
            lens = np.clip(np.rint(np.random.default_rng(230).normal(230.0, 40.0, VARLEN_CHUNKS)), 16, SEQ).astype(np.int64).tolist()
            assert sum(lens) <= CHUNKS_PER_STEP * SEQ
        cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        ids = torch.from_numpy(rng.integers(1000, 30522, size=int(cu[-1])).astype(np.int32)).to(dev)
        step = lambda: bert.embed_device(ids, cu)  # noqa: E731
        units = len(lens)
        # host-side inputs of the e2e leg: real chunk texts (~300 tokens each, the corpus the ingest path sees)
        _, texts, _ = synth.make_chunks(2048, seed=7 + rank)
        e2e_units = len(texts)
        e2e_call = lambda: model.encode(texts)  # noqa: E731
        tok_ids, tok_cu = tok.pack_texts(texts)
        h2d, d2h = int(tok_ids.nbytes), e2e_units * 384 * 4
    elif w == "rerank":
        model = Reranker(device=local_rank, max_tokens=NQ * LIMIT * 360, tokenizer=tok)
        bert, layers = model.bert, 6
        _, texts, _ = synth.make_chunks(NQ * LIMIT, seed=9 + rank)
        qs, _ = synth.make_queries(NQ, seed=5 + rank)
        pairs = [[qs[i // LIMIT], texts[i]] for i in range(NQ * LIMIT)]
        pi, pt, cu = tok.pack_pairs(pairs)
        lens = np.diff(cu).tolist()
        ids, tts = torch.from_numpy(pi).to(dev), torch.from_numpy(pt).to(dev)
        step = lambda: bert.score_device(ids, tts, cu)  # noqa: E731
        units = e2e_units = NQ * LIMIT
        e2e_call = lambda: model.predict(pairs)  # noqa: E731
        h2d, d2h = int(pi.nbytes + pt.nbytes), NQ * LIMIT * 4
    else:
        from financial_rag_system_b200.api import Retriever
        from financial_rag_system_b200.collection import Collection

        emb = Embedder(device=local_rank, max_tokens=65536, tokenizer=tok)
        rr = Reranker(device=local_rank, max_tokens=NQ * LIMIT * 360, tokenizer=tok)
        n_text = min(PIPE_ROWS, 20_000)  # distinct chunk texts; the rest of the rows are perturbed copies of their vectors
        cids, texts, payloads = synth.make_chunks(n_text, seed=11)
        col = Collection(PIPE_ROWS, dtype="bf16", device=local_rank)
        vecs = np.concatenate([emb.encode(texts[s:s + 2048]) for s in range(0, n_text, 2048)])
        col.upsert(cids, vecs, payloads)
        g = np.random.default_rng(3)
        while len(col) < PIPE_ROWS:
            m = min(n_text, PIPE_ROWS - len(col))
            src = g.integers(0, n_text, m)
            col.upsert([f"dup{len(col) + i}" for i in range(m)], vecs[src] + 0.05 * g.standard_normal((m, 384)).astype(np.float32),
                       [payloads[j] for j in src])
        retr = Retriever(col, emb, rr)
        qs, _ = synth.make_queries(NQ, seed=21 + rank)
        ts = [payloads[int(j)]["ticker"] for j in g.integers(0, n_text, NQ)]
        # device-timed leg: the three GPU passes on pre-tokenised, device-resident inputs
        q_ids, q_cu = tok.pack_texts(qs)
        q_ids_d = torch.from_numpy(q_ids).to(dev)
        pred = [col.predicate(t) for t in ts]
        qc = torch.tensor([p[0] for p in pred], dtype=torch.int64).to(torch.int32).to(dev)
        qm = torch.tensor([p[1] - (1 << 32) if p[1] >= (1 << 31) else p[1] for p in pred], dtype=torch.int64).to(torch.int32).to(dev)
        hits0 = retr.retrieve_batch(qs, ts, TOP_K)
        ids0, _ = retr.search(retr.embed(qs), ts, LIMIT)
        pairs = [[qs[i], payloads_text] for i in range(NQ) for payloads_text in
                 [col.payloads[int(r_)].get("text", "") for r_ in ids0[i] if r_ >= 0]]
        pi, pt, pcu = tok.pack_pairs(pairs)
        pi_d, pt_d = torch.from_numpy(pi).to(dev), torch.from_numpy(pt).to(dev)
        lens = np.diff(q_cu).tolist()
        ce_lens = np.diff(pcu).tolist()

        def step():
            v = emb.bert.embed_device(q_ids_d, q_cu)
            col.index.search(v, qc, qm, LIMIT)
            return rr.bert.score_device(pi_d, pt_d, pcu)

        units = e2e_units = NQ
        e2e_call = lambda: retr.retrieve_batch(qs, ts, TOP_K)  # noqa: E731
        h2d = int(q_ids.nbytes + pi.nbytes + pt.nbytes + NQ * 384 * 4 + NQ * 8)
        d2h = int(NQ * 384 * 4 + NQ * LIMIT * 12 + len(pairs) * 4)
        bert, layers = rr.bert, 6
        assert len(hits0) == NQ

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    if rank == 0:
        sampler.sample()
    barrier()
    samples = sampler.finish()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # per-kernel-class events (one profiled pass of the dominant encoder)
    bert.set_profiling(True)
    step()
    torch.cuda.synchronize(dev)
    prof = bert.read_profile()
    bert.set_profiling(False)

    # e2e through the public API, host strings in / host arrays out
    e2e_steps = max(3, min(args.steps, 20))
    e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank == 0:
        peak, peak_src = peaks()
        if w == "pipeline":
            fl = flops(12, lens) + flops(6, ce_lens)
            kernel, kernel_fl = "cross-encoder forward (6 layers)", flops(6, ce_lens)
        else:
            fl = kernel_fl = flops(layers, lens)
            kernel = "encoder forward (all kernels of one pass)"
        # per-class times come from a separate profiled pass (events between the kernels switch off the overlap of
        # programmatic dependent launch, so their sum exceeds the step); a step of embed / rerank IS one forward pass,
        # so the roofline uses the step time of the timed region
        prof_ms = sum(v for k, v in prof.items() if k.endswith("_ms"))
        pass_ms = prof_ms if w == "pipeline" else ms / args.steps
        achieved = kernel_fl / (pass_ms * 1e-3) / 1e12 if pass_ms > 0 else 0.0
        top = max((k for k in prof if k.endswith("_ms")), key=lambda k: prof[k])
        line = {
            "metric": metric_name(w), "value": units * world * args.steps / (ms * 1e-3), "unit": unit_name(w), "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(w, world),
            "e2e": {"value": e2e_units * world * e2e_steps / e2e_s, "unit": unit_name(w), "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "note": "host WordPiece tokenisation (tokenizers, all host threads) is inside the timed region"},
            "gpu_launches": int(prof["launches"]) * args.steps if w != "pipeline" else (63 + 3 + 33) * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": kernel, "kernel_ms": pass_ms, "profiled_pass_ms": prof_ms, "algorithmic_flops_per_launch": kernel_fl,
                         "peak_source": peak_src, "per_pass_kernel_class_ms": {k: round(v, 4) for k, v in prof.items() if k.endswith("_ms")},
                         "dominant_class": top, "step_flops": fl},
            "clocks": summarize_clocks(samples),
        }
        if world == 1 and with_cpu and w != "embed_varlen":  # (the CPU leg is timed on 512-token chunks: see --workload embed)
            if w == "pipeline":
                r_e, _ = cpu_encode_rate("embed", 6.0)
                r_r, cb = cpu_encode_rate("rerank", 10.0)
                per_batch = NQ * 20 / (r_e * 512) + NQ * LIMIT / r_r
                cb = dict(cb, value=NQ / per_batch, unit="queries/s",
                          sample=cb["sample"] + "; per 32-query batch = one query-embedding call + 32 rerank calls of 15 pairs")
            else:
                _, cb = cpu_encode_rate(w, 15.0)
            line["cpu_baseline"] = cb
    else:
        line = None
    if retr is not None:
        retr.close()
    else:
        model.close()
    return line


def run_ours(args, ClockSampler, summarize_clocks, Ctx):
    cx = Ctx(args)
    line = measure(cx, args.workload, args.steps, args.warmup, ClockSampler, summarize_clocks)
    if line is not None:
        print(json.dumps(line), flush=True)
    if cx.world > 1:
        cx.dist.destroy_process_group()


def measure_compact(cx, w, ClockSampler, summarize_clocks, steps=None):
    """The same measurement, boiled down for the `secondary` object of the headline line."""
    steps = steps or 10
    line = measure(cx, w, steps, 3, ClockSampler, summarize_clocks, with_cpu=cx.world == 1 and w != "embed_varlen")
    if line is None:
        return None
    out = {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "steps": steps,
           "e2e": line["e2e"]["value"], "frac_of_sustained_bf16": line["roofline"]["frac"], "achieved_tflops": line["roofline"]["achieved"],
           "per_pass_kernel_class_ms": line["roofline"]["per_pass_kernel_class_ms"], "clocks": line["clocks"],
           "workload": line["config"]["workload"], "n_gpus": line["n_gpus"]}
    if "cpu_baseline" in line:
        out["cpu_baseline"] = {k: line["cpu_baseline"][k] for k in ("value", "unit", "cores", "kind", "sample")}
    return out

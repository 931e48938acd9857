import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_data as bd
from financial_rag_system_b200.index import VectorIndex

n = int(os.environ.get("ROWS", 10_000_000))
dev = torch.device("cuda", 0)
ix = VectorIndex(n)
cent = bd.centroids_torch(dev); cdf = torch.from_numpy(bd.zipf_cdf()).to(dev)
for s in range(0, n, 1 << 17):
    m = min(1 << 17, n - s)
    x, c = bd.rows_torch(s, m, dev, cent=cent, cdf=cdf); ix.add(x, c)
ix.set_pipeline_reserve(int(os.environ.get("RES", 8)))
q, t, m = bd.queries_np("self")
qc, qm = t.astype(np.uint32), m.astype(np.uint32)

def run(steps, depth, prof=0):
    ix.set_profiling(prof)
    inflight = []
    t0 = time.perf_counter()
    tsub = tcol = 0.0
    for _ in range(steps):
        a = time.perf_counter(); inflight.append(ix.submit_host(q, qc, qm, 15)); b = time.perf_counter(); tsub += b - a
        if len(inflight) >= depth:
            ix.collect_host(inflight.pop(0)); tcol += time.perf_counter() - b
    while inflight:
        ix.collect_host(inflight.pop(0))
    dt = time.perf_counter() - t0
    p = ix.read_profile_ex() if prof else None
    ix.set_profiling(0)
    return dt / steps * 1e3, tsub / steps * 1e6, tcol / steps * 1e6, p

run(20, 3)
for depth in (1, 2, 3, 4):
    ms, us_sub, us_col, _ = run(100, depth)
    _, _, _, p3 = run(100, depth, 3)
    _, _, _, p1 = run(100, depth, 1)
    print(f"depth {depth}: {ms:.4f} ms/step  submit {us_sub:.1f} us  collect {us_col:.1f} us | bracket scan {p3['scan_ms']/max(p3['n'],1):.4f} | detailed prep {p1['prep_ms']/p1['n']:.4f} scan {p1['scan_ms']/p1['n']:.4f} merge {p1['merge_ms']/p1['n']:.4f} gap {p1['scan_gap_ms']/(p1['n']-1):.4f} span/n {p1['span_ms']/p1['n']:.4f}")
# device path for comparison
qd = torch.from_numpy(q).to(dev); qcd = torch.from_numpy(t.astype(np.int64)).to(torch.int32).to(dev); qmd = torch.from_numpy(m.astype(np.int64)).to(torch.int32).to(dev)
for _ in range(5): ix.search_async(qd, qcd, qmd, 15)
ix.sync(-1)
t0 = time.perf_counter()
for _ in range(100): ix.search_async(qd, qcd, qmd, 15)
ix.sync(-1)
print("device pipelined ms/step", (time.perf_counter() - t0) * 10)

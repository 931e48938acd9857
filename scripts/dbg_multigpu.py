import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_data as bd
from financial_rag_system_b200.multigpu import MultiGpuIndex

rows = int(os.environ.get("ROWS", 4_000_000))
dev = torch.device("cuda", 0)
cent = bd.centroids_torch(dev); cdf = torch.from_numpy(bd.zipf_cdf()).to(dev)
q, t, m = bd.queries_np("self")
qc, qm = t.astype(np.uint32), m.astype(np.uint32)
for devices in ([0], [0, 0], [0, 1] if torch.cuda.device_count() > 1 else [0, 0, 0]):
    mg = MultiGpuIndex(rows, devices=devices)
    for s in range(0, rows, 1 << 18):
        n = min(1 << 18, rows - s)
        x, c = bd.rows_torch(s, n, dev, cent=cent, cdf=cdf); mg.add_device(x, c)
    for _ in range(5): mg.search(q, qc, qm, 15)
    ts, tc = [], []
    for _ in range(50):
        a = time.perf_counter(); tk = mg.submit(q, qc, qm, 15); b = time.perf_counter(); mg.collect(tk); c_ = time.perf_counter()
        ts.append(b - a); tc.append(c_ - b)
    print(f"devices {devices}: rows {rows}: submit {np.median(ts)*1e6:.0f} us  collect {np.median(tc)*1e6:.0f} us  (ideal scan {rows/len(devices)*772/7.2e6:.0f} us per shard)")
    # pipelined depth 3
    t0 = time.perf_counter(); infl = []
    for _ in range(60):
        infl.append(mg.submit(q, qc, qm, 15))
        if len(infl) >= 3: mg.collect(infl.pop(0))
    while infl: mg.collect(infl.pop(0))
    print(f"   depth 3: {(time.perf_counter()-t0)/60*1e6:.0f} us per batch")
    mg.close()

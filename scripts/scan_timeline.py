"""Per-CTA timeline of one scan launch (frs_index_set_profiling(2)): where a launch's time goes at a given shard size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_data as bd
from financial_rag_system_b200.index import VectorIndex

dev = torch.device("cuda", 0)
for n in [int(a) for a in sys.argv[1:]] or [1_250_000, 10_000_000]:
    ix = VectorIndex(n)
    cent = bd.centroids_torch(dev); cdf = torch.from_numpy(bd.zipf_cdf()).to(dev)
    for s in range(0, n, 1 << 17):
        m = min(1 << 17, n - s)
        x, c = bd.rows_torch(s, m, dev, cent=cent, cdf=cdf); ix.add(x, c)
    for kind in ("self", "unrelated"):
        q, t, m = bd.queries_np(kind)
        qd = torch.from_numpy(q).to(dev); qc = torch.from_numpy(t.astype(np.int64)).to(torch.int32).to(dev); qm = torch.from_numpy(m.astype(np.int64)).to(torch.int32).to(dev)
        for grid in (148, 140):
            ix.set_scan_grid(grid)
            for _ in range(3): ix.search(qd, qc, qm, 15)
            torch.cuda.synchronize()
            ix.set_profiling(2)
            ix.search(qd, qc, qm, 15)
            torch.cuda.synchronize()
            tl = ix.read_timeline(grid).astype(np.int64)
            p = ix.read_profile()
            ix.set_profiling(0)
            st = ix.last_stats()
            t0 = tl[:, 0].min()
            start, first, lastmma, firsttile, lasttile, exit_ = [(tl[:, i] - t0) / 1e3 for i in range(6)]
            nslow, twait, tslow = tl[:, 12], tl[:, 13] / 1e3, tl[:, 14] / 1e3
            tiles = (n + 127) // 128
            print(f"n={n} {kind} grid={grid} tiles/CTA={tiles / grid:.1f} scan_ms={p['scan_ms']:.4f} ideal_us@7.2TB/s={n * 772 / 7.2e6:.1f} stats={st}")
            print(f"   start  us  min {start.min():.1f} max {start.max():.1f}")
            print(f"   first slab landed - start: mean {np.mean(first - start):.1f} max {np.max(first - start):.1f}")
            print(f"   first tile consumed: mean {firsttile.mean():.1f} max {firsttile.max():.1f}")
            print(f"   last tile consumed: min {lasttile.min():.1f} mean {lasttile.mean():.1f} max {lasttile.max():.1f}")
            print(f"   exit: min {exit_.min():.1f} mean {exit_.mean():.1f} max {exit_.max():.1f}")
            print(f"   rare path: invocations/CTA mean {nslow.mean():.1f} max {nslow.max()}  time/CTA us mean {tslow.mean():.1f} max {tslow.max():.1f}; epilogue wait for accumulators/CTA us mean {twait.mean():.1f}")
    ix.close()

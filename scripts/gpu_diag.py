"""GPU diagnostics (development aid): validates the tensor-core pre-filter against fp64, then the
full search against the oracle, then times the scan.  Run on the B200 box:  python scripts/gpu_diag.py"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from financial_rag_system_b200.index import VectorIndex  # noqa: E402
from oracle import search_oracle as so  # noqa: E402

ANY = np.uint32(0x80000000)  # mask: only the tombstone bit


def diag_scores(dtype, n):
    g = torch.Generator(device="cuda").manual_seed(3)
    ix = VectorIndex(n, dtype=dtype)
    x = torch.randn((n, 384), generator=g, device="cuda")
    ix.add(x)
    q = torch.randn((32, 384), generator=g, device="cuda")
    s = ix.debug_scores(q).cpu().numpy().astype(np.float64)
    rows = ix.read_rows().cpu().numpy()
    qp = ix.last_queries().cpu().numpy()
    ref = so.scores_f64(rows, qp)
    err = np.abs(s - ref)
    print(f"[scores {dtype} n={n}] max|tc-f64|={err.max():.3e} mean={err.mean():.3e} "
          f"ref absmax={np.abs(ref).max():.3f} bad rows={(err.max(axis=0) > 1e-2).sum()}", flush=True)
    if err.max() > 1e-2:
        bad = np.argwhere(err > 1e-2)
        print("  first bad (q,row):", bad[:10].tolist(), flush=True)
        print("  got", s[bad[0][0], bad[0][1]], "want", ref[bad[0][0], bad[0][1]])
    # normalisation parity
    o = so.store_rows(x.cpu().numpy(), dtype)
    print(f"  stored rows vs oracle store_rows: max diff {np.abs(o - rows).max():.3e} mismatching elems {(o != rows).sum()}", flush=True)
    ix.close()
    return err.max()


def diag_search(dtype, n, nq=32, k=15, filt=True, grid=0):
    g = torch.Generator(device="cuda").manual_seed(5)
    ix = VectorIndex(n, dtype=dtype)
    x = torch.randn((n, 384), generator=g, device="cuda")
    codes = torch.randint(0, 5, (n,), generator=g, device="cuda", dtype=torch.int32)
    ix.add(x, codes)
    if grid:
        ix.set_scan_grid(grid)
    q = x[:nq] + 0.2 * torch.randn((nq, 384), generator=g, device="cuda")
    qc = codes[:nq].clone()
    mval = 0x80FFFFFF if filt else 0x80000000
    qm = torch.full((nq,), mval - (1 << 32), dtype=torch.int64).to(torch.int32).cuda()
    ids, sc = ix.search(q, qc, qm, k)
    torch.cuda.synchronize()
    st = ix.last_stats()
    rows = ix.read_rows().cpu().numpy()
    qp = ix.last_queries().cpu().numpy()[:nq]
    oi, os_ = so.exact_topk(rows, qp, codes.cpu().numpy().astype(np.uint32), qc.cpu().numpy().astype(np.uint32),
                            qm.cpu().numpy().astype(np.uint32), k)
    ok_ids = np.array_equal(ids.cpu().numpy(), oi)
    dsc = np.abs(sc.cpu().numpy().astype(np.float64) - os_)
    dsc = dsc[np.isfinite(dsc)]
    print(f"[search {dtype} n={n} filt={filt} grid={grid or 'auto'}] ids equal={ok_ids} max score diff={dsc.max() if dsc.size else 0:.2e} stats={st}", flush=True)
    if not ok_ids:
        bad = np.argwhere(ids.cpu().numpy() != oi)
        print("  first mismatches (q,rank):", bad[:8].tolist())
        qi = bad[0][0]
        print("  got ", ids[qi].tolist(), "\n  want", oi[qi].tolist())
        print("  got s", sc[qi].tolist(), "\n  want s", os_[qi].tolist())
    ix.close()
    return ok_ids


def bench(dtype, n, iters=20, nomatch=False):
    g = torch.Generator(device="cuda").manual_seed(9)
    ix = VectorIndex(n, dtype=dtype)
    chunk = 1 << 18
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        ix.add(torch.randn((m, 384), generator=g, device="cuda"),
               torch.randint(0, 500, (m,), generator=g, device="cuda", dtype=torch.int32))
    q = torch.randn((32, 384), generator=g, device="cuda")
    qc = torch.zeros(32, dtype=torch.int32, device="cuda")
    qm = torch.full((32,), 0x80000000 - (1 << 32), dtype=torch.int64).to(torch.int32).cuda()
    if nomatch:  # a ticker no row has: zero candidates -> raw streaming rate of the scan
        qc = torch.full((32,), 0xFFFFFF, dtype=torch.int32, device="cuda")
        qm = torch.full((32,), 0x80FFFFFF - (1 << 32), dtype=torch.int64).to(torch.int32).cuda()
    for _ in range(3):
        ix.search(q, qc, qm, 15)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ix.search(q, qc, qm, 15)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_ = n * 384 * (4 if dtype == "f32" else 2) + n * 4
    print(f"[bench {dtype} n={n} nomatch={nomatch}] {ms*1e3:.1f} us/batch  {bytes_/ms/1e6:.0f} GB/s  {32/ms*1e3:.0f} QPS  stats={ix.last_stats()}", flush=True)
    ix.set_profiling(2)
    for _ in range(iters):
        ix.search(q, qc, qm, 15)
    torch.cuda.synchronize()
    pr = ix.read_profile()
    nn = max(pr["n"], 1)
    print(f"   per-kernel (events): prep {pr['prep_ms']/nn*1e3:.1f} us  scan {pr['scan_ms']/nn*1e3:.1f} us  merge {pr['merge_ms']/nn*1e3:.1f} us"
          f"  -> scan {bytes_/(pr['scan_ms']/nn)/1e6:.0f} GB/s", flush=True)
    grid = ix.last_stats()["grid"]
    tl = ix.read_timeline(grid).astype(np.int64)
    t0 = tl[:, 0].min()
    rel = (tl[:, :6] - t0) / 1e3
    names = ["start", "first_slab", "last_mma", "first_tile_done", "last_tile_done", "exit"]
    for j, nm in enumerate(names):
        print(f"   timeline {nm:>16}: min {rel[:, j].min():8.1f}  median {np.median(rel[:, j]):8.1f}  max {rel[:, j].max():8.1f} us", flush=True)
    print(f"   epilogue warp0 per CTA (median): slow-path calls={np.median(tl[:, 12]):.0f}  wait_tfull={np.median(tl[:, 13])/1e3:.1f} us"
          f"  in_slow_path={np.median(tl[:, 14])/1e3:.1f} us  in_refresh={np.median(tl[:, 15])/1e3:.1f} us", flush=True)
    ix.set_profiling(0)
    ix.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    steps = [
        lambda: diag_scores("bf16", 50000),
        lambda: diag_scores("f32", 50000),
        lambda: diag_search("bf16", 1000),
        lambda: diag_search("bf16", 100000),
        lambda: diag_search("bf16", 100000, filt=False),
        lambda: diag_search("bf16", 100000, grid=7),
        lambda: diag_search("f32", 100000),
        lambda: diag_search("f32", 100000, filt=False, grid=3),
        lambda: bench("bf16", 1_000_000),
        lambda: bench("bf16", 1_000_000, nomatch=True),
        lambda: bench("bf16", 10_000_000),
    ]
    for s in steps:
        try:
            s()
        except Exception:
            traceback.print_exc()
            if "CUDA" in traceback.format_exc() or "frs_b200 error -2" in traceback.format_exc():
                print("CUDA failure: stopping", flush=True)
                break

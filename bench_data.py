"""Synthetic SEC-style corpus of bench.py — ONE recipe for both arms.

Every value is a pure function of (seed, global row, column) through a 64-bit integer hash, computed with
the same integer / IEEE operations by numpy on the host (`rows_np`: the CPU reference arm and the oracle
checks) and by torch on the GPU (`rows_torch`: our arm builds its shards on the device, never on the host),
so both arms search bit-identical fp32 input rows, tickers and queries — whatever the number of ranks.

  row r   = centroid[cid(r)] + noise(r)          1024 centroids of norm ~1, noise of norm ~0.3
            (stored L2-normalised by the index / the oracle, as a COSINE collection does, ingest.py:89-95)
  ticker  ~ Zipf(1.1) over 500 tickers            (payload keyword of main.py:218-223)
  queries : see `queries_*` — perturbed copies of corpus rows with their tickers (a user asking about a
            filing that exists), plus the unfriendly variants bench.py --queries selects.

Gaussian-like values are Irwin-Hall sums of four 16-bit fields of the hash (exact in fp32, no
transcendental functions, hence bit-identical across devices).
"""
from __future__ import annotations

import numpy as np

DIM = 384
N_CENTROIDS = 1024
N_TICKERS = 500
SEED = 7
_G1, _G2, _G3 = 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9
_M1, _M2 = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
_IH_STD = float(np.sqrt(4 * (65536.0 ** 2 - 1) / 12.0))   # std of the sum of four uniform 16-bit integers
_IH_MEAN = 2.0 * 65535.0
NOISE_SCALE = np.float32(0.3 / np.sqrt(DIM) / _IH_STD)      # per-component std 0.3 / sqrt(384): noise norm ~0.3
CENT_SCALE = np.float32(1.0 / np.sqrt(DIM) / _IH_STD)       # centroid norm ~1
UNIT_SCALE = np.float32(1.0 / _IH_STD)                      # std 1


def zipf_cdf(n: int = N_TICKERS, a: float = 1.1) -> np.ndarray:
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** a
    return np.cumsum(p / p.sum())


# ---- numpy ---------------------------------------------------------------------------------------
def _mix_np(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
    return z ^ (z >> np.uint64(31))


def _ih_np(h: np.ndarray) -> np.ndarray:
    m = np.uint64(0xFFFF)
    s = (h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))
    return s.astype(np.float32) - np.float32(_IH_MEAN)


def _field_np(rows: np.ndarray, cols: int, seed: int, stream: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = rows.astype(np.uint64)[:, None] * np.uint64(cols) + np.arange(cols, dtype=np.uint64)[None, :]
        return _ih_np(_mix_np(idx * np.uint64(_G1) + np.uint64((seed * 0x10001 + stream * 0x9E37) & 0xFFFFFFFFFFFFFFFF)))


def centroids_np(seed: int = SEED) -> np.ndarray:
    return _field_np(np.arange(N_CENTROIDS), DIM, seed, 1) * CENT_SCALE


def rows_np(row0: int, m: int, seed: int = SEED, cent: np.ndarray | None = None):
    """(x float32 [m,384] un-normalised, ticker uint32 [m]) of global rows [row0, row0 + m)."""
    cent = centroids_np(seed) if cent is None else cent
    r = np.arange(row0, row0 + m, dtype=np.uint64)
    with np.errstate(over="ignore"):
        cid = (_mix_np(r * np.uint64(_G2) + np.uint64(seed + 2)) >> np.uint64(40)) & np.uint64(N_CENTROIDS - 1)
        u = (_mix_np(r * np.uint64(_G3) + np.uint64(seed + 3)) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    x = cent[cid.astype(np.int64)] + _field_np(r, DIM, seed, 4) * NOISE_SCALE
    t = np.minimum(np.searchsorted(zipf_cdf(), u, side="right"), N_TICKERS - 1).astype(np.uint32)
    return x.astype(np.float32), t


def unit_noise_np(n: int, seed: int, stream: int) -> np.ndarray:
    return _field_np(np.arange(n), DIM, seed, stream) * UNIT_SCALE


# ---- torch ---------------------------------------------------------------------------------------
def _s64(c: int) -> int:
    c &= 0xFFFFFFFFFFFFFFFF
    return c - (1 << 64) if c >= (1 << 63) else c


def _shr(z, k: int):
    return (z >> k) & ((1 << (64 - k)) - 1)   # logical shift of the int64 bit pattern


def _mix_t(z):
    z = (z ^ _shr(z, 30)) * _s64(_M1)
    z = (z ^ _shr(z, 27)) * _s64(_M2)
    return z ^ _shr(z, 31)


def _ih_t(h):
    import torch

    s = (h & 0xFFFF) + (_shr(h, 16) & 0xFFFF) + (_shr(h, 32) & 0xFFFF) + _shr(h, 48)
    return s.to(torch.float32) - float(_IH_MEAN)


def _field_t(rows, cols: int, seed: int, stream: int):
    import torch

    idx = rows.to(torch.int64)[:, None] * cols + torch.arange(cols, dtype=torch.int64, device=rows.device)[None, :]
    return _ih_t(_mix_t(idx * _s64(_G1) + _s64(seed * 0x10001 + stream * 0x9E37)))


def centroids_torch(device, seed: int = SEED):
    import torch

    return _field_t(torch.arange(N_CENTROIDS, device=device), DIM, seed, 1) * float(CENT_SCALE)


def rows_torch(row0: int, m: int, device, seed: int = SEED, cent=None, cdf=None):
    import torch

    cent = centroids_torch(device, seed) if cent is None else cent
    cdf = torch.from_numpy(zipf_cdf()).to(device) if cdf is None else cdf
    r = torch.arange(row0, row0 + m, dtype=torch.int64, device=device)
    cid = _shr(_mix_t(r * _s64(_G2) + (seed + 2)), 40) & (N_CENTROIDS - 1)
    u = _shr(_mix_t(r * _s64(_G3) + (seed + 3)), 11).to(torch.float64) * (1.0 / 9007199254740992.0)
    x = cent[cid] + _field_t(r, DIM, seed, 4) * float(NOISE_SCALE)
    t = torch.searchsorted(cdf, u, right=True).clamp_(max=N_TICKERS - 1).to(torch.int32)
    return x, t


# ---- queries -------------------------------------------------------------------------------------
QUERY_KINDS = ("self", "any", "unrelated", "one_ticker", "rare")
TICKER_MASK = 0x80FFFFFF   # "ticker == T and not a tombstone"
ANY_MASK = 0x80000000      # "not a tombstone"


def queries_np(kind: str = "self", nq: int = 32, seed: int = SEED):
    """(q float32 [nq,384], ticker uint32 [nq], mask uint32 [nq]).

    self        rows 0..nq-1 + 0.02 * unit noise, each with its row's ticker (every query has a ~0.999 hit)
    any         the same vectors, no ticker condition (unfiltered search: every row competes)
    unrelated   unit noise vectors (no near neighbour anywhere: best scores ~0.25), tickers ~ Zipf
    one_ticker  unrelated vectors, all on the hottest ticker (0): ~19 % of the rows pass the filter for every query
    rare        unrelated vectors on the rarest tickers (499, 498, ...: ~2e-4 of the rows each)"""
    if kind not in QUERY_KINDS:
        raise ValueError(f"queries must be one of {QUERY_KINDS}")
    x, t = rows_np(0, nq, seed)
    if kind in ("self", "any"):
        q = x + np.float32(0.02) * unit_noise_np(nq, seed, 5)
        mask = np.full(nq, TICKER_MASK if kind == "self" else ANY_MASK, np.uint32)
        return q.astype(np.float32), t, mask
    q = unit_noise_np(nq, seed, 6)
    if kind == "unrelated":
        _, t = rows_np(1_000, nq, seed)   # tickers of some other rows: Zipf-distributed
    elif kind == "one_ticker":
        t = np.zeros(nq, np.uint32)
    else:
        t = (N_TICKERS - 1 - np.arange(nq)).astype(np.uint32)
    return q.astype(np.float32), t, np.full(nq, TICKER_MASK, np.uint32)


# ---- host fast path (CPU arm only): oracle/synth_gen.c ---------------------------------------------
def rows_host(row0: int, m: int, seed: int = SEED, normalise: bool = False, out: np.ndarray | None = None):
    """rows_np through the C + OpenMP restatement under oracle/ when it has been built (bit-identical, ~100x
    faster on a many-core host); the numpy path otherwise.  Only bench.py's CPU legs and tests call this."""
    import ctypes as C
    import os

    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_build", "libfrs_synth.so")
    x = np.empty((m, DIM), np.float32) if out is None else out
    t = np.empty(m, np.uint32)
    if os.path.exists(so):
        lib = C.CDLL(so)
        cent, cdf = np.ascontiguousarray(centroids_np(seed)), np.ascontiguousarray(zipf_cdf())
        lib.frs_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_float,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.frs_synth_rows.restype = None
        lib.frs_synth_rows(row0, m, seed, DIM, cent.ctypes.data, N_CENTROIDS, float(NOISE_SCALE), cdf.ctypes.data,
                           N_TICKERS, x.ctypes.data, t.ctypes.data)
        if normalise:
            lib.frs_synth_normalize.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
            lib.frs_synth_normalize.restype = None
            lib.frs_synth_normalize(x.ctypes.data, m, DIM)
        return x, t
    cent = centroids_np(seed)
    for s in range(0, m, 1 << 16):
        e = min(m, s + (1 << 16))
        x[s:e], t[s:e] = rows_np(row0 + s, e - s, seed, cent)
    if normalise:
        n = np.sqrt(np.sum(x * x, axis=1, keepdims=True, dtype=np.float32))
        np.divide(x, n, out=x, where=n > 0)
    return x, t

"""ctypes binding of libfrs_b200.so (C ABI declared in include/frs_b200.h).

The library is the only compute path: if it is missing this module raises — there is no CPU or
PyTorch fallback (tests/ and bench.py rely on that to prove the CUDA path is the one that runs).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FRS_B200_LIB: a differently-flagged build of the same library (kernel experiments, scripts/); never a fallback
LIB_PATH = os.environ.get("FRS_B200_LIB") or os.path.join(_HERE, "csrc", "libfrs_b200.so")

FRS_OK = 0
FRS_E_TIMEOUT = -5
FRS_DTYPE_F32 = 0
FRS_DTYPE_BF16 = 1
FRS_DIM = 384
FRS_MAX_BATCH = 32
FRS_MAX_K = 32
CODE_TICKER_MASK = 0x00FFFFFF
CODE_DOCTYPE_SHIFT = 24
CODE_DOCTYPE_MASK = 0x7F000000
CODE_TOMBSTONE = 0x80000000

_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

# name -> (restype, argtypes); mirrors include/frs_b200.h one to one
PROTOTYPES = {
    "frs_version": (_int, []),
    "frs_last_error": (C.c_char_p, []),
    "frs_device_count": (_int, []),
    "frs_index_create": (_int, [_int, _int, _i64, _int, C.POINTER(_vp)]),
    "frs_index_destroy": (_int, [_vp]),
    "frs_index_size": (_i64, [_vp]),
    "frs_index_capacity": (_i64, [_vp]),
    "frs_index_dtype": (_int, [_vp]),
    "frs_index_device": (_int, [_vp]),
    "frs_index_set_base": (_int, [_vp, _i64]),
    "frs_index_set_scan_grid": (_int, [_vp, _int]),
    "frs_index_set_pipeline_reserve": (_int, [_vp, _int]),
    "frs_index_set_scan_streams": (_int, [_vp, _int]),
    "frs_index_add": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "frs_index_add_host": (_int, [_vp, _vp, _vp, _i64]),
    "frs_index_set_rows": (_int, [_vp, _i64, _vp, _vp, _i64, _vp]),
    "frs_index_set_codes": (_int, [_vp, _i64, _vp, _i64, _vp]),
    "frs_index_read_rows": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "frs_index_read_rows_host": (_int, [_vp, _i64, _i64, _vp]),
    "frs_index_rows_ptr": (_vp, [_vp]),
    "frs_index_codes_ptr": (_vp, [_vp]),
    "frs_index_set_size": (_int, [_vp, _i64]),
    "frs_index_export_raw": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "frs_index_import_raw": (_int, [_vp, _vp, _vp, _i64]),
    "frs_index_search": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp]),
    "frs_index_search_tiles": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _i64, _vp, _vp, _vp]),
    "frs_index_search_host": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _vp]),
    "frs_index_search_async": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, C.POINTER(_int)]),
    "frs_index_wait": (_int, [_vp, _int, _vp]),
    "frs_index_sync": (_int, [_vp, _int]),
    "frs_index_search_host_submit": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, C.POINTER(_int)]),
    "frs_index_search_host_collect": (_int, [_vp, _vp, _int, _vp, _vp]),
    "frs_index_search_local": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp]),
    "frs_merge_shards": (_int, [_int, _vp, _vp, _int, _int, _int, _vp, _vp, _vp]),
    "frs_merge_shards_packed": (_int, [_int, _vp, _int, _int, _int, _vp, _vp, _vp]),
    "frs_exchange_create": (_int, [_int, _int, _int, _int, _int, C.POINTER(_vp)]),
    "frs_exchange_destroy": (_int, [_vp]),
    "frs_exchange_handle": (_int, [_vp, _vp]),
    "frs_exchange_connect": (_int, [_vp, _vp]),
    "frs_exchange_connect_local": (_int, [_vp, _vp]),
    "frs_exchange_push": (_int, [_vp, _vp, _vp]),
    "frs_exchange_wait_merge": (_int, [_vp, _vp, _vp, _vp]),
    "frs_exchange_wait_merge_n": (_int, [_vp, _int, _int, _vp, _vp, _vp]),
    "frs_exchange_set_timeout_ms": (_int, [_vp, _i64]),
    "frs_exchange_status": (_int, [_vp]),
    "frs_index_search_push": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _vp]),
    "frs_sharded_create": (_int, [_int, C.POINTER(_int), _int, _i64, _int, C.POINTER(_vp)]),
    "frs_sharded_destroy": (_int, [_vp]),
    "frs_sharded_n_shards": (_int, [_vp]),
    "frs_sharded_size": (_i64, [_vp]),
    "frs_sharded_capacity": (_i64, [_vp]),
    "frs_sharded_block_rows": (_i64, [_vp]),
    "frs_sharded_shard": (_vp, [_vp, _int]),
    "frs_sharded_set_size": (_int, [_vp, _i64]),
    "frs_sharded_add_host": (_int, [_vp, _vp, _vp, _i64]),
    "frs_sharded_set_rows_host": (_int, [_vp, _i64, _vp, _vp, _i64]),
    "frs_sharded_read_rows_host": (_int, [_vp, _i64, _i64, _vp]),
    "frs_sharded_export_raw": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "frs_sharded_import_raw": (_int, [_vp, _vp, _vp, _i64]),
    "frs_sharded_search_host": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _vp]),
    "frs_sharded_search_host_submit": (_int, [_vp, _vp, _vp, _vp, _int, _int, C.POINTER(_int)]),
    "frs_sharded_search_host_collect": (_int, [_vp, _int, _vp, _vp]),
    "frs_index_last_queries": (_int, [_vp, _vp, _vp]),
    "frs_index_debug_scores": (_int, [_vp, _vp, _int, _vp, _vp]),
    "frs_index_last_stats": (_int, [_vp, C.POINTER(_i64)]),
    "frs_index_set_profiling": (_int, [_vp, _int]),
    "frs_index_read_profile": (_int, [_vp, C.POINTER(C.c_double)]),
    "frs_index_read_profile_ex": (_int, [_vp, C.POINTER(C.c_double)]),
    "frs_index_read_profile_bracket_rel": (_int, [_vp, _vp, _vp, C.POINTER(C.c_double)]),
    "frs_index_read_profile_raw": (_int, [_vp, C.POINTER(C.c_double), _int]),
    "frs_index_read_timeline": (_int, [_vp, _vp, _int]),
    "frs_encoder_create": (_int, [_int, _vp, _vp, _int, _int, _int, C.POINTER(_vp)]),
    "frs_encoder_destroy": (_int, [_vp]),
    "frs_encoder_max_tokens": (_int, [_vp]),
    "frs_encoder_embed": (_int, [_vp, _vp, _vp, _int, _int, _vp, _vp]),
    "frs_encoder_embed_host": (_int, [_vp, _vp, _vp, _int, _int, _vp]),
    "frs_encoder_score_pairs": (_int, [_vp, _vp, _vp, _vp, _int, _vp, _vp]),
    "frs_encoder_score_pairs_host": (_int, [_vp, _vp, _vp, _vp, _int, _vp]),
    "frs_encoder_last_hidden": (_int, [_vp, _vp, _int, _vp]),
    "frs_encoder_debug_read": (_int, [_vp, _int, _vp, _i64, _vp]),
    "frs_encoder_set_profiling": (_int, [_vp, _int]),
    "frs_encoder_read_profile": (_int, [_vp, C.POINTER(C.c_double)]),
}

_lib = None


class FrsError(RuntimeError):
    """Non-zero status from the C ABI (message from frs_last_error())."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"frs_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load libfrs_b200.so once.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
                "or `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the retrieval path."
            )
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int) -> None:
    if rc != FRS_OK:
        raise FrsError(rc, lib().frs_last_error().decode("utf-8", "replace"))

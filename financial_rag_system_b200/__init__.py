"""B200-native retrieval hot path for financial-rag-system (embed -> exact cosine top-15 -> rerank).

The compute path is libfrs_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/frs_b200.h); this package is the host-side mirror of the reference's call surface.
"""
from ._lib import FRS_DIM, FRS_MAX_BATCH, FRS_MAX_K, FrsError  # noqa: F401

__all__ = ["FRS_DIM", "FRS_MAX_BATCH", "FRS_MAX_K", "FrsError"]

"""The C-ABI library: loads, exports every symbol include/frs_b200.h declares, the ctypes table
mirrors the header, and without a CUDA device the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from financial_rag_system_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "frs_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/frs_b200.h but not exported"


def test_ctypes_table_mirrors_header():
    assert sorted(_lib.PROTOTYPES) == _declared()


def test_version_and_error_string():
    lib = _lib.lib()
    assert lib.frs_version() == 200
    assert isinstance(lib.frs_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_a_device():
    lib = _lib.lib()
    h = ctypes.c_void_p()
    assert lib.frs_index_create(0, 128, 10, _lib.FRS_DTYPE_BF16, ctypes.byref(h)) == -1  # FRS_E_INVALID: dim
    assert b"dim" in lib.frs_last_error()
    assert lib.frs_index_create(0, 384, 10, 7, ctypes.byref(h)) == -1
    assert lib.frs_index_create(0, 384, 0, _lib.FRS_DTYPE_BF16, ctypes.byref(h)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a CUDA device")
def test_no_cpu_fallback():
    """No GPU => creating the chunk store is an error, never a silent CPU path."""
    lib = _lib.lib()
    h = ctypes.c_void_p()
    rc = lib.frs_index_create(0, 384, 1000, _lib.FRS_DTYPE_BF16, ctypes.byref(h))
    assert rc == -2 and not h.value  # FRS_E_CUDA
    from financial_rag_system_b200.index import VectorIndex

    with pytest.raises(_lib.FrsError):
        VectorIndex(1000)


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "financial_rag_system_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f

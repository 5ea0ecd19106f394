"""cammiq_b200 -- B200 (sm_100a) implementation of CAMMiQ's query-time read-matching path.

The product is the C-ABI shared library ``libcammiq_gpu.so`` (include/cammiq_gpu.h) and the
C++ ``cammiq`` CLI built from ``cammiq_b200/csrc``.  This Python package is a thin ctypes
binding of that ABI used by tests, bench.py and the multi-GPU launcher; it contains no
compute and no fallback: if the library is missing, importing :mod:`cammiq_b200.capi`
raises.
"""
from .capi import (  # noqa: F401
    CLASS_CONFLICT, CLASS_D_INTER, CLASS_D_PAIR, CLASS_U, CLASS_UD, CLASS_UNLABELED,
    MODE_P, MODE_SC, TABLE_D, TABLE_U, CammiqError, Context, Index, MultiContext, lib, library_path, pack_isa, pack_reads,
)

__all__ = ["Index", "Context", "MultiContext", "CammiqError", "lib", "library_path", "MODE_P", "MODE_SC",
           "TABLE_U", "TABLE_D", "pack_reads", "pack_isa"]

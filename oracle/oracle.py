"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_engine.hpp).

ctypes binding of oracle/_build/liboracle.so for tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. Never imported by the gamba_b200 package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")
DRIVER = os.path.join(_HERE, "_build", "gamba_oracle")


class _Result(C.Structure):
    _fields_ = [("n_new", C.c_uint32), ("nnz_new", C.c_uint64), ("row_ptr", C.POINTER(C.c_uint64)),
                ("col_rel", C.POINTER(C.c_uint32)), ("coeff", C.POINTER(C.c_uint32)),
                ("fma_top", C.c_uint64), ("seconds", C.c_double)]


def build(force: bool = False) -> None:
    if force or not (os.path.exists(LIB) and os.path.exists(DRIVER)):
        subprocess.run(["make", "-C", _HERE, "-j4"], check=True, capture_output=True)


_lib = None


def host_threads() -> int:
    """Threads for the checker mode: all the cores this process may run on."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.gamba_oracle_reduce.argtypes = [C.c_uint32] * 4 + [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p,
                                                                                  C.POINTER(_Result)]
        _lib.gamba_oracle_reduce_mt.argtypes = [C.c_uint32] * 4 + [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                                                                     C.POINTER(_Result)]
        _lib.gamba_oracle_free.argtypes = [C.POINTER(_Result)]
    return _lib


def reduce(step, threads: int = 1):
    """Canonical RREF of D' for a gamba_b200.stepfile.Step -> (Rows, fma_top, seconds).
    threads <= 1: the single-thread timing mode; threads > 1: the checker mode (oracle_checker.cpp)."""
    from gamba_b200.stepfile import Rows
    lib = _load()
    pool = np.ascontiguousarray(step.pool, dtype=np.uint32)
    off = step.coeff_offsets()
    ptrs = (pool.ctypes.data + off[:-1] * 4).astype(np.uint64)
    row_ptr = np.ascontiguousarray(step.row_ptr, dtype=np.uint64)
    col = np.ascontiguousarray(step.col, dtype=np.uint32)
    cid = np.ascontiguousarray(step.coeff_id, dtype=np.uint32)
    clen = np.ascontiguousarray(step.coeff_len, dtype=np.uint32)
    res = _Result()
    rc = lib.gamba_oracle_reduce_mt(step.p, step.n_top, step.n_bot, step.n_cols, row_ptr.ctypes.data, col.ctypes.data,
                                    cid.ctypes.data, len(clen), ptrs.ctypes.data, clen.ctypes.data, max(1, int(threads)), C.byref(res))
    if rc:
        raise RuntimeError("oracle failed")
    n, nnz = res.n_new, res.nnz_new
    rows = Rows(n, np.ctypeslib.as_array(res.row_ptr, (n + 1,)).copy(),
                np.ctypeslib.as_array(res.col_rel, (max(nnz, 1),))[:nnz].copy(),
                np.ctypeslib.as_array(res.coeff, (max(nnz, 1),))[:nnz].copy())
    out = (rows, int(res.fma_top), float(res.seconds))
    lib.gamba_oracle_free(C.byref(res))
    return out

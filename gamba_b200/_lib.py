"""ctypes binding of include/gamba_cuda.h. Loading fails loudly: there is no fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgamba_cuda.so")

REPR_STANDARD, REPR_MONTGOMERY = 0, 1
STEP_FINAL = 1
STEP_DEVICE_INDICES = 2
STEP_DEVICE_COLIDX = 4


class Config(C.Structure):
    _fields_ = [("p", C.c_uint32), ("coeff_bytes", C.c_uint32), ("repr", C.c_uint32), ("device", C.c_uint32),
                ("seed", C.c_uint64), ("rank", C.c_uint32), ("world_size", C.c_uint32)]


class StepIn(C.Structure):
    _fields_ = [("n_top", C.c_uint32), ("n_bot", C.c_uint32), ("n_cols", C.c_uint32), ("n_coeff_arrays", C.c_uint32),
                ("nnz", C.c_uint64), ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("coeff_id", C.c_void_p),
                ("coeff_ptr", C.c_void_p), ("coeff_len", C.c_void_p), ("flags", C.c_uint32),
                ("col_map_len", C.c_uint32), ("row_src", C.c_void_p), ("col_map", C.c_void_p), ("col_idx_len", C.c_uint64)]


class StepOut(C.Structure):
    _fields_ = [("n_new", C.c_uint32), ("n_nonpiv", C.c_uint32), ("nnz_new", C.c_uint64),
                ("row_ptr", C.c_void_p), ("col_rel", C.c_void_p), ("coeff", C.c_void_p), ("piv_rel", C.c_void_p),
                ("zero_reductions", C.c_uint64), ("seconds_total", C.c_double), ("seconds_device", C.c_double),
                ("seconds_eliminate", C.c_double), ("seconds_echelon", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("steps", C.c_uint64), ("sm_count", C.c_uint32),
                ("tile_cols", C.c_uint32), ("n_tiles", C.c_uint32), ("acc_bits", C.c_uint32),
                ("pivots_applied", C.c_uint64), ("fma_applied", C.c_uint64), ("fma_top", C.c_uint64), ("mc_combinations", C.c_uint32), ("mc_attempts", C.c_uint32),
                ("seconds_stage_last", C.c_double), ("seconds_sparse_last", C.c_double),
                ("seconds_schur_last", C.c_double), ("schur_macs", C.c_uint64),
                ("mc_ech_combinations", C.c_uint32), ("solve_mode", C.c_uint32), ("dense_macs", C.c_uint64),
                ("seconds_inverse_last", C.c_double), ("panel_t", C.c_uint32), ("reserved0", C.c_uint32),
                ("seconds_operands_last", C.c_double)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
BROADCAST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)

# every symbol include/gamba_cuda.h declares
SYMBOLS = ["gamba_cuda_abi_version", "gamba_cuda_last_error", "gamba_cuda_create", "gamba_cuda_destroy",
           "gamba_cuda_reduce", "gamba_cuda_stage", "gamba_cuda_reduce_staged", "gamba_cuda_set_allgather", "gamba_cuda_set_broadcast", "gamba_cuda_host_alloc", "gamba_cuda_host_free",
           "gamba_cuda_get_counters", "gamba_cuda_set_option", "gamba_cuda_measure_int8_peak",
           "gamba_cuda_rowgen_begin", "gamba_cuda_rowgen_add_monomials", "gamba_cuda_rowgen_add_polys", "gamba_cuda_rowgen_add_rows", "gamba_cuda_rowgen_fetch_monomials",
           "gamba_cuda_rowgen_rows", "gamba_cuda_rowgen_download_rows"]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gamba_cuda_abi_version.restype = C.c_int
    lib.gamba_cuda_last_error.restype = C.c_char_p
    lib.gamba_cuda_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.gamba_cuda_destroy.argtypes = [C.c_void_p]
    lib.gamba_cuda_destroy.restype = None
    lib.gamba_cuda_reduce.argtypes = [C.c_void_p, C.POINTER(StepIn), C.POINTER(StepOut)]
    lib.gamba_cuda_stage.argtypes = [C.c_void_p, C.POINTER(StepIn)]
    lib.gamba_cuda_reduce_staged.argtypes = [C.c_void_p, C.POINTER(StepOut)]
    lib.gamba_cuda_set_allgather.argtypes = [C.c_void_p, ALLGATHER_FN, C.c_void_p]
    lib.gamba_cuda_set_broadcast.argtypes = [C.c_void_p, BROADCAST_FN, C.c_void_p]
    lib.gamba_cuda_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.gamba_cuda_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.gamba_cuda_measure_int8_peak.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    lib.gamba_cuda_rowgen_begin.argtypes = [C.c_void_p, C.c_uint32]
    lib.gamba_cuda_rowgen_add_monomials.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    lib.gamba_cuda_rowgen_add_polys.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gamba_cuda_rowgen_add_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
    lib.gamba_cuda_rowgen_fetch_monomials.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.gamba_cuda_rowgen_rows.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.gamba_cuda_rowgen_download_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    _lib = lib
    return lib

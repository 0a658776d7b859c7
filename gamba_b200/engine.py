"""Python mirror of the reference's engine interface on top of the C ABI.

`LinalgEngine(p, seed)` plays `linalg_v3<Coeff,Order>(field, seed)` (source/f4.hpp:66);
`reduce(step)` plays `initialize(matrix)` + `reduce(params)` (source/f4.hpp:91-93) and returns the
new rows that `matrix_f4::extract_new_rows` (source/matrix.hpp:799-895) would read;
`reduce_final(step)` plays `linalg<Matrix>::reduce` (source/reduce.hpp:51-52).
Everything runs in libgamba_cuda.so on the GPU; errors raise RuntimeError like the
reference's std::exception path (source/main.cpp:149-160)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .stepfile import Rows, Step


class LinalgEngine:
    def __init__(self, p: int, seed: int = 967557673, device: int = 0, coeff_bytes: int = 4,
                 repr: int = _lib.REPR_STANDARD, rank: int = 0, world_size: int = 1):
        self.lib = _lib.load()
        self.p, self.coeff_bytes, self.device = p, coeff_bytes, device
        cfg = _lib.Config(p, coeff_bytes, repr, device, seed, rank, world_size)
        h = C.c_void_p()
        if self.lib.gamba_cuda_create(C.byref(cfg), C.byref(h)):
            raise RuntimeError("gamba_cuda_create: " + self.lib.gamba_cuda_last_error().decode())
        self.h = h
        self._keep = None
        self._cb = None
        self.last = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.gamba_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # -- helpers -----------------------------------------------------------------
    def _step_in(self, s: Step, flags: int) -> _lib.StepIn:
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[self.coeff_bytes]
        pool = np.ascontiguousarray(s.pool.astype(dt, copy=False))
        off = s.coeff_offsets()
        ptrs = (pool.ctypes.data + off[:-1] * self.coeff_bytes).astype(np.uint64)
        row_ptr = np.ascontiguousarray(s.row_ptr, dtype=np.uint64)
        col = np.ascontiguousarray(s.col, dtype=np.uint32)
        cid = np.ascontiguousarray(s.coeff_id, dtype=np.uint32)
        clen = np.ascontiguousarray(s.coeff_len, dtype=np.uint32)
        self._keep = (pool, ptrs, row_ptr, col, cid, clen)
        sin = _lib.StepIn(s.n_top, s.n_bot, s.n_cols, len(clen), s.nnz, row_ptr.ctypes.data, col.ctypes.data,
                          cid.ctypes.data, ptrs.ctypes.data, clen.ctypes.data, flags)
        if s.row_src is not None:       # indirect form: rows anywhere in `col`, entries are handles
            row_src = np.ascontiguousarray(s.row_src, dtype=np.uint64)
            col_map = np.ascontiguousarray(s.col_map, dtype=np.uint32)
            self._keep += (row_src, col_map)
            sin.row_src, sin.col_map, sin.col_map_len, sin.col_idx_len = row_src.ctypes.data, col_map.ctypes.data, len(col_map), len(col)
        return sin

    def make_resident(self, s: Step, device: int | None = None):
        """Copy the index arrays of a step to the engine's device (torch is only the allocator here) and return a handle
        for reduce_resident(): the GAMBA_CUDA_STEP_DEVICE_INDICES form of the hand-over (inputs resident in HBM; the
        coefficient arrays stay host arrays, kept on the device by the engine's own `coeff_cache`)."""
        import torch
        dev = torch.device("cuda", self.device if device is None else device)
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[self.coeff_bytes]
        pool = np.ascontiguousarray(s.pool.astype(dt, copy=False))
        off = s.coeff_offsets()
        ptrs = (pool.ctypes.data + off[:-1] * self.coeff_bytes).astype(np.uint64)
        clen = np.ascontiguousarray(s.coeff_len, dtype=np.uint32)

        def up(a, dtype):     # torch has no uint64/uint32 arithmetic, but it can hold the bytes
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype).view(np.uint8)).to(dev)
        row_ptr, col, cid = up(s.row_ptr, np.uint64), up(s.col, np.uint32), up(s.coeff_id, np.uint32)
        keep = [pool, ptrs, clen, row_ptr, col, cid]
        sin = _lib.StepIn(s.n_top, s.n_bot, s.n_cols, len(clen), s.nnz, row_ptr.data_ptr(), col.data_ptr(), cid.data_ptr(),
                          ptrs.ctypes.data, clen.ctypes.data, _lib.STEP_DEVICE_INDICES)
        if s.row_src is not None:
            row_src, col_map = up(s.row_src, np.uint64), up(s.col_map, np.uint32)
            keep += [row_src, col_map]
            sin.row_src, sin.col_map, sin.col_map_len, sin.col_idx_len = row_src.data_ptr(), col_map.data_ptr(), len(s.col_map), len(s.col)
        return (sin, keep)

    def reduce_resident(self, resident, copy: bool = True) -> Rows:
        out = _lib.StepOut()
        self._check(self.lib.gamba_cuda_reduce(self.h, C.byref(resident[0]), C.byref(out)), "gamba_cuda_reduce")
        return self._rows(out, copy)

    def _rows(self, out: _lib.StepOut, copy: bool = True) -> Rows:
        """copy=False: the arrays are views of the engine-owned result buffers (valid until the next call on the
        engine), which is what the C ABI hands out; the default copies them."""
        n, nnz = out.n_new, out.nnz_new
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[self.coeff_bytes]

        def arr(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            v = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (count,))
            return v.copy() if copy else v
        self.last = {k: getattr(out, k) for k, _ in _lib.StepOut._fields_ if not k.endswith(("ptr", "rel", "coeff"))}
        val = arr(out.coeff, nnz, dt)
        return Rows(n, arr(out.row_ptr, n + 1, np.uint64), arr(out.col_rel, nnz, np.uint32),
                    val if val.dtype == np.uint32 else val.astype(np.uint32))

    def _check(self, rc: int, what: str):
        if rc:
            raise RuntimeError(f"{what}: " + self.lib.gamba_cuda_last_error().decode())

    # -- the reference-facing calls ------------------------------------------------
    def reduce(self, s: Step | None, final: bool = False, copy: bool = True) -> Rows:
        """Host buffers in, host buffers out (H2D + kernels + D2H). `s` may be None on ranks > 0 of a
        broadcasting group (set_broadcast): the step then arrives from rank 0 over NCCL."""
        out = _lib.StepOut()
        sin = C.byref(self._step_in(s, _lib.STEP_FINAL if final else 0)) if s is not None else None
        self._check(self.lib.gamba_cuda_reduce(self.h, sin, C.byref(out)), "gamba_cuda_reduce")
        return self._rows(out, copy)

    def reduce_final(self, s: Step) -> Rows:
        return self.reduce(s, final=True)

    def stage(self, s: Step | None, final: bool = False) -> None:
        sin = C.byref(self._step_in(s, _lib.STEP_FINAL if final else 0)) if s is not None else None
        self._check(self.lib.gamba_cuda_stage(self.h, sin), "gamba_cuda_stage")

    def reduce_staged(self, copy: bool = True) -> Rows:
        out = _lib.StepOut()
        self._check(self.lib.gamba_cuda_reduce_staged(self.h, C.byref(out)), "gamba_cuda_reduce_staged")
        return self._rows(out, copy)

    def counters(self) -> dict:
        c = _lib.Counters()
        self._check(self.lib.gamba_cuda_get_counters(self.h, C.byref(c)), "gamba_cuda_get_counters")
        return {k: getattr(c, k) for k, _ in _lib.Counters._fields_}

    def set_option(self, key: str, value: int) -> None:
        self._check(self.lib.gamba_cuda_set_option(self.h, key.encode(), int(value)), "gamba_cuda_set_option")

    def set_broadcast(self, fn) -> None:
        """fn(buf_ptr, nbytes, root, stream_ptr) -> None; see parallel.make_broadcast."""
        def tramp(_user, buf, nbytes, root, stream):
            try:
                fn(buf, nbytes, root, stream)
                return 0
            except Exception:  # noqa: BLE001 - reported through the C status code
                import traceback
                traceback.print_exc()
                return 1
        self._bcb = _lib.BROADCAST_FN(tramp)
        self._check(self.lib.gamba_cuda_set_broadcast(self.h, self._bcb, None), "gamba_cuda_set_broadcast")

    def set_allgather(self, fn) -> None:
        """fn(send_ptr, recv_ptr, bytes_per_rank, stream_ptr) -> None; see parallel.py."""
        def tramp(_user, send, recv, nbytes, stream):
            try:
                fn(send, recv, nbytes, stream)
                return 0
            except Exception:  # noqa: BLE001 - reported through the C status code
                import traceback
                traceback.print_exc()
                return 1
        self._cb = _lib.ALLGATHER_FN(tramp)
        self._check(self.lib.gamba_cuda_set_allgather(self.h, self._cb, None), "gamba_cuda_set_allgather")

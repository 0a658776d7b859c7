"""gamba_b200 — B200-native F4 Macaulay-matrix reduction over F_p (hot path of GamBa).

The product is `libgamba_cuda.so` (hand-written sm_100a kernels behind the C ABI of
include/gamba_cuda.h) plus the `bin/gamba` host driver; this package is the thin Python
mirror of the reference's engine interface used by the tests and the benchmark."""
from .stepfile import Rows, Step, load_rows, load_step, save_step, synthetic_step  # noqa: F401

__all__ = ["Rows", "Step", "load_rows", "load_step", "save_step", "synthetic_step", "LinalgEngine"]


def __getattr__(name):
    if name == "LinalgEngine":
        from .engine import LinalgEngine
        return LinalgEngine
    raise AttributeError(name)

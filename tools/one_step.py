#!/usr/bin/env python
"""Stage one step matrix of a system and reduce it N times (for ncu launch lists). usage: one_step.py SYSTEM MIN_NNZ INDEX [key=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gamba_b200 import LinalgEngine, load_step
system, min_nnz, idx = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
files = bench.generate_workload(system, 2000, min_nnz, 0, 1)
s = load_step(files[idx])
eng = LinalgEngine(s.p)
for kv in sys.argv[4:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
for _ in range(3):      # the engine picks its kernels for a step from what the previous steps measured
    eng.stage(s)
    eng.reduce_staged()
c = eng.counters()
print(os.path.basename(files[idx]), s.n_top, s.n_bot, s.n_cols, "T", c["tile_cols"], "elim ms", eng.last["seconds_eliminate"] * 1e3, "sparse", c["seconds_sparse_last"] * 1e3, "inverse", c["seconds_inverse_last"] * 1e3,
      "mode", c["solve_mode"], "panel_t", c["panel_t"], "echelon ms", eng.last["seconds_echelon"] * 1e3, "launches", c["kernel_launches"])

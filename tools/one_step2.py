#!/usr/bin/env python
"""Reduce one step matrix of a bench workload N times (for ncu). usage: one_step2.py SYSTEM MAX_SPAIRS EVERY INDEX RANK_HINT [key=value ...]
INDEX -1 = the step with the most non-zeros; RANK_HINT < 0 = none (the engine predicts from its previous reduction)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gamba_b200 import LinalgEngine, load_step
system, sp, every, idx, hint = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
files = bench.generate_workload(system, sp, 0, 0, 1, every)
if idx < 0:
    idx = max(range(len(files)), key=lambda i: os.path.getsize(files[i]))
s = load_step(files[idx])
eng = LinalgEngine(s.p)
for kv in sys.argv[6:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
for _ in range(int(os.environ.get("ONE_STEP_REPEATS", "3"))):      # the engine picks its kernels for a step from what the previous steps measured
    if hint >= 0:
        eng.set_option("rank_hint", hint)
    r = eng.reduce(s)
    c = eng.counters()
    print(os.path.basename(files[idx]), s.n_top, s.n_bot, s.n_cols, "rank", r.n_new, "K", c["mc_combinations"], "T", c["tile_cols"], "elim ms", eng.last["seconds_eliminate"] * 1e3,
          "operands", c["seconds_operands_last"] * 1e3, "solve", c["seconds_sparse_last"] * 1e3, "mode", c["solve_mode"], "echelon ms", eng.last["seconds_echelon"] * 1e3,
          "launches", c["kernel_launches"], flush=True)

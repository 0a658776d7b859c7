#!/usr/bin/env python
"""One engine through the bench workload of a system for several passes (as bench.py's value leg does): solve mode, combinations and
times per step and pass.  usage: mc_seq.py SYSTEM MAX_SPAIRS EVERY PASSES [key=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gamba_b200 import LinalgEngine, load_step  # noqa: E402

system, sp, every, passes = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
steps = [load_step(f) for f in bench.generate_workload(system, sp, 0, 0, 1, every)]
eng = LinalgEngine(steps[0].p)
eng.set_option("coeff_cache", 1)
for kv in sys.argv[5:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
for ps in range(passes):
    for i, s in enumerate(steps):
        r = eng.reduce(s)
        c, L = eng.counters(), eng.last
        print(f"pass {ps} step {i} {s.n_top}+{s.n_bot} rank {r.n_new} K {c['mc_combinations']} draws {c['mc_attempts']} mode {c['solve_mode']} T={c['tile_cols']} "
              f"elim {L['seconds_eliminate'] * 1e3:.1f} ms (operands {c['seconds_operands_last'] * 1e3:.1f} solve {c['seconds_sparse_last'] * 1e3:.1f}) ech {L['seconds_echelon'] * 1e3:.1f}", flush=True)

#!/usr/bin/env python
"""Per-step comparison of the elimination with and without Monte-Carlo combinations of the rows to reduce, on the bench
workload of a system (matrices produced by the product driver, as bench.py does), one engine per setting driven through
the steps in F4 order.  usage: mc_stats.py SYSTEM MAX_SPAIRS EVERY [key=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gamba_b200 import LinalgEngine, load_step  # noqa: E402


def main():
    system, sp, every = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    opts = [kv.split("=") for kv in sys.argv[4:]]
    files = bench.generate_workload(system, sp, 0, 0, 1, every)
    settings = [("exact-nobo", {"mc": 0, "bank_order": 0}), ("exact", {"mc": 0}), ("mc", {"mc": 1, "mc_min_top": 1024, "mc_min_rows": 256})]
    engs = {}
    for f in files:
        s = load_step(f)
        if s.nnz < int(os.environ.get("MC_STATS_MIN_NNZ", "1000000")):
            continue
        rank = None
        for name, o in settings:
            if name not in engs:
                engs[name] = LinalgEngine(s.p)
                engs[name].set_option("coeff_cache", 0)
                for k, v in list(o.items()) + [(k, int(v)) for k, v in opts]:
                    engs[name].set_option(k, v)
            e = engs[name]
            if False:
                e.set_option("rank_hint", rank)
            e.reduce(s)
            if False:
                e.set_option("rank_hint", rank)
            r = e.reduce(s)
            rank = r.n_new
            c, L = e.counters(), e.last
            print(f"{os.path.basename(f)} {name:8s} {s.n_top}+{s.n_bot} x {s.n_cols} nnz {s.nnz} rank {r.n_new} K {c['mc_combinations']} draws {c['mc_attempts']} mode {c['solve_mode']} "
                  f"T={c['tile_cols']} elim {L['seconds_eliminate'] * 1e3:.1f} ms (operands {c['seconds_operands_last'] * 1e3:.1f} solve {c['seconds_sparse_last'] * 1e3:.1f} "
                  f"schur {c['seconds_schur_last'] * 1e3:.1f}) ech {L['seconds_echelon'] * 1e3:.1f} ms density {c['pivots_applied'] / max(1, s.n_bot * s.n_top):.3f}", flush=True)


if __name__ == "__main__":
    main()

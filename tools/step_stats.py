#!/usr/bin/env python
"""Per-step device statistics of the product engine on the step matrices of a named system
(matrices produced by the product driver, as bench.py does): sizes, tiles, kernel times, pivot hits.
usage: step_stats.py SYSTEM MIN_NNZ [key=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gamba_b200 import LinalgEngine, load_step  # noqa: E402


def main():
    system, min_nnz = sys.argv[1], int(sys.argv[2])
    opts = [kv.split("=") for kv in sys.argv[3:]]
    files = bench.generate_workload(system, 2000, min_nnz, 0, 1)
    if os.environ.get("STEP_STATS_MAX"):
        files = files[: int(os.environ["STEP_STATS_MAX"])]
    eng = None
    for f in files:
        s = load_step(f)
        if eng is None:
            eng = LinalgEngine(s.p)
            for k, v in opts:
                eng.set_option(k, int(v))
        eng.stage(s)
        for _ in range(3):
            eng.reduce_staged(); eng.stage(s)      # kernels are chosen from the previous steps' measurements
        best = None
        for _ in range(2):
            r = eng.reduce_staged()
            L = dict(eng.last)
            if best is None or L["seconds_eliminate"] < best["seconds_eliminate"]:
                best = L
        c = eng.counters()
        hits = c["pivots_applied"] / max(1, s.n_bot)
        print(f"{os.path.basename(f)} {s.n_top}+{s.n_bot} x {s.n_cols} nonpiv {s.n_nonpiv} nnz {s.nnz} new {r.n_new} T={c['tile_cols']} nt={c['n_tiles']} "
              f"elim {best['seconds_eliminate'] * 1e3:.2f} ms (sparse {c['seconds_sparse_last'] * 1e3:.2f} schur {c['seconds_schur_last'] * 1e3:.2f}) ech {best['seconds_echelon'] * 1e3:.2f} ms hits/row {hits:.0f} "
              f"fma_applied {c['fma_applied']:.3e} fma_top {c['fma_top']:.3e} Gfma/s {c['fma_top'] / best['seconds_eliminate'] / 1e9:.1f}", flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Run the CUDA engine on dumped step matrices and compare with the oracle's rows (bit-exact).

usage: check_steps.py [--gen NAME ...] [--dir DIR] [--min-nnz N] [--max-steps K] [--opt key=value ...]
  --gen cyclic7-15   generate the system, run the oracle driver with --dump-steps into a temp dir"""
import argparse
import glob
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gamba_b200 import LinalgEngine, load_rows, load_step, systems  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gen", nargs="*", default=[])
    ap.add_argument("--dir", default=None)
    ap.add_argument("--min-nnz", type=int, default=0)
    ap.add_argument("--max-steps", type=int, default=10 ** 9)
    ap.add_argument("--opt", nargs="*", default=[])
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--stop-after", type=int, default=0)
    a = ap.parse_args()
    dirs = [a.dir] if a.dir else []
    oracle.build()
    for name in a.gen:
        d = tempfile.mkdtemp(prefix=f"steps_{name}_")
        inp = systems.write_system(name, os.path.join(d, name + ".txt"))
        t = time.time()
        subprocess.run([oracle.DRIVER, "-i", inp, "-o", os.path.join(d, "out.txt"), "-v", "0", "--dump-steps", d,
                        "--dump-min-nnz", str(a.min_nnz)] + (["--stop-after", str(a.stop_after)] if a.stop_after else []), check=True)
        print(f"# {name}: oracle driver {time.time() - t:.1f}s -> {d}", flush=True)
        dirs.append(d)
    bad = 0
    engines = {}
    for d in dirs:
        files = sorted(glob.glob(os.path.join(d, "step_*.bin")))[: a.max_steps]
        for f in files:
            s = load_step(f)
            if s.nnz < a.min_nnz:
                continue
            want = load_rows(f.replace("step_", "rows_"))
            eng = engines.get(s.p)
            if eng is None:
                eng = engines[s.p] = LinalgEngine(s.p)
                for kv in a.opt:
                    k, v = kv.split("=")
                    eng.set_option(k, int(v))
            for _ in range(a.repeat):
                got = eng.reduce(s, final=f.endswith("step_final.bin"))
            ok = got.same_as(want)
            c = eng.counters()
            L = eng.last
            print(f"{os.path.basename(d)[:24]:24s} {os.path.basename(f):16s} {s.n_top:7d}+{s.n_bot:<5d} x {s.n_cols:<7d} nnz {s.nnz:9d} "
                  f"new {got.n_new:5d}/{want.n_new:<5d} {'OK ' if ok else 'DIFF'} T={c['tile_cols']} nt={c['n_tiles']} "
                  f"elim {L['seconds_eliminate'] * 1e3:9.3f} ms ech {L['seconds_echelon'] * 1e3:8.3f} ms dev {L['seconds_device'] * 1e3:9.3f} "
                  f"tot {L['seconds_total'] * 1e3:9.3f} ms fma {c['fma_applied']:.3e}", flush=True)
            bad += not ok
    print("steps with differences:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

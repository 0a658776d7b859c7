import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gamba_b200 import LinalgEngine
from gamba_b200.stepfile import synthetic_step
from oracle import oracle
for p in (2147483647, 65521, 32003):
    for tile in (0, 1024, 256, 128):
        for shape in ((700, 900, 0.04, 0.9), (400, 1500, 0.05, 0.5)):
            s = synthetic_step(shape[0], shape[1], shape[2], p, seed=7, top_frac=shape[3], share=10)
            want, _, _ = oracle.reduce(s)
            for mma in (1, 0):
                e = LinalgEngine(p)
                if tile: e.set_option("tile_cols", tile)
                e.set_option("echelon_mma", mma)
                got = e.reduce(s)
                c = e.counters()
                print(f"p={p} tile={tile} T={c['tile_cols']} nt={c['n_tiles']} shape={shape} mma={mma} new {got.n_new}/{want.n_new} {'OK' if got.same_as(want) else 'DIFF'}", flush=True)
                e.close()

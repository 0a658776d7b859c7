#!/usr/bin/env python
"""How full is A at the granularity a block-sparse tensor-core solve would see? For a dumped step: the fraction of 128 x 128
(and 128 x 64 ... ) blocks of the pivot part A (rows < n_top, columns < n_top) that hold at least one entry, and the entries
per occupied block. usage: block_occupancy.py STEP.bin [STEP.bin ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gamba_b200 import load_step
for f in sys.argv[1:]:
    s = load_step(f)
    nt = s.n_top
    rp = s.row_ptr.astype(np.int64)
    rows = np.repeat(np.arange(nt, dtype=np.int64), np.diff(rp[: nt + 1]))
    cols = s.col[: rp[nt]].astype(np.int64)
    inA = cols < nt
    r, c = rows[inA], cols[inA]
    nnzA = len(r)
    out = [f"{os.path.basename(f)} n_top={nt} n_bot={s.n_bot} nnz={s.nnz} nnz(A)={nnzA} ({nnzA / (nt * nt / 2):.5f} of the triangle)"]
    for bk, bc in ((128, 128), (128, 64), (256, 128), (512, 128)):
        key = (r // bk) * ((nt + bc - 1) // bc) + (c // bc)
        occ = len(np.unique(key))
        tot = ((nt + bk - 1) // bk) * ((nt + bc - 1) // bc) / 2
        out.append(f"  blocks {bk}(k) x {bc}(col): {occ} occupied of {tot:.0f} in the triangle = {occ / tot:.4f}; {nnzA / occ:.1f} entries per occupied block")
    # distribution of the distance from the diagonal
    d = (c - r)
    qs = np.quantile(d, [0.1, 0.5, 0.9, 0.99])
    out.append(f"  column - row distance of A's entries: 10/50/90/99 % quantiles = {qs.astype(np.int64).tolist()}")
    print("\n".join(out), flush=True)

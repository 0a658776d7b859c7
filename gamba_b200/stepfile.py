"""Step-matrix and result files written by the host driver (`gamba --dump-steps`,
format in gamba_b200/host/engine.hpp). One file = one F4 step in A|B / C|D block form."""
from __future__ import annotations

import dataclasses

import numpy as np

STEP_MAGIC = 0x3150545342474A
ROWS_MAGIC = 0x3157524E42474A


@dataclasses.dataclass
class Step:
    p: int
    n_top: int
    n_bot: int
    n_cols: int
    row_ptr: np.ndarray     # uint64 [n_rows + 1]
    col: np.ndarray         # uint32 [nnz], block-form columns
    coeff_id: np.ndarray    # uint32 [n_rows]
    coeff_len: np.ndarray   # uint32 [n_arrays]
    pool: np.ndarray        # uint32, the coefficient arrays back to back
    # indirect form of the ABI (include/gamba_cuda.h): `col` then holds caller handles, stored anywhere
    row_src: np.ndarray | None = None   # uint64 [n_rows]: where row r's entries start in `col`
    col_map: np.ndarray | None = None   # uint32: handle -> block-form column

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])

    @property
    def n_rows(self) -> int:
        return self.n_top + self.n_bot

    @property
    def n_nonpiv(self) -> int:
        return self.n_cols - self.n_top

    def coeff_offsets(self) -> np.ndarray:
        off = np.zeros(len(self.coeff_len) + 1, dtype=np.uint64)
        np.cumsum(self.coeff_len, out=off[1:])
        return off


@dataclasses.dataclass
class Rows:
    n_new: int
    row_ptr: np.ndarray     # uint64 [n_new + 1]
    col: np.ndarray         # uint32, relative non-pivot columns
    val: np.ndarray         # uint32, residues

    def same_as(self, other: "Rows") -> bool:
        return (self.n_new == other.n_new and np.array_equal(self.row_ptr, other.row_ptr)
                and np.array_equal(self.col, other.col) and np.array_equal(self.val, other.val))


def to_indirect(s: Step, seed: int = 1) -> Step:
    """The same step in the ABI's indirect form: rows stored in a random order with gaps between them,
    entries as handles under a random relabelling of the columns."""
    rng = np.random.default_rng(seed)
    n_rows = s.n_rows
    rp = s.row_ptr.astype(np.int64)
    lens = np.diff(rp)
    order = rng.permutation(n_rows)
    gaps = rng.integers(0, 3, size=n_rows)
    starts = np.zeros(n_rows, dtype=np.int64)
    pos = 0
    for r in order:
        pos += int(gaps[r]); starts[r] = pos; pos += int(lens[r])
    handle_of_col = rng.permutation(s.n_cols).astype(np.uint32)     # column c is called handle_of_col[c]
    col_map = np.empty(s.n_cols, dtype=np.uint32); col_map[handle_of_col] = np.arange(s.n_cols, dtype=np.uint32)
    store = rng.integers(0, s.n_cols, size=pos + 2, dtype=np.int64).astype(np.uint32)   # garbage in the gaps
    for r in range(n_rows):
        store[starts[r]: starts[r] + lens[r]] = handle_of_col[s.col[rp[r]: rp[r + 1]]]
    return Step(s.p, s.n_top, s.n_bot, s.n_cols, s.row_ptr, store, s.coeff_id, s.coeff_len, s.pool,
                row_src=starts.astype(np.uint64), col_map=col_map)


def load_step(path: str) -> Step:
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:64].view(np.uint64)
    if int(hdr[0]) != STEP_MAGIC:
        raise ValueError(f"{path}: not a step file")
    p, n_top, n_bot, n_cols, nnz, n_arr, pool_len = (int(x) for x in hdr[1:8])
    n_rows = n_top + n_bot
    o = 64
    row_ptr = raw[o:o + 8 * (n_rows + 1)].view(np.uint64).copy(); o += 8 * (n_rows + 1)
    col = raw[o:o + 4 * nnz].view(np.uint32).copy(); o += 4 * nnz
    coeff_id = raw[o:o + 4 * n_rows].view(np.uint32).copy(); o += 4 * n_rows
    coeff_len = raw[o:o + 4 * n_arr].view(np.uint32).copy(); o += 4 * n_arr
    pool = raw[o:o + 4 * pool_len].view(np.uint32).copy(); o += 4 * pool_len
    assert o == raw.size, (o, raw.size)
    return Step(p, n_top, n_bot, n_cols, row_ptr, col, coeff_id, coeff_len, pool)


def save_step(path: str, s: Step) -> None:
    hdr = np.array([STEP_MAGIC, s.p, s.n_top, s.n_bot, s.n_cols, s.nnz, len(s.coeff_len), len(s.pool)], dtype=np.uint64)
    with open(path, "wb") as f:
        for a in (hdr, s.row_ptr.astype(np.uint64), s.col.astype(np.uint32), s.coeff_id.astype(np.uint32),
                  s.coeff_len.astype(np.uint32), s.pool.astype(np.uint32)):
            f.write(a.tobytes())


def load_rows(path: str) -> Rows:
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:24].view(np.uint64)
    if int(hdr[0]) != ROWS_MAGIC:
        raise ValueError(f"{path}: not a rows file")
    n_new, nnz = int(hdr[1]), int(hdr[2])
    o = 24
    row_ptr = raw[o:o + 8 * (n_new + 1)].view(np.uint64).copy(); o += 8 * (n_new + 1)
    col = raw[o:o + 4 * nnz].view(np.uint32).copy(); o += 4 * nnz
    val = raw[o:o + 4 * nnz].view(np.uint32).copy(); o += 4 * nnz
    return Rows(n_new, row_ptr, col, val)


def thin_step(step: Step, every: int) -> Step:
    """All pivot rows, every `every`-th row to reduce: a smaller valid step on the same A|B (bounded CPU samples of
    bench.py's baseline legs; the checks on the uncapped kat15 / eco15 steps, whose 8-20 k rows cost the oracle minutes)."""
    if every <= 1:
        return step
    nt = step.n_top
    keep = np.arange(0, step.n_bot, every) + nt
    rp = step.row_ptr.astype(np.int64)
    parts, lens = [step.col[: rp[nt]]], []
    for r in keep:
        parts.append(step.col[rp[r]: rp[r + 1]]); lens.append(rp[r + 1] - rp[r])
    row_ptr = np.concatenate([rp[: nt + 1], rp[nt] + np.cumsum(lens)]).astype(np.uint64)
    return Step(step.p, nt, len(keep), step.n_cols, row_ptr, np.concatenate(parts),
                np.concatenate([step.coeff_id[:nt], step.coeff_id[keep]]), step.coeff_len, step.pool)


def synthetic_step(n_rows: int, n_cols: int, density: float, p: int, seed: int = 967557673,
                   top_frac: float = 0.92, share: int = 25) -> Step:
    """F4-shaped random step (BASELINE.json configs[4]): n_top pivot rows with
    distinct sorted pivot columns, unit lead, remaining entries uniform over the columns right of
    the lead; n_bot rows whose lead is a pivot column, sorted by descending lead; coefficient
    arrays drawn from a pool of n_rows/share arrays."""
    rng = np.random.default_rng(seed)
    n_top = max(1, int(top_frac * n_rows))
    n_bot = max(1, n_rows - n_top)
    n_top = min(n_top, n_cols - 1)
    row_len = max(2, int(round(density * n_cols)))
    piv_orig = np.sort(rng.choice(n_cols - 1, size=n_top, replace=False))  # original pivot columns
    is_piv = np.zeros(n_cols, dtype=bool); is_piv[piv_orig] = True
    newcol = np.empty(n_cols, dtype=np.int64)
    newcol[is_piv] = np.arange(n_top)
    newcol[~is_piv] = n_top + np.arange(n_cols - n_top)
    n_arr = max(1, n_rows // share)
    pool_rows = rng.integers(1, p, size=(n_arr, row_len), dtype=np.int64).astype(np.uint32)
    pool_rows[:, 0] = 1
    leads = np.concatenate([piv_orig, np.sort(rng.choice(piv_orig, size=n_bot))[::-1]])
    cols = np.empty((n_rows, row_len), dtype=np.uint32)
    for r in range(n_rows):
        lead = int(leads[r])
        avail = n_cols - lead - 1
        k = min(row_len - 1, avail)
        rest = lead + 1 + np.sort(rng.choice(avail, size=k, replace=False)) if k > 0 else np.empty(0, dtype=np.int64)
        row = np.concatenate([[lead], rest])
        if len(row) < row_len:  # pad by repeating is not allowed: shorten via sentinel
            row = np.concatenate([row, np.full(row_len - len(row), -1)])
        cols[r] = np.where(row >= 0, newcol[np.maximum(row, 0)], 0xFFFFFFFF).astype(np.uint32)
    lens = (cols != 0xFFFFFFFF).sum(axis=1)
    row_ptr = np.zeros(n_rows + 1, dtype=np.uint64); np.cumsum(lens, out=row_ptr[1:])
    col = cols[cols != 0xFFFFFFFF].astype(np.uint32)
    coeff_id = rng.integers(0, n_arr, size=n_rows).astype(np.uint32)
    return Step(p, n_top, n_bot, n_cols, row_ptr, col, coeff_id,
                np.full(n_arr, row_len, dtype=np.uint32), pool_rows.reshape(-1).copy())


def synthetic_step_large(n_rows: int, n_cols: int, density: float, p: int, seed: int = 967557673,
                         top_frac: float = 0.92, share: int = 25, chunk_rows: int = 8192) -> Step:
    """Vectorised generator of the same F4-shaped family for the BASELINE.json configs[4] sweep
    (50k x 80k ... 500k x 800k at ~0.5 % density): per row `density * n_cols` columns drawn
    uniformly right of the lead (duplicates dropped, so rows near the right edge get shorter)."""
    rng = np.random.default_rng(seed)
    n_top = min(max(1, int(top_frac * n_rows)), n_cols - 1)
    n_bot = max(1, n_rows - n_top)
    n_rows = n_top + n_bot
    row_len = max(2, int(round(density * n_cols)))
    piv_orig = np.sort(rng.choice(n_cols - 1, size=n_top, replace=False))
    is_piv = np.zeros(n_cols, dtype=bool); is_piv[piv_orig] = True
    newcol = np.empty(n_cols, dtype=np.uint32)
    newcol[is_piv] = np.arange(n_top, dtype=np.uint32)
    newcol[~is_piv] = n_top + np.arange(n_cols - n_top, dtype=np.uint32)
    leads = np.concatenate([piv_orig, np.sort(rng.choice(piv_orig, size=n_bot))[::-1]]).astype(np.int64)
    parts, lens = [], np.empty(n_rows, dtype=np.int64)
    for r0 in range(0, n_rows, chunk_rows):
        ld = leads[r0:r0 + chunk_rows]
        avail = (n_cols - 1 - ld)[:, None]
        c = ld[:, None] + 1 + (rng.random((len(ld), row_len - 1)) * avail).astype(np.int64)
        c.sort(axis=1)
        c = np.concatenate([ld[:, None], c], axis=1)
        keep = np.ones(c.shape, dtype=bool)
        keep[:, 1:] = (c[:, 1:] != c[:, :-1]) & (avail > 0)
        lens[r0:r0 + chunk_rows] = keep.sum(axis=1)
        parts.append(newcol[c[keep]])
    row_ptr = np.zeros(n_rows + 1, dtype=np.uint64); np.cumsum(lens, out=row_ptr[1:])
    n_arr = max(1, n_rows // share)
    pool_rows = rng.integers(1, p, size=(n_arr, row_len), dtype=np.int64).astype(np.uint32)
    pool_rows[:, 0] = 1
    coeff_id = rng.integers(0, n_arr, size=n_rows).astype(np.uint32)
    return Step(p, n_top, n_bot, n_cols, row_ptr, np.concatenate(parts), coeff_id,
                np.full(n_arr, row_len, dtype=np.uint32), pool_rows.reshape(-1).copy())

"""Properties of the CPU oracle itself on small steps (pure-numpy cross-check, idempotence, rank)."""
import numpy as np
import pytest

from gamba_b200.stepfile import Rows, Step, synthetic_step


def dense_rref_numpy(step: Step) -> np.ndarray:
    """Independent restatement: dense Gauss-Jordan over F_p of the whole [A|B; C|D] matrix with Python ints."""
    p, nt, nc = step.p, step.n_top, step.n_cols
    off = step.coeff_offsets()
    rows = []
    for r in range(step.n_rows):
        v = [0] * nc
        a, b = int(step.row_ptr[r]), int(step.row_ptr[r + 1])
        cf = step.pool[int(off[step.coeff_id[r]]):]
        for k in range(a, b):
            v[int(step.col[k])] = int(cf[k - a])
        rows.append(v)
    piv = {}
    for r in range(nt):  # pivot rows are already in echelon shape (lead at column r)
        piv[r] = rows[r]
    out = []
    for v in rows[nt:]:
        v = v[:]
        for c in range(nc):
            if v[c] and c in piv:
                f = v[c] * pow(piv[c][c], -1, p) % p
                pr = piv[c]
                for j in range(c, nc):
                    v[j] = (v[j] - f * pr[j]) % p
            elif v[c] and c >= nt:
                inv = pow(v[c], -1, p)
                v = [x * inv % p for x in v]
                # clear this column from the new pivots found so far
                for q in out:
                    if q[c]:
                        f = q[c]
                        for j in range(c, nc):
                            q[j] = (q[j] - f * v[j]) % p
                piv[c] = v
                out.append(v)
                break
    # full inter-reduction among the new rows
    out.sort(key=lambda q: next(j for j in range(nt, nc) if q[j]))
    for i in range(len(out) - 1, -1, -1):
        ci = next(j for j in range(nt, nc) if out[i][j])
        for q in out[:i]:
            if q[ci]:
                f = q[ci]
                for j in range(ci, nc):
                    q[j] = (q[j] - f * out[i][j]) % p
    return np.array([q[nt:] for q in out], dtype=np.int64).reshape(len(out), nc - nt)


def rows_to_dense(r: Rows, nnp: int) -> np.ndarray:
    d = np.zeros((r.n_new, nnp), dtype=np.int64)
    for i in range(r.n_new):
        a, b = int(r.row_ptr[i]), int(r.row_ptr[i + 1])
        d[i, r.col[a:b]] = r.val[a:b]
    return d


@pytest.mark.parametrize("p", [7, 251, 32003, 65521, 2147483647])
@pytest.mark.parametrize("shape", [(30, 40, 0.2), (60, 90, 0.1)])
def test_oracle_matches_independent_dense_elimination(built, p, shape):
    from oracle import oracle
    s = synthetic_step(shape[0], shape[1], shape[2], p, seed=11 + p % 97, top_frac=0.7, share=3)
    got, _, _ = oracle.reduce(s)
    want = dense_rref_numpy(s)
    assert np.array_equal(rows_to_dense(got, s.n_nonpiv), want)


def test_oracle_output_is_canonical(built):
    """Rows monic, pivots strictly ascending, pivot columns cleared in all other rows, columns increasing."""
    from oracle import oracle
    s = synthetic_step(300, 420, 0.05, 32003, seed=3, top_frac=0.8, share=5)
    r, fma, _ = oracle.reduce(s)
    assert fma > 0 and r.n_new > 0
    d = rows_to_dense(r, s.n_nonpiv)
    piv = [int(r.col[int(r.row_ptr[i])]) for i in range(r.n_new)]
    assert piv == sorted(set(piv))
    for i in range(r.n_new):
        a, b = int(r.row_ptr[i]), int(r.row_ptr[i + 1])
        assert r.val[a] == 1 and np.all(np.diff(r.col[a:b].astype(np.int64)) > 0)
        assert np.all(r.val[a:b] > 0) and np.all(r.val[a:b] < s.p)
    for i, c in enumerate(piv):
        col = d[:, c].copy(); col[i] = 0
        assert not col.any()


def test_oracle_empty_and_degenerate(built):
    from oracle import oracle
    # no rows to reduce
    s = synthetic_step(20, 30, 0.2, 101, seed=1, top_frac=0.7, share=2)
    s0 = Step(s.p, s.n_top, 0, s.n_cols, s.row_ptr[: s.n_top + 1].copy(), s.col[: int(s.row_ptr[s.n_top])].copy(),
              s.coeff_id[: s.n_top].copy(), s.coeff_len, s.pool)
    r, _, _ = oracle.reduce(s0)
    assert r.n_new == 0 and list(r.row_ptr) == [0]
    # a row to reduce that is a copy of a pivot row reduces to zero
    t = int(s.n_top // 2)
    a, b = int(s.row_ptr[t]), int(s.row_ptr[t + 1])
    rp = np.concatenate([s.row_ptr[: s.n_top + 1], [s.row_ptr[s.n_top] + (b - a)]]).astype(np.uint64)
    col = np.concatenate([s.col[: int(s.row_ptr[s.n_top])], s.col[a:b]])
    cid = np.concatenate([s.coeff_id[: s.n_top], [s.coeff_id[t]]]).astype(np.uint32)
    s1 = Step(s.p, s.n_top, 1, s.n_cols, rp, col, cid, s.coeff_len, s.pool)
    r, _, _ = oracle.reduce(s1)
    assert r.n_new == 0

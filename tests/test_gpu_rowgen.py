"""Row generation of symbolic preprocessing on the device (gamba_cuda_rowgen_*; reference:
matrix_f4::create_multiplied_poly_matrix_row + the matrix' monomial table, source/matrix.hpp:387-444) against a plain
dictionary on the host: same monomial sets, rows that decode to the same monomials, dense handles (one per distinct
monomial; their order inside a batch is whatever the device's atomics made it - nothing depends on it), table growth
(rehash) and several batches per step. The end-to-end check is that the product driver, which
generates every row this way, reproduces the goldens (tests/test_gpu_e2e.py) and the oracle driver's bases
(tests/test_gpu_zz_named.py)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("nv,n_polys,max_terms,n_rows,seed", [(6, 5, 40, 60, 1), (15, 30, 900, 4000, 2), (40, 8, 300, 500, 3)])
def test_rows_match_a_dictionary(built, nv, n_polys, max_terms, n_rows, seed):
    from gamba_b200 import LinalgEngine
    rng = np.random.default_rng(seed)
    eng = LinalgEngine(32003)
    lib, h = eng.lib, eng.h
    stride = nv + 2
    w = rng.integers(1, 2 ** 32, size=nv, dtype=np.uint64) | 1

    def vec(e):      # [deg, deg_block, e_0 ..]: any layout works, only + and == matter
        v = np.zeros(stride, dtype=np.uint16); v[2:] = e; v[0] = e.sum(); return v

    def hsh(e):
        return np.uint32((w * e.astype(np.uint64)).sum() & 0xffffffff)

    assert lib.gamba_cuda_rowgen_begin(h, stride) == 0
    # the driver's basis table: monomials uploaded in two instalments, polynomials as handles into it
    polys, firsts, basis_ex, basis_hv = [], [], [], []
    for _ in range(n_polys):
        n = int(rng.integers(1, max_terms + 1))
        es = rng.integers(0, 4, size=(n, nv)).astype(np.uint16)
        ex = np.ascontiguousarray(np.stack([vec(e) for e in es]))
        hv = np.array([hsh(e) for e in es], dtype=np.uint32)
        mon = np.arange(len(basis_hv), len(basis_hv) + n, dtype=np.uint32)
        basis_ex.extend(ex); basis_hv.extend(hv)
        polys.append((ex, hv, mon))
    bex = np.ascontiguousarray(np.stack(basis_ex)); bhv = np.array(basis_hv, dtype=np.uint32)
    half = len(bhv) // 2
    first = C.c_uint64()
    assert lib.gamba_cuda_rowgen_add_monomials(h, half, ptr(bex[:half]), ptr(bhv[:half]), C.byref(first)) == 0 and first.value == 0
    assert lib.gamba_cuda_rowgen_add_monomials(h, len(bhv) - half, ptr(np.ascontiguousarray(bex[half:])), ptr(np.ascontiguousarray(bhv[half:])), C.byref(first)) == 0
    assert first.value == half
    nts = np.array([len(p_[1]) for p_ in polys], dtype=np.uint32)
    mptr = (C.c_void_p * n_polys)(*[p_[2].ctypes.data for p_ in polys])
    fts = np.zeros(n_polys, dtype=np.uint64)
    assert lib.gamba_cuda_rowgen_add_polys(h, n_polys, ptr(nts), mptr, ptr(fts)) == 0
    firsts = [int(x) for x in fts]
    table, rows_host, n_seen = {}, [], 0
    handles_dev = []
    for batch in range(3):                                    # several batches per step, the table grows in between
        nb = n_rows // 3
        which = rng.integers(0, n_polys, size=nb)
        qs = rng.integers(0, 3, size=(nb, nv)).astype(np.uint16)
        qex = np.ascontiguousarray(np.stack([vec(q) for q in qs]))
        qh = np.array([hsh(q) for q in qs], dtype=np.uint32)
        ft = np.array([firsts[i] for i in which], dtype=np.uint64)
        nt = np.array([len(polys[i][1]) for i in which], dtype=np.uint32)
        lead = np.zeros(nb, dtype=np.uint32)
        nmon = C.c_uint32()
        assert lib.gamba_cuda_rowgen_add_rows(h, nb, ptr(ft), ptr(nt), ptr(qex), ptr(qh), ptr(lead), C.byref(nmon)) == 0, \
            lib.gamba_cuda_last_error()
        for i, g in enumerate(which):
            row = [tuple((polys[g][0][k] + qex[i]).tolist()) for k in range(nt[i])]
            for m in row:
                table.setdefault(m, len(table))
            rows_host.append(row)
        assert nmon.value == len(table)                       # dense handles, one per distinct monomial
        # the monomials this batch introduced, with their hashes
        cnt = nmon.value - n_seen
        ex_new = np.zeros((max(cnt, 1), stride), dtype=np.uint16); hv_new = np.zeros(max(cnt, 1), dtype=np.uint32)
        assert lib.gamba_cuda_rowgen_fetch_monomials(h, n_seen, cnt, ptr(ex_new), ptr(hv_new)) == 0
        for k in range(cnt):
            handles_dev.append(tuple(ex_new[k].tolist()))
            assert hv_new[k] == hsh(ex_new[k][2:])
        n_seen = nmon.value
        base = len(rows_host) - nb
        for i in range(nb):
            assert handles_dev[lead[i]] == rows_host[base + i][0]
    assert sorted(handles_dev) == sorted(table)                # same monomial set, no duplicates
    dev = C.c_void_p(); n_ent = C.c_uint64()
    assert lib.gamba_cuda_rowgen_rows(h, C.byref(dev), C.byref(n_ent)) == 0
    assert n_ent.value == sum(len(r) for r in rows_host)
    flat = np.zeros(n_ent.value, dtype=np.uint32)
    assert lib.gamba_cuda_rowgen_download_rows(h, 0, n_ent.value, ptr(flat)) == 0
    o = 0
    for row in rows_host:                                      # every handle decodes to the monomial the host computed
        for m in row:
            assert handles_dev[flat[o]] == m
            o += 1
    # a new step starts from an empty table; the uploaded polynomials stay
    assert lib.gamba_cuda_rowgen_begin(h, stride) == 0
    lead = np.zeros(1, dtype=np.uint32); nmon = C.c_uint32()
    ft = np.array([firsts[0]], dtype=np.uint64); nt = np.array([len(polys[0][1])], dtype=np.uint32)
    q0 = np.ascontiguousarray(vec(np.zeros(nv, dtype=np.uint16))[None, :]); h0 = np.zeros(1, dtype=np.uint32)
    assert lib.gamba_cuda_rowgen_add_rows(h, 1, ptr(ft), ptr(nt), ptr(q0), ptr(h0), ptr(lead), C.byref(nmon)) == 0
    assert nmon.value == len({tuple(v.tolist()) for v in polys[0][0]}) and lead[0] < nmon.value
    eng.close()

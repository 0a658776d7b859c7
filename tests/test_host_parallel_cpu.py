"""Host driver: the data-parallel phases (row generation against the concurrent monomial table, reducer
search, column mapping; gamba_b200/host/par.hpp) must hand the engine the same matrices whatever the
number of threads — handle numbering is the only thing that may differ, and nothing depends on it."""
import filecmp
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD_IN, golden_bytes


@pytest.mark.parametrize("name", ["cyclic7-15", "kat10-15"])
def test_step_matrices_and_output_do_not_depend_on_threads(built, tmp_path, name):
    from oracle import oracle
    outs = []
    for t in (1, 4, 7):
        d = tmp_path / f"t{t}"
        d.mkdir()
        subprocess.run([oracle.DRIVER, "-i", os.path.join(GOLD_IN, name + ".txt"), "-o", str(d / "out.txt"), "-v", "0",
                        "-t", str(t), "--dump-steps", str(d)], check=True)
        outs.append(d)
    assert (outs[0] / "out.txt").read_bytes() == golden_bytes(name)
    files = sorted(os.path.basename(f) for f in glob.glob(str(outs[0] / "*.bin")))
    assert len(files) > 10
    for d in outs[1:]:
        assert sorted(os.path.basename(f) for f in glob.glob(str(d / "*.bin"))) == files
        for f in files:
            assert filecmp.cmp(outs[0] / f, d / f, shallow=False), f


def test_large_synthetic_generator_is_f4_shaped():
    """BASELINE.json configs[4] generator (vectorised): block form invariants the engine relies on."""
    from gamba_b200.stepfile import synthetic_step_large
    s = synthetic_step_large(4000, 6400, 0.005, 2147483647, seed=3)
    assert s.n_top == 3680 and s.n_bot == 320 and s.n_cols == 6400
    rp = s.row_ptr.astype(np.int64)
    for r in list(range(0, s.n_rows, 97)) + [s.n_top - 1, s.n_top, s.n_rows - 1]:
        c = s.col[rp[r]:rp[r + 1]].astype(np.int64)
        a, b = c[c < s.n_top], c[c >= s.n_top]
        assert np.all(np.diff(a) > 0) and np.all(np.diff(b) > 0)
        assert len(c) <= s.coeff_len[s.coeff_id[r]]
        if r < s.n_top:
            assert c[0] == r                      # pivot row r leads at column r
        else:
            assert c[0] < s.n_top                 # rows to reduce lead at a pivot column
    leads = s.col[rp[s.n_top:-1]]
    assert np.all(np.diff(leads.astype(np.int64)) <= 0)      # descending lead
    assert np.all(s.pool[np.cumsum(s.coeff_len) - s.coeff_len] == 1)   # monic


def test_large_synthetic_matches_oracle_rank(built):
    from gamba_b200.stepfile import synthetic_step_large
    from oracle import oracle
    s = synthetic_step_large(2500, 4000, 0.01, 32003, seed=5)
    rows, fma_top, _ = oracle.reduce(s)
    assert rows.n_new == min(s.n_bot, s.n_nonpiv) and fma_top > 0      # random rows: full rank


@pytest.mark.parametrize("name,extra", [("cyclic7-15", ()), ("kat10-15", ("--max-spairs", "0")), ("eco10-31", ())])
def test_batched_pair_update_equals_one_at_a_time(built, tmp_path, name, extra):
    """F4::update_batch installs a step's new generators as one batch over all host threads (Gebauer-Moeller criteria,
    reference source/update.hpp:29-269); GAMBA_UPDATE_BATCH=0 installs them one at a time. Same queue: same steps, pairs,
    criteria counts and basis."""
    import os
    import re
    import subprocess
    from conftest import GOLD_IN
    from oracle import oracle
    res = []
    for batch in ("0", "1"):
        out = tmp_path / f"out{batch}.txt"
        r = subprocess.run([oracle.DRIVER, "-i", os.path.join(GOLD_IN, name + ".txt"), "-o", str(out), "-v", "1", "-t", "4", *extra],
                           capture_output=True, text=True, env=dict(os.environ, GAMBA_UPDATE_BATCH=batch, GAMBA_ORACLE_THREADS="4"))
        assert r.returncode == 0, r.stderr
        m = re.search(r"steps (\d+)  pairs (\d+)  gm-removed (\d+)  rows-reduced (\d+)  zero-reductions (\d+)", r.stdout)
        res.append((m.groups(), out.read_bytes()))
    assert res[0] == res[1]

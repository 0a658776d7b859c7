"""Parity on the step matrices of BASELINE.json configs[1..3] — kat12-15 / eco12-15 (every step), cyclic9-15 (the bench
workload itself) and cyclic9-31 (steps >= 1 M nnz), the largest steps of kat15-15 and eco15-31 with the pair cap and
without it (the reference's two test modes, test/test_output.sh:15-30) — with the engine's DEFAULT options and ONE engine
driven through the steps in F4 order, so that the cost-model-chosen kernels (blocked tensor-core solve, sparse kernels,
Monte-Carlo echelon with the rank predicted from the previous step) are what is checked. Bit-exact against the oracle's
checker mode (oracle/oracle_checker.cpp, all host threads)."""
import glob
import hashlib
import os
import subprocess

import pytest

from conftest import ROOT
from gamba_b200.stepfile import load_step, thin_step

pytestmark = pytest.mark.gpu
GAMBA = os.path.join(ROOT, "gamba_b200", "bin", "gamba")


def check_steps_in_order(files, p, thin_rows_above=0):
    """One default-option engine through the steps in order; returns the solve modes it used."""
    from gamba_b200 import LinalgEngine
    from oracle import oracle
    eng = LinalgEngine(p)
    modes = []
    for f in files:
        s = load_step(f)
        if thin_rows_above and s.n_bot > thin_rows_above:
            s = thin_step(s, -(-s.n_bot // thin_rows_above))
        want, fma_top, _ = oracle.reduce(s, threads=oracle.host_threads())
        got = eng.reduce(s)
        assert got.same_as(want), f
        c = eng.counters()
        assert c["fma_top"] == fma_top, f
        assert eng.last["zero_reductions"] == s.n_bot - want.n_new
        modes.append(c["solve_mode"])
    eng.close()
    return modes


def step_files(d):
    return sorted(f for f in glob.glob(os.path.join(d, "step_*.bin")) if not f.endswith("step_final.bin"))


@pytest.mark.parametrize("name", ["kat12-15", "eco12-15"])
def test_config2_every_step_bit_exact(step_dirs, name):
    """BASELINE.json configs[1]: "per-F4-step D block checked bit-exact"."""
    files = step_files(step_dirs(name))
    assert len(files) >= 12
    modes = check_steps_in_order(files, 32003)
    assert 2 in modes or 1 in modes or 0 in modes


def test_config3_cyclic9_15_bench_matrices_default_options(product_step_dirs):
    """The ten matrices bench.py times (>= 15 M nnz), default options: the blocked tensor-core solve must be what runs."""
    files = step_files(product_step_dirs("cyclic9-15", ("--dump-min-nnz", "15000000")))
    assert len(files) == 10
    modes = check_steps_in_order(files, 32003)
    assert modes.count(2) >= 7, modes


def test_config3_cyclic9_31_steps_bit_exact(product_step_dirs):
    files = step_files(product_step_dirs("cyclic9-31", ("--dump-min-nnz", "1000000")))
    assert len(files) >= 20
    modes = check_steps_in_order(files, 2147483647)
    assert 2 in modes, modes


@pytest.mark.parametrize("name,extra", [("kat15-15", ()), ("kat15-15", ("--max-spairs", "0")),
                                        ("eco15-31", ()), ("eco15-31", ("--max-spairs", "0"))])
def test_config4_largest_steps_bit_exact(product_step_dirs, name, extra):
    """The three largest steps of the run; without the pair cap a step has 8-20 k rows to reduce: the check keeps all
    pivot rows and thins the rows to reduce to ~2500 (the oracle needs minutes for the full step)."""
    files = step_files(product_step_dirs(name, ("--dump-top", "3") + tuple(extra)))
    assert len(files) == 3
    check_steps_in_order(files, 32003 if name.endswith("-15") else 2147483647, thin_rows_above=2500)


def named_md5():
    out = {}
    with open(os.path.join(ROOT, "tests", "golden", "named_md5.txt")) as f:
        for line in f:
            if line.strip() and not line.startswith("#"):
                name, md5 = line.split()[:2]
                out[name] = md5
    return out


@pytest.mark.parametrize("name,extra", [("cyclic9-15", ()), ("cyclic9-31", ()), ("kat14-15", ()), ("eco14-15", ()),
                                        ("eco15-31", ()), ("kat15-15", ("--max-spairs", "0")), ("kat15-15", ())])
def test_named_systems_output_md5(built, tmp_path, name, extra):
    """No golden exists for the benchmark systems: the reduced Groebner basis is unique, so the product's
    `-o` file must equal the oracle driver's, whose md5 is committed (tests/golden/named_md5.txt, tools/make_named_md5.sh)."""
    from gamba_b200 import systems
    want = named_md5().get(name)
    if want is None:
        pytest.skip(f"no committed md5 for {name}")
    inp = systems.write_system(name, str(tmp_path / (name + ".txt")))
    out = tmp_path / "out.txt"
    r = subprocess.run([GAMBA, "-i", inp, "-o", str(out), "-v", "0", *extra], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:]
    assert hashlib.md5(out.read_bytes()).hexdigest() == want

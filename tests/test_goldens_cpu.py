"""The oracle is pinned end to end: host driver + CPU oracle engine must reproduce the reference's
MAGMA-generated golden bases (reference test/results/, harness test/test_output.sh:1-33) byte for
byte, with the default --max-spairs and with --max-spairs=0, like the reference's own tests."""
import os
import re
import subprocess

import pytest

from conftest import GOLD_IN, GOLD_OUT, golden_bytes

CASES = sorted(f[: -len(".out.txt.tar.gz")] for f in os.listdir(GOLD_OUT))


@pytest.mark.parametrize("case", CASES)
def test_driver_with_oracle_matches_golden(built, case, tmp_path):
    from oracle import oracle
    m = re.match(r"(.*)-elim(\d+)$", case)
    name, elim = (m.group(1), m.group(2)) if m else (case, "0")
    gold = golden_bytes(case)
    for extra in ([], ["--max-spairs=0"]):
        out = tmp_path / "out.txt"
        r = subprocess.run([oracle.DRIVER, "-i", os.path.join(GOLD_IN, name + ".txt"), "-e", elim, "-o", str(out),
                            "-v", "0", *extra], capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert out.read_bytes() == gold, (case, extra)


def test_generated_systems_match_reference_inputs(built, tmp_path):
    """gamba_b200.systems re-creates the benchmark families: same basis as the shipped input file."""
    from gamba_b200 import systems
    from oracle import oracle
    for name in ["cyclic5-15", "cyclic6-15", "cyclic7-31", "kat8-15", "kat9-8", "eco10-31", "eco11-16"]:
        inp = systems.write_system(name, str(tmp_path / (name + ".txt")))
        out = tmp_path / "o.txt"
        subprocess.run([oracle.DRIVER, "-i", inp, "-o", str(out), "-v", "0"], check=True, timeout=600)
        assert out.read_bytes() == golden_bytes(name), name

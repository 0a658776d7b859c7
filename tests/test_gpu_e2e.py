"""End to end through the product binary (host driver + CUDA engine behind the C ABI): the `-o` file is
byte-identical to the reference's golden basis (BASELINE.json configs[0]: cyclic6-15, and more)."""
import os
import subprocess

import pytest

from conftest import GOLD_IN, ROOT, golden_bytes

pytestmark = pytest.mark.gpu
GAMBA = os.path.join(ROOT, "gamba_b200", "bin", "gamba")


@pytest.mark.parametrize("name,extra", [
    ("cyclic6-15", []), ("cyclic6-15", ["--max-spairs=0"]), ("cyclic6-31", []), ("cyclic7-31", []), ("cyclic7-8", []),
    ("kat9-16", []), ("kat10-15", []), ("eco10-31", ["--max-spairs=0"]), ("eco11-8", []), ("noon6-8", []),
    ("one-ideal", []), ("zero-ideal", []), ("henrion5-31", []), ("cyclic7-3", []),
])
def test_gamba_output_matches_golden(built, tmp_path, name, extra):
    out = tmp_path / "out.txt"
    r = subprocess.run([GAMBA, "-i", os.path.join(GOLD_IN, name + ".txt"), "-o", str(out), "-v", "1", *extra],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    assert "engine = cuda-b200" in r.stdout
    assert out.read_bytes() == golden_bytes(name)


def test_gamba_elimination_order(built, tmp_path):
    out = tmp_path / "out.txt"
    subprocess.run([GAMBA, "-i", os.path.join(GOLD_IN, "cyclic5-15.txt"), "-e", "2", "-o", str(out), "-v", "0"], check=True)
    assert out.read_bytes() == golden_bytes("cyclic5-15-elim2")

"""Drop-in proof on the GPU: the reference's OPEN driver (compiled from /root/reference by oracle/build_ref.sh in the build
container; the binary oracle/_ref/gamba_ref travels to the GPU box) calls libgamba_cuda.so through the adapter headers of
integration/reference_adapter/ — its own f4_main, extract_new_rows (NUM_ROWS_BUFFER layout, source/matrix.hpp:799-895),
insert_new_rows_echelon / insert_new_rows_reduce (source/basis.hpp:652-769) consume the engine's output, and the basis it
writes is byte-identical to the reference's goldens. uint8 / uint16 / uint32 instantiations, Montgomery form in and out."""
import os
import subprocess

import pytest

from conftest import GOLD_IN, ROOT, golden_bytes

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref", "gamba_ref")


@pytest.mark.parametrize("name,extra", [
    ("cyclic6-15", []), ("cyclic6-15", ["--max-spairs=0"]), ("kat9-16", []), ("cyclic7-31", []), ("cyclic7-31", ["--max-spairs=0"]),
    ("cyclic7-8", []), ("cyclic7-3", []), ("kat10-15", []), ("eco10-31", []), ("eco11-16", []), ("noon6-31", []), ("one-ideal", []),
])
def test_reference_driver_on_cuda_engine_matches_goldens(built, tmp_path, name, extra):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/gamba_ref was not built (needs the reference sources in the build container)")
    out = tmp_path / "out.txt"
    r = subprocess.run([REF, "-i", os.path.join(GOLD_IN, name + ".txt"), "-o", str(out), *extra],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_bytes() == golden_bytes(name)


def test_reference_driver_elimination_order_on_cuda_engine(built, tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/gamba_ref was not built")
    out = tmp_path / "out.txt"
    subprocess.run([REF, "-i", os.path.join(GOLD_IN, "cyclic5-15.txt"), "-e", "2", "-o", str(out)], check=True, timeout=600)
    assert out.read_bytes() == golden_bytes("cyclic5-15-elim2")

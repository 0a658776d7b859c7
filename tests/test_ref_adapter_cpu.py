"""The reference's OPEN driver (its own f4_main / matrix_f4::extract_new_rows / insert_new_rows_reduce, compiled from
/root/reference where it lies by oracle/build_ref.sh) running on the two adapter headers of integration/reference_adapter/
— here with the CPU oracle behind the C ABI (oracle/_ref/libcabi_oracle.so), so the adapters themselves (block-form
flattening, Montgomery form in and out for uint8 / uint16 / uint32 coefficients, the NUM_ROWS_BUFFER column-major repack,
new[] / allocate_polynomial ownership) are checked against the reference's goldens without a GPU. The same binary linked
with libgamba_cuda.so is tests/test_gpu_ref_adapter.py. Skipped when neither the reference sources nor a prebuilt binary exist."""
import os
import subprocess

import pytest

from conftest import GOLD_IN, ROOT, golden_bytes

REF_CPU = os.path.join(ROOT, "oracle", "_ref", "gamba_ref_cpu")


@pytest.fixture(scope="module")
def ref_cpu():
    if os.path.isdir("/root/reference/source"):
        subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True, capture_output=True)
    if not os.path.exists(REF_CPU):
        pytest.skip("no reference sources and no prebuilt oracle/_ref/gamba_ref_cpu")
    return REF_CPU


# uint32 (31-bit), uint16 (15-/16-bit) and uint8 (8-, 3-, 2-bit primes) instantiations of the reference's templates
@pytest.mark.parametrize("name,extra", [
    ("cyclic6-15", []), ("cyclic6-15", ["--max-spairs=0"]), ("cyclic6-31", []), ("cyclic6-16", []), ("cyclic6-2", []),
    ("cyclic7-8", []), ("cyclic7-3", ["--max-spairs=0"]), ("kat9-16", []), ("kat8-15", []), ("kat7-31", []), ("eco10-31", []),
    ("noon6-8", []), ("henrion5-31", []), ("one-ideal", []), ("zero-ideal", []),
])
def test_reference_driver_on_the_adapters_matches_goldens(ref_cpu, tmp_path, name, extra):
    out = tmp_path / "out.txt"
    r = subprocess.run([ref_cpu, "-i", os.path.join(GOLD_IN, name + ".txt"), "-o", str(out), *extra],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_bytes() == golden_bytes(name)


def test_reference_driver_elimination_order(ref_cpu, tmp_path):
    out = tmp_path / "out.txt"
    subprocess.run([ref_cpu, "-i", os.path.join(GOLD_IN, "cyclic5-15.txt"), "-e", "2", "-o", str(out)], check=True, timeout=600)
    assert out.read_bytes() == golden_bytes("cyclic5-15-elim2")

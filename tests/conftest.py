import os
import subprocess
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD_IN = os.path.join(ROOT, "tests", "golden", "inputs")
GOLD_OUT = os.path.join(ROOT, "tests", "golden", "outputs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_bytes(name: str) -> bytes:
    with tarfile.open(os.path.join(GOLD_OUT, name + ".out.txt.tar.gz")) as tf:
        return tf.extractfile(tf.getmembers()[0]).read()


@pytest.fixture(scope="session")
def built():
    """Build the checker (oracle/) and, when missing, the product. The .so files are prebuilt and
    travel to the GPU box; make is a no-op there."""
    from oracle import oracle
    oracle.build()
    lib = os.path.join(ROOT, "gamba_b200", "libgamba_cuda.so")
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, "gamba_b200", "bin", "gamba")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "gamba_b200"), "-j8"], check=True, capture_output=True)
    return True


@pytest.fixture(scope="session")
def step_dirs(built, tmp_path_factory):
    """Per system: directory of step matrices + oracle results, produced by the oracle driver."""
    from gamba_b200 import systems
    from oracle import oracle
    cache = {}

    def get(name: str, extra=()):
        key = (name,) + tuple(extra)
        if key not in cache:
            d = str(tmp_path_factory.mktemp("steps_" + name.replace("-", "_")))
            gold_in = os.path.join(GOLD_IN, name + ".txt")
            inp = gold_in if os.path.exists(gold_in) else systems.write_system(name, os.path.join(d, name + ".txt"))
            subprocess.run([oracle.DRIVER, "-i", inp, "-o", os.path.join(d, "out.txt"), "-v", "0", "--dump-steps", d,
                            *extra], check=True, env=dict(os.environ, GAMBA_ORACLE_THREADS=str(oracle.host_threads())))
            cache[key] = d
        return cache[key]
    return get


@pytest.fixture(scope="session")
def product_step_dirs(built, tmp_path_factory):
    """Step matrices of the big named systems (BASELINE.json configs[1..3]), dumped by the PRODUCT driver on the GPU
    (the oracle driver needs minutes to hours for them); the tests check every dumped step against the oracle."""
    from gamba_b200 import systems
    cache = {}

    def get(name: str, extra=()):
        key = (name,) + tuple(extra)
        if key not in cache:
            d = str(tmp_path_factory.mktemp("psteps_" + name.replace("-", "_")))
            inp = systems.write_system(name, os.path.join(d, name + ".txt"))
            r = subprocess.run([os.path.join(ROOT, "gamba_b200", "bin", "gamba"), "-i", inp, "-o", os.path.join(d, "out.txt"),
                                "-v", "0", "--dump-steps", d, *extra], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            cache[key] = d
        return cache[key]
    return get

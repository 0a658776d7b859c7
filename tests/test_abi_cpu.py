"""The C-ABI library loads and exports every symbol include/gamba_cuda.h declares; without a GPU the
product fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "gamba_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gamba_cuda_[a-z0-9_]+)\s*\(", text)) - {"gamba_cuda_allgather_fn"})


def test_header_symbols_exported(built):
    from gamba_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SYMBOLS) == names
    assert lib.gamba_cuda_abi_version() == 4


def test_struct_sizes_match_header(built, tmp_path):
    """ctypes mirrors of the ABI structs have the sizes the C compiler gives the header's structs."""
    from gamba_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gamba_cuda.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(gamba_cuda_config),sizeof(gamba_cuda_step_in),sizeof(gamba_cuda_step_out),'
                   'sizeof(gamba_cuda_counters));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.Config), C.sizeof(_lib.StepIn), C.sizeof(_lib.StepOut), C.sizeof(_lib.Counters)]


def test_create_rejects_bad_arguments(built):
    from gamba_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    for p, cb in [(2, 4), (4, 4), (2 ** 31 + 11, 4), (32003, 3), (65521, 1)]:
        cfg = _lib.Config(p, cb, 0, 0, 1, 0, 1)
        assert lib.gamba_cuda_create(C.byref(cfg), C.byref(h)) != 0
        assert lib.gamba_cuda_last_error()


def test_no_cpu_fallback_without_gpu(built):
    """On a box without a CUDA device the engine refuses to be created (and the driver exits non-zero)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gamba_b200 import LinalgEngine
    with pytest.raises(RuntimeError):
        LinalgEngine(32003)
    inp = os.path.join(ROOT, "tests", "golden", "inputs", "cyclic4-15.txt")
    r = subprocess.run([os.path.join(ROOT, "gamba_b200", "bin", "gamba"), "-i", inp, "-v", "0"], capture_output=True, text=True)
    assert r.returncode != 0 and "gamba_cuda_create" in r.stderr


def test_product_does_not_link_the_oracle(built):
    """Neither the library nor the driver binary contains oracle code."""
    for f in ("libgamba_cuda.so", os.path.join("bin", "gamba")):
        out = subprocess.run(["nm", "-C", "--defined-only", os.path.join(ROOT, "gamba_b200", f)], capture_output=True, text=True).stdout
        assert "OracleEngine" not in out and "gamba_oracle" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "gamba_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle_engine" not in text, f

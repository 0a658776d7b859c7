"""Input grammar / output layout of `gamba -i/-o` (reference source/io.cpp:46-449)."""
import subprocess

import pytest


def run(built_driver, text, tmp_path, *args):
    inp = tmp_path / "in.txt"
    inp.write_text(text)
    out = tmp_path / "out.txt"
    r = subprocess.run([built_driver, "-i", str(inp), "-o", str(out), "-v", "0", *args], capture_output=True, text=True)
    return r, (out.read_text() if out.exists() else None)


@pytest.fixture
def drv(built):
    from oracle import oracle
    return oracle.DRIVER


def test_output_layout(drv, tmp_path):
    r, out = run(drv, "x, y,z\n32003\nx^2 + y*z - 1,\n 2*x*y - 3/2*z ,\n\nz^2*y+x\n", tmp_path)
    assert r.returncode == 0, r.stderr
    lines = out.split("\n")
    assert lines[0] == "x,y,z," and lines[1] == "32003"
    body = [l for l in lines[2:] if l]
    assert all(l.endswith(",") for l in body[:-1]) and not body[-1].endswith(",")
    assert "-" not in "".join(body) and " " not in "".join(body)


def test_rational_and_repeated_variables(drv, tmp_path):
    a, out_a = run(drv, "x,y\n101\nx*x*y^2-1/3,\ny-2\n", tmp_path)
    b, out_b = run(drv, "x,y\n101\n3*x^2*y^2-1,\ny-2\n", tmp_path)
    assert a.returncode == 0 and b.returncode == 0 and out_a == out_b


def test_longest_variable_name_first(drv, tmp_path):
    r, out = run(drv, "x,x1,x11\n32003\nx11*x1-x,\nx1^2-1\n", tmp_path)
    assert r.returncode == 0 and out.startswith("x,x1,x11,\n32003\n")


@pytest.mark.parametrize("text,msg", [
    ("", "Input file is empty."),
    ("x,y\n", "Missing line containing field characteristic."),
    ("x,y\nabc\nx\n", "Field characteristic not valid."),
    ("x,y\n32004\nx\n", "Field characteristic is not a prime number."),
    ("x,y\n2\nx\n", "Field characteristic = 2 not supported."),
    ("x,y\n4294967291\nx\n", "Field characteristic too large."),
    ("x,x\n7\nx\n", "Duplicated variable name: x."),
    ("x,1y\n7\nx\n", "Invalid variable name: 1y."),
    ("x,,y\n7\nx\n", "Empty variable name."),
    ("x,y\n7\nx+*y\n", "Error parsing generator number 1."),
    ("x,y\n7\nx,\nx*w\n", "Error parsing generator number 2."),
    ("x,y\n7\nx,\n3/14*y\n", "Division by zero in generator number 2."),
    ("x,y\n7\nx^40000\n", "Monomial exponent in generator 1 must be >= 0 or < 32768."),
])
def test_errors(drv, tmp_path, text, msg):
    r, _ = run(drv, text, tmp_path)
    assert r.returncode != 0
    assert msg in r.stderr


def test_elimination_block_validation(drv, tmp_path):
    r, _ = run(drv, "x,y\n7\nx*y-1\n", tmp_path, "-e", "2")
    assert r.returncode != 0 and "Elimination block size" in r.stderr


def test_zero_and_one_ideal(drv, tmp_path):
    r, out = run(drv, "x,y\n7\nx-x\n", tmp_path)
    assert r.returncode == 0 and out == "x,y,\n7\n"
    r, out = run(drv, "x,y\n7\nx,\nx-1\n", tmp_path)
    assert r.returncode == 0 and out == "x,y,\n7\n1\n"


def test_parallel_writer_equals_the_serial_one(drv, tmp_path):
    """A basis of more than 200 k terms (kat11-15: 383 k) is formatted by all host threads into per-chunk buffers
    (gamba_b200/host/io.cpp); GAMBA_WRITE_THREADS=1 takes the serial path. Same bytes."""
    import os
    from gamba_b200 import systems
    inp = systems.write_system("kat11-15", str(tmp_path / "kat11.txt"))
    outs = []
    for threads in ("1", "5"):
        out = tmp_path / f"out{threads}.txt"
        r = subprocess.run([drv, "-i", inp, "-o", str(out), "-v", "0"], capture_output=True, text=True,
                           env=dict(os.environ, GAMBA_WRITE_THREADS=threads, GAMBA_ORACLE_THREADS="8"))
        assert r.returncode == 0, r.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1] and outs[0].count(b"+") + outs[0].count(b",") > 200000

"""gamba_b200/systems.py re-creates the benchmark systems from their definitions (the reference ships them as
examples/*.txt, absent on the GPU box): for every system that also exists among the committed golden inputs the
generated text must describe the same polynomials (same variables, prime and terms; term order and trailing commas may differ)."""
import os
import re
from fractions import Fraction

import pytest

from conftest import GOLD_IN
from gamba_b200 import systems


def parse(text):
    lines = text.strip().split("\n")
    names = [v.strip() for v in lines[0].split(",") if v.strip()]
    p = int(lines[1].strip())
    body = "".join("".join(lines[2:]).split())
    polys = []
    for g in body.split(","):
        if not g:
            continue
        terms = {}
        for t in re.findall(r"[+-]?[^+-]+", g):
            sign = -1 if t[0] == "-" else 1
            t = t.lstrip("+-")
            coef, mon = Fraction(1), {}
            for f in t.split("*"):
                m = re.fullmatch(r"(\d+)(?:/(\d+))?", f)
                if m:
                    coef *= Fraction(int(m.group(1)), int(m.group(2) or 1))
                else:
                    v, _, e = f.partition("^")
                    mon[v] = mon.get(v, 0) + int(e or 1)
            key = tuple(sorted(mon.items()))
            terms[key] = terms.get(key, 0) + sign * coef
        polys.append({k: v for k, v in terms.items() if v != 0})
    return names, p, polys


# the reference's own file names are not uniform for the smallest primes (cyclic6-2 uses other variable names than cyclic6-15,
# kat10-3 means p = 3 while cyclic7-3 means the 3-bit prime 7): those two are read from the committed inputs, never generated
NAMES = sorted(f[:-4] for f in os.listdir(GOLD_IN) if re.fullmatch(r"(cyclic|kat|eco)\d+-\d+\.txt", f) and f[:-4] not in ("cyclic6-2", "kat10-3"))


@pytest.mark.parametrize("name", NAMES)
def test_generated_system_equals_the_committed_input(name):
    want = parse(open(os.path.join(GOLD_IN, name + ".txt")).read())
    got = parse(systems.system_text(name))
    assert got[0] == want[0] and got[1] == want[1]
    assert got[2] == want[2]


def test_benchmark_systems_are_generated():
    for name in ("cyclic9-15", "cyclic9-31", "kat12-15", "eco12-15", "kat15-15", "kat15-31", "eco15-31"):
        names, p, polys = parse(systems.system_text(name))
        n = int(re.search(r"\d+", name).group())
        assert len(names) == n and len(polys) == n and p == systems.PRIMES[name.rsplit("-", 1)[1]]

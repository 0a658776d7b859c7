"""Generators for the benchmark polynomial systems named in BASELINE.json.

The reference ships them as text files (examples/cyclic9-15.txt, kat12-15.txt,
eco12-15.txt ...). Those files do not exist on the GPU box, and the systems are
standard families, so they are re-created here from their definitions, with the
variable names of the reference's files (the reduced Groebner basis — and hence
the `-o` file — depends only on the ideal, the variable order and p).
`tests/test_systems_cpu.py` checks the generated text against the committed golden inputs (same polynomials).
"""
from __future__ import annotations

PRIMES = {"2": 3, "3": 7, "8": 251, "15": 32003, "16": 65521, "31": 2147483647}


def _fmt(names, p, polys):
    return ",".join(names) + "\n" + str(p) + "\n" + ",\n".join(polys) + "\n"


def cyclic(n: int, p: int) -> str:
    names = list("xyztuv") if n == 6 else [chr(ord("a") + i) for i in range(n)]
    polys = []
    for k in range(1, n):
        terms = ["*".join(names[(i + j) % n] for j in range(k)) for i in range(n)]
        polys.append("+".join(terms))
    polys.append("*".join(names) + "-1")
    return _fmt(names, p, polys)


def katsura(n: int, p: int) -> str:
    """`katN` of the reference: N variables x1..xN (Katsura-(N-1))."""
    names = [f"x{i}" for i in range(1, n + 1)]

    def u(l):
        l = abs(l)
        return names[l] if l < n else None

    polys = [names[0] + "".join(f"+2*{v}" for v in names[1:]) + "-1"]
    for m in range(0, n - 1):
        coef = {}
        for l in range(-(n - 1), n):
            a, b = u(l), u(m - l)
            if a is None or b is None:
                continue
            key = tuple(sorted((a, b), key=lambda s: int(s[1:])))
            coef[key] = coef.get(key, 0) + 1
        terms = []
        for (a, b), c in coef.items():
            mon = f"{a}^2" if a == b else f"{a}*{b}"
            terms.append(mon if c == 1 else f"{c}*{mon}")
        polys.append("+".join(terms) + f"-{names[m]}")
    return _fmt(names, p, polys)


def eco(n: int, p: int) -> str:
    names = [f"x{i}" for i in range(n)]
    last = names[n - 1]
    polys = []
    for k in range(1, n):
        terms = [f"{names[i]}*{names[i + k]}*{last}" for i in range(0, n - 1 - k)]
        terms.append(f"{names[k - 1]}*{last}")
        polys.append("+".join(terms) + f"-{k}")
    polys.append("+".join(names[: n - 1]) + "+1")
    return _fmt(names, p, polys)


FAMILIES = {"cyclic": cyclic, "kat": katsura, "eco": eco}


def system_text(name: str) -> str:
    """`cyclic9-15`, `kat12-15`, `eco15-31` ... -> input file text."""
    stem, bits = name.rsplit("-", 1)
    for fam, fn in FAMILIES.items():
        if stem.startswith(fam) and stem[len(fam):].isdigit():
            return fn(int(stem[len(fam):]), PRIMES[bits])
    raise KeyError(name)


def write_system(name: str, path: str) -> str:
    with open(path, "w") as f:
        f.write(system_text(name))
    return path

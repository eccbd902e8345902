"""Evaluate the reference's generated ``odes.f90`` (GETYDOT) from Python.

Test-infrastructure only, and only usable in the authoring container where
``/root/reference`` exists.  The Fortran text is *read and interpreted*, never
copied into this repo: continuation lines are joined and each assignment is
translated to a Python expression (``RATE(12)`` -> ``RATE[12]``), then executed
with 1-based array shims.  This gives us the reference's own RHS arithmetic
(same operation order, IEEE doubles) to pin the oracle's table-driven RHS and
to generate the golden vectors in ``tests/golden/``.
"""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np

REF_ODES = Path("/root/reference/src/fortran_src/odes.f90")


class _OneBased:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, i):
        return self.a[i - 1]

    def __setitem__(self, i, v):
        self.a[i - 1] = v


def _join(text: str):
    cur = ""
    for raw in text.splitlines():
        line = raw
        if cur:
            s = line.lstrip()
            if s.startswith("&"):
                s = s[1:]
            line = s
        if line.rstrip().endswith("&"):
            cur += line.rstrip()[:-1]
            continue
        yield cur + line
        cur = ""


def compile_getydot(path: Path = REF_ODES):
    src = []
    body = False
    indent = "    "
    for line in _join(path.read_text()):
        s = line.strip()
        if s.upper().startswith("SUBROUTINE GETYDOT"):
            body = True
            continue
        if not body or not s or s.startswith("!") or s.upper().startswith("REAL("):
            continue
        if s.upper().startswith("END SUBROUTINE"):
            break
        up = s.upper()
        if up.startswith("IF ") and up.endswith("THEN"):
            cond = s[s.index("(") + 1: s.rindex(")")]
            cond = cond.replace(".lt.", "<").replace(".LT.", "<")
            cond = re.sub(r"\b(RATE|Y|YDOT)\((\d+)\)", r"\1[\2]", cond)
            src.append(f"{indent}if {cond}:")
            indent = "        "
            continue
        if up == "ELSE":
            src.append("    else:")
            continue
        if up == "ENDIF":
            indent = "    "
            continue
        s = re.sub(r"\b(RATE|Y|YDOT)\((\d+)\)", r"\1[\2]", s)
        s = re.sub(r"\bMIN\(", "min(", s)
        src.append(indent + s)
    code = (
        "def getydot(RATE, Y, bulkLayersReciprocal, surfaceCoverage, safeMantle, safeBulk, D, YDOT):\n"
        "    safebulk = safeBulk\n"
        + "\n".join(src)
        + "\n    return SURFGROWTHUNCORRECTED\n"
    )
    ns: dict = {}
    exec(compile(code, "<odes.f90>", "exec"), ns)
    fn = ns["getydot"]

    def call(rate, y, blr, cov, safe_mantle, safe_bulk, dens):
        ydot = np.zeros(len(y), dtype=np.float64)
        fn(_OneBased([float(v) for v in rate]), _OneBased([float(v) for v in y]), float(blr), float(cov),
           float(safe_mantle), float(safe_bulk), float(dens), _OneBased(ydot))
        return ydot

    return call


if __name__ == "__main__":
    f = compile_getydot()
    rng = np.random.default_rng(0)
    y = 10 ** rng.uniform(-12, -4, 336)
    y[335] = 1e4
    rate = 10 ** rng.uniform(-12, -9, 3203)
    print(f(rate, y, 0.5, 0.1, 1e-6, 1e-5, 1e4)[:5])

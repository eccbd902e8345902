"""The reference's text outputs (``io.f90``), host side.

``write_full_output``   -- ``outputFile``: header ``io.f90:27-29`` (format 335), one row per output
                           time ``io.f90:80-83`` (format 8020)
``write_abundances``    -- ``abundSaveFile``: one line of final abundances, ``io.f90:48-56`` (format 8010)
``read_abundances``     -- ``abundLoadFile``: list-directed read of that line, ``io.f90:36-46``
``write_column_output`` -- ``columnFile``: header ``io.f90:23-24`` (format 333), selected species every
                           ``writeStep``-th output ``io.f90:108-118`` (format 8030)
``write_rate_output``   -- ``rateFile``: no header, physics + every rate coefficient per output time
                           ``io.f90:94-97`` (format 8021, three-digit exponents)
``read_output_file``    -- what ``uclchem.analysis.read_output_file`` returns for a full output file

The rows carry six significant digits like the reference's (``1pe15.5``); Fortran writes a
three-digit exponent without the ``E`` (``1.00000-100``), reproduced by :func:`_e`.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

PHYSICS_HEADER = "Time,Density,gasTemp,dustTemp,Av,radfield,zeta,point,"


def _e(x: float, width: int, digits: int) -> str:
    """Fortran ``1pe<width>.<digits>`` (two exponent digits; ``d.ddddd-100`` style beyond that)."""
    s = f"{x:.{digits}E}"
    mant, exp = s.split("E")
    e = int(exp)
    if abs(e) >= 100:
        s = f"{mant}{'+' if e >= 0 else '-'}{abs(e):03d}"
    return s.rjust(width)


def _species_header(species) -> str:
    w = max(len(s) for s in species)          # specName is CHARACTER(LEN=<longest name>)
    return ",".join(s.ljust(w) for s in species)


def format_row(physics_row, abund_row) -> str:
    """One line of the full output (format 8020): time, density, gas T, dust T, Av, radfield, zeta,
    point, then every species."""
    t, dens, tg, td, av, rad, zeta, point = physics_row[:8]
    head = (f"{_e(t, 11, 3)},{_e(dens, 11, 4)},{tg:8.2f},{td:8.2f},{_e(av, 11, 4)},{_e(rad, 11, 4)},"
            f"{_e(zeta, 11, 4)},{int(point):4d},")
    return head + ",".join(_e(v, 15, 5) for v in abund_row)


def write_full_output(path, species, physics, abund) -> None:
    """physics [nrows, 8], abund [nrows, nspec] (the trimmed trajectory of one model, point = 1)."""
    lines = [PHYSICS_HEADER + _species_header(species)]
    lines += [format_row(p, a) for p, a in zip(np.asarray(physics), np.asarray(abund))]
    Path(path).write_text("\n".join(lines) + "\n")


def _e3(x: float, width: int, digits: int) -> str:
    """Fortran ``1pe<width>.<digits>e3``: always a three-digit exponent with the ``E``."""
    mant, exp = f"{x:.{digits}E}".split("E")
    e = int(exp)
    return f"{mant}E{'+' if e >= 0 else '-'}{abs(e):03d}".rjust(width)


def _physics7(physics_row) -> str:
    t, dens, tg, td, av, rad, zeta = physics_row[:7]
    return (f"{_e(t, 11, 3)},{_e(dens, 11, 4)},{tg:8.2f},{td:8.2f},{_e(av, 11, 4)},{_e(rad, 11, 4)},"
            f"{_e(zeta, 11, 4)},")


def write_column_output(path, species, out_species, physics, abund, write_step: int = 1) -> None:
    """columnFile: the species of `out_species` every `write_step`-th output.  The reference's counter starts at
    zero (chemistry.f90:28) and is compared before it is advanced (io.f90:110-117), so the initial state is not
    written and, with writeStep = 1, every later row is."""
    species = list(species)
    idx = [species.index(sp) for sp in out_species]
    w = max(len(sp) for sp in species)
    lines = ["Time,Density,gasTemp,dustTemp,av,radfield,zeta," + ",".join(sp.ljust(w) for sp in out_species)]
    counter = 0
    for p_, a_ in zip(np.asarray(physics), np.asarray(abund)):
        if counter == write_step:
            counter = 1
            lines.append(_physics7(p_) + ",".join(_e(a_[i], 15, 5) for i in idx))
        else:
            counter += 1
    Path(path).write_text("\n".join(lines) + "\n")


def write_rate_output(path, physics, rates) -> None:
    """rateFile: one row per output time, physics columns then every rate coefficient (format 8021)."""
    lines = [_physics7(p_) + f"{int(p_[7]):4d}," + ",".join(_e3(v, 15, 5) for v in r_)
             for p_, r_ in zip(np.asarray(physics), np.asarray(rates))]
    Path(path).write_text("\n".join(lines) + "\n")


def write_abundances(path, abund) -> None:
    Path(path).write_text(",".join(_e(v, 15, 5) for v in np.asarray(abund).ravel()) + "\n")


def read_abundances(path, nspec=None) -> np.ndarray:
    txt = Path(path).read_text().replace("\n", " ")
    vals = np.array([float(_fix_exp(v)) for v in txt.replace(",", " ").split()])
    return vals if nspec is None else vals[:nspec]


def _fix_exp(tok: str) -> str:
    """``1.00000-100`` -> ``1.00000E-100``"""
    tok = tok.strip()
    for i in range(1, len(tok)):
        if tok[i] in "+-" and tok[i - 1] not in "eEdD":
            return tok[:i] + "E" + tok[i:]
    return tok.replace("D", "E").replace("d", "e")


def read_output_file(path):
    """(column names, table) of a full output file; names stripped like analysis.read_output_file."""
    lines = Path(path).read_text().splitlines()
    names = [c.strip() for c in lines[0].split(",")]
    data = np.array([[float(_fix_exp(v)) for v in ln.split(",")] for ln in lines[1:] if ln.strip()])
    return names, data

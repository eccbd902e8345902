"""io.f90 text formats on the host (uclchem_b200/datio.py), pinned on an excerpt of the reference's own
example output (tests/golden/static_full_excerpt.dat: header, first three rows and the 1 Myr row of
examples/example-output/static-full.dat)."""
import numpy as np
import pytest
from conftest import GOLDEN

from uclchem_b200 import datio


@pytest.fixture(scope="module")
def excerpt():
    return (GOLDEN / "static_full_excerpt.dat").read_text().splitlines()


def test_full_output_rows_are_byte_identical(net, excerpt, tmp_path):
    g = np.load(GOLDEN / "static_full.npz")           # the same file, parsed (tools/make_golden.py)
    rows = [0, 1, 2, len(g["physics"]) - 1]
    out = tmp_path / "full.dat"
    datio.write_full_output(out, net.names, g["physics"][rows], g["abund"][rows])
    got = out.read_text().splitlines()
    assert got[0] == excerpt[0]                        # header: io.f90 format 335, names padded to LEN(specName)
    assert got[1:] == excerpt[1:]                      # rows: io.f90 format 8020
    names, data = datio.read_output_file(out)
    assert names[:8] == ["Time", "Density", "gasTemp", "dustTemp", "Av", "radfield", "zeta", "point"]
    assert names[8:] == net.names and np.array_equal(data[:, 8:], g["abund"][rows])


def test_abundance_file_round_trip(net, tmp_path):
    v = np.load(GOLDEN / "startstatic.npy")
    f = tmp_path / "start.dat"
    datio.write_abundances(f, v)
    txt = f.read_text()
    assert txt.count("\n") == 1 and len(txt.split(",")) == net.nspec and txt.startswith("    1.97293E-05,")
    assert np.array_equal(datio.read_abundances(f, net.nspec), v)


def test_three_digit_exponents():
    assert datio._e(1e-100, 15, 5) == "    1.00000-100" and datio._e(1.0e-30, 15, 5) == "    1.00000E-30"
    assert datio._fix_exp("1.00000-100") == "1.00000E-100" and float(datio._fix_exp(" 2.5E+03")) == 2500.0
    assert datio._e(0.0, 11, 3) == "  0.000E+00"


def test_disk_mode_refuses_to_mix_with_memory_mode():
    from uclchem_b200 import model
    with pytest.raises(RuntimeError, match="Offending keys"):
        model.pre_flight_checklist(True, False, False, None, {"outputfile": "a.dat"})
    model.pre_flight_checklist(False, False, False, None, {"outputfile": "a.dat"})   # disk mode: fine

"""io.f90 text formats on the host (uclchem_b200/datio.py), pinned on an excerpt of the reference's own
example output (tests/golden/static_full_excerpt.dat: header, first three rows and the 1 Myr row of
examples/example-output/static-full.dat)."""
from pathlib import Path

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


class _OracleBackedLibrary:
    """Test double for uclchem_b200._capi.Library: same `run_grid` contract, computed by the oracle.
    Lets the host-side logic of uclchem_b200.model (argument handling, disk mode, return tuples) run on
    the CPU; the product path never sees it."""

    def __init__(self, oracle, net):
        from uclchem_b200._capi import STAT_FIELDS
        self.orc, self.net = oracle, net
        self.nspec, self.neq, self.nreac, self.species = net.nspec, net.neq, net.nreac, list(net.names)
        self._nstat, self._iint = len(STAT_FIELDS), STAT_FIELDS.index("nintervals")

    def run_grid(self, kind, params, y0=None, timepoints=0, want_physics=False, want_chem=False, want_rates=False,
                 step_budget=0, coefficients=None, pp_grid=None, pp_coldens=False):
        assert not coefficients, "the double has no per-reaction overrides"
        ncell = params.shape[1]
        tp = max(timepoints, 1)
        out = {"y_final": np.zeros((ncell, self.neq)), "phys_final": np.zeros((ncell, 8)),
               "flag": np.zeros(ncell, np.int32), "stats": np.zeros((ncell, self._nstat), np.int64),
               "dissipation_time": np.zeros(ncell)}
        if want_physics:
            out["physics"] = np.zeros((ncell, timepoints + 1, 8))
        if want_chem:
            out["abund"] = np.zeros((ncell, timepoints + 1, self.nspec))
        if want_rates:
            out["rates"] = np.zeros((ncell, timepoints + 1, self.nreac))
        for c in range(ncell):
            r = self.orc.run_model(kind, np.ascontiguousarray(params[:, c]), y0=None if y0 is None else y0[c],
                                   timepoints=tp if timepoints else 500, rates=want_rates,
                                   pp_grid=None if pp_grid is None else pp_grid[c], pp_coldens=pp_coldens)
            out["y_final"][c], out["phys_final"][c], out["flag"][c] = r["y_final"], r["phys_final"], r["flag"]
            out["stats"][c, self._iint] = r["stats"]["nintervals"]
            out["dissipation_time"][c] = r["dissipation_time"]
            n = len(r["physics"])
            if want_physics:
                out["physics"][c, :n] = r["physics"]
            if want_chem:
                out["abund"][c, :n] = r["abund"]
            if want_rates:
                out["rates"][c, :n] = r["rates"]
        return out


def test_model_disk_mode_host_logic(oracle, net, tmp_path, monkeypatch):
    """The sequence of tests/test_gpu_parity.py::test_disk_mode_files with the oracle behind the library
    interface: outputFile / abundSaveFile written after the run, abundLoadFile seeding the next one."""
    from uclchem_b200 import model
    monkeypatch.setattr(model, "get_library", lambda *a, **k: _OracleBackedLibrary(oracle, net))
    pd_ = {"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 1e3}
    full, save = tmp_path / "full.dat", tmp_path / "final.dat"
    res = model.cloud(param_dict={**pd_, "outputFile": str(full), "abundSaveFile": str(save)}, out_species=["CO"])
    assert res[0] == 0 and len(res) == 2 and res[1] > 0
    phys, chem, _, start, flag = model.cloud(param_dict=pd_, return_array=True)
    names, data = datio.read_output_file(full)
    assert flag == 0 and names[8:] == net.names and data.shape == (phys.shape[0], 8 + net.nspec)
    assert np.allclose(data[:, 8:], chem[:, 0, :], rtol=1e-5, atol=0) and np.allclose(data[:, 0], phys[:, 0, 0], rtol=1e-3)
    assert phys[-1, 0, 0] == pytest.approx(1e3)
    final = datio.read_abundances(save, net.nspec)
    assert np.allclose(final, start, rtol=1e-5, atol=0)
    res2 = model.cloud(param_dict={**pd_, "finalTime": 1e2, "abundLoadFile": str(save), "outputFile": str(full)})
    _, data2 = datio.read_output_file(full)
    keep = np.array([n not in ("BULK", "SURFACE", "E-") for n in net.names])
    assert res2 == [0] and np.allclose(data2[0, 8:][keep], final[keep], rtol=1e-4, atol=1e-29)
    # columnFile (format 8030, every writeStep-th output after the initial state) and rateFile (format 8021)
    col, rat = tmp_path / "column.dat", tmp_path / "rates.dat"
    res3 = model.cloud(param_dict={**pd_, "columnFile": str(col), "rateFile": str(rat), "writeStep": 2}, out_species=["OH", "CO"])
    assert res3[0] == 0
    lines = col.read_text().splitlines()
    assert lines[0] == "Time,Density,gasTemp,dustTemp,av,radfield,zeta," + "OH".ljust(len(max(net.names, key=len))) + "," + "CO".ljust(len(max(net.names, key=len)))
    nrows = phys.shape[0]
    assert len(lines) - 1 == (nrows - 1) // 2
    first = [float(datio._fix_exp(v)) for v in lines[1].split(",")]
    assert len(first) == 9 and first[0] == pytest.approx(phys[2, 0, 0], rel=1e-3)
    assert first[8] == pytest.approx(chem[2, 0, net.names.index("CO")], rel=1e-5)
    rl = rat.read_text().splitlines()
    assert len(rl) == nrows and len(rl[-1].split(",")) == 8 + net.nreac and "E-0" in rl[-1]
    with pytest.raises(NotImplementedError):
        model.cloud(param_dict={**pd_, "fluxFile": "f.dat"})
    with pytest.raises(ValueError, match="out_species"):
        model.cloud(param_dict={**pd_, "columnFile": str(col)})
    with pytest.raises(RuntimeError, match="Offending keys"):
        model.cloud(param_dict={**pd_, "outputFile": str(full)}, return_array=True)


def test_grid_trajectories_host_logic(oracle, net, monkeypatch):
    """`*_grid(return_array=True)`: the reference's in-memory layout [time, point, field], one point per cell."""
    from uclchem_b200 import model
    monkeypatch.setattr(model, "get_library", lambda *a, **k: _OracleBackedLibrary(oracle, net))
    g = model.cloud_grid({"initialDens": [1e3, 1e4, 1e5], "finalTime": [1e2, 1e3, 1e2]}, out_species=["CO"],
                         return_array=True, timepoints=40)
    assert g["physics_array"].shape == (41, 3, 8) and g["chemical_abun_array"].shape == (41, 3, net.nspec)
    assert (g["flag"] == 0).all() and g["out_species"].shape == (3, 1)
    for c in range(3):
        n = int(g["nrows"][c])
        t = g["physics_array"][:n, c, 0]
        assert t[0] == 0.0 and (np.diff(t) > 0).all() and t[-1] == pytest.approx([1e2, 1e3, 1e2][c])
        assert (g["physics_array"][n:, c, 0] == 0.0).all()
        assert np.allclose(g["chemical_abun_array"][n - 1, c], g["abundances"][c], rtol=1e-12)
        assert g["physics_array"][0, c, 1] == pytest.approx([1e3, 1e4, 1e5][c])
    assert g["nrows"][1] > g["nrows"][0]
    with pytest.raises(RuntimeError, match="Offending keys"):
        model.cloud_grid({"initialDens": [1e3, 1e4], "outputFile": "x.dat"})


def test_cshock_return_conventions_host_logic(oracle, net, monkeypatch):
    """cshock's return tuples follow the reference (model.py:606-643): disk mode [flag, dissipation_time, *abunds],
    in-memory (physics, chem, rates, dissipation_time, abundanceStart, flag); on failure the dissipation time is
    None.  An unknown key is PARAMETER_READ_ERROR (-1) for a single model, not an exception."""
    from uclchem_b200 import model
    monkeypatch.setattr(model, "get_library", lambda *a, **k: _OracleBackedLibrary(oracle, net))
    sc = np.load(Path(__file__).parent / "golden" / "shockstart.npy")
    pd_ = {"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 2.0}
    phys, chem, rates, tdiss, start, flag = model.cshock(20.0, param_dict=pd_, return_array=True, starting_chemistry=sc)
    assert flag == 0 and rates is None and isinstance(tdiss, float) and 1e2 < tdiss < 1e5      # years
    assert start.shape == (net.nspec,) and np.array_equal(start, chem[-1, 0]) and phys.shape[1:] == (1, 8)
    res = model.cshock(20.0, param_dict=pd_, out_species=["CO"])
    assert res[0] == 0 and res[1] == pytest.approx(tdiss) and len(res) == 3
    # a failing model: too few time points in memory mode -> flag -6, dissipation time None (model.py:606-607)
    out = model.cshock(20.0, param_dict={**pd_, "finalTime": 40.0}, return_array=True, starting_chemistry=sc, timepoints=1)
    assert out[-1] == -6 and out[3] is None
    # disk mode has no row limit: the same model with the same tiny buffer still writes every row
    assert model.cloud(param_dict={"initialDens": 1e4, "finalTime": 1e2}, timepoints=2)[0] == 0
    assert model.cloud(param_dict={"initialDensity": 1e4}) == [-1]
    assert model.cshock(20.0, param_dict={"nonsense": 1.0}, return_array=True)[-1] == -1
    with pytest.raises(KeyError):
        model.cloud_grid({"initialDensity": [1e3, 1e4]})


def test_collapse_model_host_logic_and_oracle_physics(oracle, net, monkeypatch):
    """uclchem.model.collapse mirror (model.py:319-426) through the oracle-backed library double, and the oracle's
    collapse physics (oracle/orc_collapse.c) against an independent numpy evaluation of the filament fits of
    collapse.f90:178-266: parcel density rho(rout, t) and Av = baseAv + N(rin..rout) / 1.6e21."""
    from uclchem_b200 import model
    monkeypatch.setattr(model, "get_library", lambda *a, **k: _OracleBackedLibrary(oracle, net))
    with pytest.raises(ValueError, match="collapse must be one of"):
        model.collapse("spherical", None)
    with pytest.raises(NotImplementedError):
        model.collapse("filament", "physics.dat")
    pd_ = {"rout": 0.2, "baseAv": 1.0, "finalTime": 2.0e3, "initialTemp": 10.0}
    phys, chem, rates, start, flag = model.collapse("filament", None, param_dict=pd_, return_array=True)
    assert flag == 0 and rates is None and phys.shape[1:] == (1, 8) and chem.shape[2] == net.nspec
    t = phys[:, 0, 0]
    assert t[0] == 0.0 and t[-1] >= 2.0e3 and np.allclose(t[2:] / t[1:-1], 10.0)          # cadence collapse.f90:63-75 below 1e3 yr
    f32 = lambda x: float(np.float32(x))
    mh, pi, pc, spy = 1.67262164e-24, f32(3.141592654), 3.086e18, 3.16e7
    unitt = (2 * pi * 6.67e-8 * 2.2e4 * mh) ** -0.5 / spy
    unitr = np.sqrt(1.38e-16 * 10 / 2 / mh) * (2 * pi * 6.67e-8 * 2.2e4 * mh) ** -0.5 / pc
    def profile(r, tt):
        tn = tt / unitt
        rho0 = 10 ** (f32(3.54) * (f32(5.47) - tn) ** f32(-0.15) - f32(2.73))
        r0 = 10 ** (f32(-1.34) * (f32(5.47) - tn) ** f32(-0.15) + f32(1.47))
        a = 2.0 - 0.5 * (tn / f32(5.47)) ** 9
        return 2.2e4 * rho0 / (1 + (r / unitr / r0) ** 2) ** a
    for row in (1, 5, len(t) - 1):
        assert phys[row, 0, 1] == pytest.approx(profile(0.2, t[row]), rel=1e-12)         # velocity term: dt = 0 (reference quirk)
        r = np.linspace(0.0, 0.2, 10001)
        rho = profile(r, t[row])
        coldens = np.sum(0.5 * (rho[1:] + rho[:-1]) * (0.2 / 10000) * pc)
        assert phys[row, 0, 4] == pytest.approx(1.0 + coldens / 1.6e21, rel=1e-10)
    # Bonnor-Ebert modes run to 0.97 of the fit's time span whatever finalTime says (collapse.f90:36-41);
    # an unknown mode is a physics initialisation error (flag -2), never an exception
    from uclchem_b200.params import params_from_dict
    assert oracle.run_model(3, params_from_dict({"collapse_mode": 7})[:, 0])["flag"] == -2


def test_postprocess_host_logic_and_oracle_physics(oracle, net, monkeypatch):
    """uclchem.model.postprocess mirror (model.py:748-880) through the oracle-backed library double: one output row
    per history point, physics rows taken from the history exactly as postprocess.f90:112-129 does (the row written
    after interval k carries the values of history point k), Av from the supplied N_H (5.348e-22 N_H)."""
    from uclchem_b200 import model
    monkeypatch.setattr(model, "get_library", lambda *a, **k: _OracleBackedLibrary(oracle, net))
    n, spy = 6, 3.16e7
    t = np.linspace(0.0, 50.0, n) * spy
    dens, tg, td = np.linspace(1e4, 2e4, n), np.linspace(10, 20, n), np.linspace(10, 15, n)
    kw = dict(time_array=t, density_array=dens, gas_temperature_array=tg, dust_temperature_array=td,
              zeta_array=np.full(n, 2.0), radfield_array=np.full(n, 1.5))
    phys, chem, rates, start, flag = model.postprocess(param_dict={"initialDens": 1e4}, return_array=True, **kw)
    assert flag == 0 and phys.shape == (n + 1, 1, 8) and chem.shape == (n + 1, 1, net.nspec)
    assert np.allclose(phys[1:, 0, 0], t / spy + 1.0)                    # targets: history time + 1 yr (postprocess.f90:102)
    assert np.allclose(phys[1:, 0, 1], dens) and np.allclose(phys[1:, 0, 2], tg) and np.allclose(phys[1:, 0, 3], td)
    assert np.allclose(phys[1:, 0, 5], 1.5) and np.allclose(phys[1:, 0, 6], 2.0)
    nh = np.linspace(1e21, 2e21, n)
    phys2, _, _, _, flag2 = model.postprocess(param_dict={"initialDens": 1e4}, return_array=True, coldens_H_array=nh,
                                             coldens_H2_array=0.4 * nh, coldens_CO_array=1e-5 * nh, coldens_C_array=1e-6 * nh, **kw)
    assert flag2 == 0 and np.allclose(phys2[1:, 0, 4], float(np.float32(5.348e-22)) * nh)
    with pytest.raises(ValueError, match="together"):
        model.postprocess(return_array=True, coldens_H_array=nh, **kw)
    with pytest.raises(AssertionError, match="same length"):
        model.postprocess(return_array=True, **{**kw, "density_array": dens[:-1]})

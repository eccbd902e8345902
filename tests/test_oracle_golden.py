"""Pin the oracle (CPU restatement of the reference algorithm) to the reference's own
known answers: generated odes.f90 outputs, the example trajectories and a notebook result."""
import numpy as np
import pytest
from conftest import GOLDEN, STATIC, max_dex

from uclchem_b200.params import params_from_dict


def test_getydot_bit_exact_against_reference_odes(oracle, net):
    g = np.load(GOLDEN / "getydot_cases.npz")
    branches = set()
    for i in range(6):
        yd, sg = oracle.getydot(g[f"rate_{i}"], g[f"y_{i}"], float(g[f"blr_{i}"]), float(g[f"cov_{i}"]),
                                float(g[f"safe_mantle_{i}"]), float(g[f"safe_bulk_{i}"]), float(g[f"dens_{i}"]))
        assert np.array_equal(yd[: net.nspec], g[f"ydot_{i}"][: net.nspec])
        branches.add(sg < 0)
    assert branches == {True, False}  # both branches of the three-phase transfer are covered


@pytest.fixture(scope="module")
def static_run(oracle):
    return oracle.run_model(0, params_from_dict(STATIC)[:, 0])


def test_static_cloud_matches_golden_trajectory(static_run):
    """G1: every stored time of static-full.dat up to 1 Myr, to the file's 6-digit precision."""
    gold = np.load(GOLDEN / "static_full.npz")
    r = static_run
    assert r["flag"] == 0
    n = r["abund"].shape[0]
    assert n == 47 and r["stats"]["nintervals"] == 46
    np.testing.assert_allclose(r["physics"][:, 0], gold["physics"][:n, 0], rtol=1e-3)
    for row in range(1, n):
        assert max_dex(r["abund"][row], gold["abund"][row]) < 1e-5, row


def test_notebook_known_answer(oracle, net):
    """G5: notebooks/1_first_model.ipynb cell 3, full double precision."""
    p = params_from_dict({"endAtFinalDensity": False, "freefall": False, "initialDens": 1e4, "initialTemp": 10.0,
                          "finalTime": 1.0e6, "rout": 0.1, "baseAv": 1.0})[:, 0]
    r = oracle.run_model(0, p)
    so, co = r["y_final"][net.names.index("SO")], r["y_final"][net.names.index("CO")]
    assert abs(so / 4.3502289723887733e-10 - 1) < 1e-6
    assert abs(co / 2.2662477663676093e-05 - 1) < 1e-6


def test_freefall_phase1_matches_golden(oracle):
    """G2: free-fall collapse 1e2 -> 1e5 cm-3 (density is the 336th ODE)."""
    gold = np.load(GOLDEN / "phase1_full.npz")
    p = params_from_dict({"endAtFinalDensity": True, "freefall": True, "initialDens": 1e2, "initialTemp": 10.0,
                          "finalDens": 1e5, "finalTime": 5.0e6})[:, 0]
    r = oracle.run_model(0, p)
    assert r["flag"] == 0 and r["abund"].shape[0] == gold["abund"].shape[0] == 90
    for row in range(1, 90):
        assert max_dex(r["abund"][row], gold["abund"][row]) < 0.01, row
    assert max_dex(r["y_final"][:335], np.load(GOLDEN / "startcollapse.npy")) < 0.01


def test_hot_core_phase2_matches_golden_prefix(oracle):
    """G3 (first 1e4 yr of 1 Myr, to keep the CPU suite short): hot_core(3, 300) from startcollapse."""
    gold = np.load(GOLDEN / "phase2_full.npz")
    sc = np.load(GOLDEN / "startcollapse.npy")
    p = params_from_dict({"endAtFinalDensity": False, "freefall": False, "initialDens": 1e5, "initialTemp": 10.0,
                          "finalDens": 1e5, "finalTime": 1.0e4, "freezeFactor": 0.0, "thermdesorb": True,
                          "temp_indx": 3, "max_temperature": 300.0})[:, 0]
    r = oracle.run_model(1, p, y0=np.append(sc, 1e5))
    assert r["flag"] == 0
    n = r["abund"].shape[0]
    assert n > 90
    np.testing.assert_allclose(r["physics"][1:n, 2], gold["physics"][1:n, 2], atol=6e-3)  # gasTemp, f8.2 format
    for row in range(1, n):
        assert max_dex(r["abund"][row], gold["abund"][row]) < 0.01, row


def test_ode_conservation(oracle, net):
    """reference tests/test_ode_conservation.py: elements are linear invariants of the RHS."""
    p = params_from_dict({"endAtFinalDensity": False, "freefall": True, "initialDens": 1e4, "initialTemp": 10.0,
                          "finalDens": 1e5, "finalTime": 1.0e3})[:, 0]
    r = oracle.run_model(0, p)
    ydot = oracle.get_odes(p, r["y_final"][:335])
    elements, counts, _ = net.element_matrix()
    for e in ("H", "N", "C", "O"):
        assert abs(counts[elements.index(e)] @ ydot[:335]) < 1e-15


def test_cshock_runs_and_conserves_elements(oracle, net):
    """No golden exists for the C-shock: smoke + invariants only (parity pinned via self-consistency)."""
    sc = np.load(GOLDEN / "shockstart.npy")
    p = params_from_dict({"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 20.0, "shock_vel": 30.0,
                          "timestep_factor": 0.01, "minimum_temperature": 0.0})[:, 0]
    r = oracle.run_model(2, p, y0=np.append(sc, 1e4))
    assert r["flag"] == 0 and r["dissipation_time"] > 0
    assert r["physics"][-1, 2] > 10.0  # the gas is being heated
    elements, counts, _ = net.element_matrix()
    for e in ("C", "O", "N"):
        t0, t1 = counts[elements.index(e)] @ sc, counts[elements.index(e)] @ r["y_final"][:335]
        assert abs(t1 / t0 - 1) < 1e-3  # sputtering's many-one store (sputtering.f90:104) is not conservative

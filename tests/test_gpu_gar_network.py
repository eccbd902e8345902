"""Third network (default + grain-assisted recombination, `libuclgpu_gar.so`) on the GPU against the oracle: the GAR
reaction type (rates.f90:316-332, density-weighted fluxes) through rate coefficients, F and whole models."""
import numpy as np
import pytest
from conftest import GOLDEN, ROOT, max_dex

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net3():
    from uclchem_b200.network import Network
    return Network.from_json(ROOT / "uclchem_b200" / "networks" / "gar.json")


@pytest.fixture(scope="module")
def lib3():
    from uclchem_b200._capi import Library
    L = Library("gar")
    L.init([0])
    return L


@pytest.fixture(scope="module")
def oracle3(net3):
    from oracle.oracle import Oracle
    return Oracle(net3)


def test_gar_rates_and_rhs_match_the_oracle(lib3, oracle3, net3):
    from uclchem_b200.params import params_from_dict
    gold = np.load(GOLDEN / "static_full.npz")
    lo, hi = net3.type_ranges["GAR"]
    for pd_ in ({"initialDens": 1e3, "initialTemp": 30.0, "radfield": 10.0, "baseAv": 0.5},
                {"initialDens": 1e5, "initialTemp": 10.0}):
        ys = np.array([np.append(np.maximum(gold["abund"][r], 1e-30), pd_["initialDens"]) for r in (3, 25, 46)])
        p1 = params_from_dict(pd_)
        pp = np.repeat(p1, len(ys), axis=1)
        rates, rhs = lib3.get_rates(pp, ys), lib3.probe_rhs(pp, ys)
        for k in range(len(ys)):
            ref = oracle3.get_rates(p1[:, 0], ys[k, :335])
            assert np.array_equal(rates[k] == 0, ref == 0) and (ref[lo:hi + 1] > 0).all()
            m = ref != 0
            assert np.abs(rates[k][m] / ref[m] - 1).max() < 1e-12
            f = oracle3.probe_rhs(p1[:, 0], ys[k, :335])
            assert np.abs(rhs[k] - f).max() <= 1e-9 * np.abs(f).max()


def test_models_with_gar_match_the_oracle_and_differ_from_the_default_network(lib3, oracle3, net3, lib):
    """Translucent, irradiated gas is where grain-assisted recombination competes with radiative recombination:
    the ionisation balance must follow the oracle of THIS network and move away from the default network's."""
    from uclchem_b200.params import params_from_dict
    p = params_from_dict({"initialDens": [3e2, 1e3, 1e4], "initialTemp": [50.0, 30.0, 10.0], "radfield": [10.0, 3.0, 1.0],
                          "baseAv": [0.3, 0.5, 2.0], "finalTime": 1e5})
    out = lib3.run_grid(0, p, step_budget=300000)
    ref, _, flag, st = oracle3.run_grid(0, p, nthreads=3)
    base = lib.run_grid(0, p, step_budget=300000)
    assert (out["flag"] == 0).all() and (flag == 0).all() and (base["flag"] == 0).all()
    for c in range(3):
        assert max_dex(out["y_final"][c, :335], ref[c, :335]) <= 0.01, c
        assert 0.5 * st[c, 0] < out["stats"][c, 0] < 1.5 * st[c, 0]
    assert max_dex(out["y_final"][0, :335], base["y_final"][0, :335]) > 0.02      # GAR matters in the diffuse cell

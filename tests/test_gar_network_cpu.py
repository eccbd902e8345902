"""Third network: default + grain-assisted recombination (GAR, rates.f90:316-332), 335 species / 3209 reactions,
produced by the reference's own MakeRates from its `gar_settings.yaml` (tools/make_gar_network.py).  It puts the GAR
reaction type -- which the default network does not contain -- through the oracle, the generator and the device
library: the RHS is pinned on the reference-generated odes.f90 of THAT network (GAR reactions carry an extra
factor of the gas density, reaction.py:792-794), the GAR rate coefficient on an independent evaluation of the
Weingartner & Draine formula, and the generated Jacobian / LU on dense algebra."""
import numpy as np
import pytest
from conftest import GOLDEN, ROOT

from uclchem_b200 import product_form
from uclchem_b200.makerates_cuda import Generated
from uclchem_b200.network import Network
from uclchem_b200.params import params_from_dict
from uclchem_b200.table_emulator import TableEngine


@pytest.fixture(scope="module")
def net3():
    return Network.from_json(ROOT / "uclchem_b200" / "networks" / "gar.json")


@pytest.fixture(scope="module")
def gen3(net3):
    return Generated(net3)


@pytest.fixture(scope="module")
def cases():
    return np.load(GOLDEN / "getydot_cases_gar.npz")


def test_sizes(net3, gen3):
    lo, hi = net3.type_ranges["GAR"]
    assert (net3.nspec, net3.nreac) == (335, 3209) and hi - lo + 1 == 6 and net3.gar_params.shape == (6, 7)
    assert {net3.names[net3.re[r, 0]] for r in range(lo, hi + 1)} >= {"H+", "HE+", "C+", "MG+"}


def test_oracle_rhs_on_reference_generated_odes(net3, cases):
    from oracle.oracle import Oracle
    orc = Oracle(net3)
    for i in range(6):
        ref = cases[f"ydot_{i}"]
        got, _ = orc.getydot(cases[f"rate_{i}"], cases[f"y_{i}"], float(cases[f"blr_{i}"]), float(cases[f"cov_{i}"]),
                             float(cases[f"safe_mantle_{i}"]), float(cases[f"safe_bulk_{i}"]), float(cases[f"dens_{i}"]))
        assert np.abs(got[:335] - ref[:335]).max() <= 1e-14 * np.abs(ref).max()


def test_generated_tables_on_reference_generated_odes_and_dense_algebra(net3, gen3, cases):
    sym = gen3.sym
    eng = TableEngine(sym)
    gas = np.array([k for k, n in enumerate(net3.names) if n[0] not in "#@" and n not in ("BULK", "SURFACE")])
    for i in (0, 1):
        y, rate, ref = cases[f"y_{i}"].copy(), cases[f"rate_{i}"], cases[f"ydot_{i}"]
        y[sym.iB], y[sym.iS] = y[net3.bulk_list].sum(), y[net3.surface_list].sum()
        got, _ = eng.rhs(y, rate)
        assert np.abs(got[gas] - ref[gas]).max() <= 1e-12 * np.abs(ref[gas]).max()
    y, rate = cases["y_0"].copy(), cases["rate_0"]
    y[sym.iB], y[sym.iS] = y[net3.bulk_list].sum(), y[net3.surface_list].sum()
    neq, gamma = sym.neq, 1e3
    f = lambda yy: eng.rhs(yy, rate)[0]
    J = np.zeros((neq, neq))
    for j in range(neq):
        h = max(abs(y[j]) * 1e-6, 1e-30)
        yp, ym = y.copy(), y.copy()
        yp[j] += h
        ym[j] -= h
        J[:, j] = (f(yp) - f(ym)) / (2 * h)
    b = np.random.default_rng(2).standard_normal(neq) * np.abs(y)
    x_ref = np.linalg.solve(np.eye(neq) - gamma * J, b)
    ba = np.zeros(sym.naug)
    ba[:neq] = b
    ba[sym.iB] = b[sym.iB] - b[net3.bulk_list].sum()
    ba[sym.iS] = b[sym.iS] - b[net3.surface_list].sum()
    val = eng.factor(eng.assemble(y, rate, gamma))
    assert np.abs(eng.solve(val, ba)[:neq] - x_ref).max() <= 1e-7 * np.abs(x_ref).max()
    xp = product_form.solve(gen3.pf, sym, product_form.invert(gen3.pf, sym, val), ba)
    assert np.abs(xp[:neq] - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


def test_gar_rate_coefficients_follow_weingartner_draine(net3):
    """rates.f90:316-332: k = 0.6 alpha c0 / (1 + c1 phi^c2 (1 + c3 T^c4 phi^(-c5 - c6 ln T))) with
    phi = G exp(-2.5 Av) sqrt(T) / (n n_e) clamped to [1e2, 1e6] (single-precision literals kept)."""
    from oracle.oracle import Oracle
    orc = Oracle(net3)
    lo, hi = net3.type_ranges["GAR"]
    gold = np.load(GOLDEN / "static_full.npz")
    f32 = lambda x: float(np.float32(x))
    for pd_ in ({"initialDens": 1e3, "initialTemp": 30.0, "radfield": 10.0, "baseAv": 0.5},
                {"initialDens": 1e5, "initialTemp": 10.0, "radfield": 1.0, "baseAv": 2.0}):
        p = params_from_dict(pd_)[:, 0]
        y = np.maximum(gold["abund"][12], 1e-30)
        rate = orc.get_rates(p, y)
        av = pd_["baseAv"] + float(np.float32(0.05)) * 3.086e18 * pd_["initialDens"] / 1.6e21
        phi = pd_["radfield"] * np.exp(-2.5 * av) * np.sqrt(pd_["initialTemp"]) / (pd_["initialDens"] * y[net3.species_idx["nelec"]])
        phi = min(max(phi, f32(1e2)), f32(1e6))
        T = pd_["initialTemp"]
        g = net3.gar_params
        ref = f32(0.6) * net3.alpha[lo:hi + 1] * g[:, 0] / (1.0 + g[:, 1] * phi ** g[:, 2] *
                                                         (1.0 + g[:, 3] * T ** g[:, 4] * phi ** (-g[:, 5] - g[:, 6] * np.log(T))))
        assert (rate[lo:hi + 1] > 0).all()
        np.testing.assert_allclose(rate[lo:hi + 1], ref, rtol=1e-12)


def test_device_library_for_the_gar_network_builds_and_loads(net3):
    from uclchem_b200 import build
    from uclchem_b200._capi import Library
    build.compile("gar")
    L = Library("gar")
    assert (L.nspec, L.nreac) == (335, 3209)

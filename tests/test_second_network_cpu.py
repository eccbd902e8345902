"""Second network (BASELINE config 5: `add_crp_photo_to_grain: True`, 335 species / 3453 reactions,
produced by the reference's own MakeRates, tools/make_second_network.py).  The oracle and the MakeRates
CUDA back-end are network-generic: both are pinned here on RHS known answers evaluated with the
reference-generated odes.f90 of THAT network, and the generated Jacobian / symbolic LU / product-form
programs are checked against dense linear algebra.  (The device library for this network is not built
yet: its Newton matrix needs 136 KB and the CTA layout has room for 125 KB -- DESIGN.md.)"""
import numpy as np
import pytest
from conftest import GOLDEN, ROOT

from uclchem_b200 import product_form
from uclchem_b200.makerates_cuda import Generated
from uclchem_b200.network import Network
from uclchem_b200.table_emulator import TableEngine


@pytest.fixture(scope="module")
def net2():
    return Network.from_json(ROOT / "uclchem_b200" / "networks" / "crp_photo.json")


@pytest.fixture(scope="module")
def gen2(net2):
    return Generated(net2)


@pytest.fixture(scope="module")
def cases():
    return np.load(GOLDEN / "getydot_cases_crp_photo.npz")


def test_sizes(net2, gen2):
    assert (net2.nspec, net2.nreac) == (335, 3453)
    assert net2.type_ranges["CRPHOT"][1] - net2.type_ranges["CRPHOT"][0] + 1 == 240   # 120 gas + 120 on grains
    s = gen2.sym.stats
    assert gen2.sym.naug == 338 and s["n0"] + s["m"] == 338
    assert len(gen2.flux_order) == net2.nreac - 2 and len(gen2.deferred) == 2


def test_oracle_rhs_on_reference_generated_odes(net2, cases):
    """The oracle's table-driven GETYDOT equals the reference's generated odes.f90 of this network."""
    from oracle.oracle import Oracle
    orc = Oracle(net2)
    for i in range(6):
        ref = cases[f"ydot_{i}"]
        got, _ = orc.getydot(cases[f"rate_{i}"], cases[f"y_{i}"], float(cases[f"blr_{i}"]), float(cases[f"cov_{i}"]),
                             float(cases[f"safe_mantle_{i}"]), float(cases[f"safe_bulk_{i}"]), float(cases[f"dens_{i}"]))
        assert np.abs(got[:335] - ref[:335]).max() <= 1e-14 * np.abs(ref).max()


def test_generated_gather_on_reference_generated_odes(net2, gen2, cases):
    """Flux table + gather program + deferred photo reactions reproduce the reaction sums of odes.f90
    (gas-phase rows: untouched by the three-phase transfer, so they compare directly)."""
    sym = gen2.sym
    for i in (0, 1):
        y, rate, ref = cases[f"y_{i}"], cases[f"rate_{i}"], cases[f"ydot_{i}"]
        ye = np.empty(sym.neq + 4)
        ye[: sym.neq] = y
        sm, sb = float(cases[f"safe_mantle_{i}"]), float(cases[f"safe_bulk_{i}"])
        blr = float(cases[f"blr_{i}"])
        swap = float(np.sum(rate[sym.swap_reacs] * y[net2.re[sym.swap_reacs, 0]] * blr))
        ye[sym.neq:] = [1.0, blr, 1.0 / sm, swap / sm]
        flux = rate * np.prod(ye[sym.flux_f], axis=1)
        out = np.zeros(sym.neq)
        gen2.gather.run(lambda t: -flux[t & 0x7FFF] if (t >> 15) & 1 else flux[t & 0x7FFF],
                        lambda tg, s: out.__setitem__(tg, s))
        for r in gen2.deferred:
            for k, sg in gen2.deferred_rows[r]:
                out[k] += sg * flux[r]
        gas = np.array([k for k, n in enumerate(net2.names) if n[0] not in "#@" and n not in ("BULK", "SURFACE")])
        assert np.abs(out[gas] - ref[gas]).max() <= 1e-13 * np.abs(ref[gas]).max()


@pytest.mark.parametrize("case,gamma", [(0, 1e3), (1, 1e3)])
def test_jacobian_lu_and_product_form(net2, gen2, cases, case, gamma):
    sym = gen2.sym
    eng = TableEngine(sym)
    y, rate = cases[f"y_{case}"].copy(), cases[f"rate_{case}"]
    y[sym.iB] = y[net2.bulk_list].sum()
    y[sym.iS] = y[net2.surface_list].sum()
    neq = sym.neq
    f = lambda yy: eng.rhs(yy, rate)[0]
    J = np.zeros((neq, neq))
    for j in range(neq):
        h = max(abs(y[j]) * 1e-6, 1e-30)
        yp, ym = y.copy(), y.copy()
        yp[j] += h
        ym[j] -= h
        J[:, j] = (f(yp) - f(ym)) / (2 * h)
    rng = np.random.default_rng(2)
    b = rng.standard_normal(neq) * np.abs(y)
    x_ref = np.linalg.solve(np.eye(neq) - gamma * J, b)
    ba = np.zeros(sym.naug)
    ba[:neq] = b
    ba[sym.iB] = b[sym.iB] - b[net2.bulk_list].sum()
    ba[sym.iS] = b[sym.iS] - b[net2.surface_list].sum()
    val = eng.factor(eng.assemble(y, rate, gamma))
    x = eng.solve(val, ba)
    assert np.abs(x[:neq] - x_ref).max() <= 1e-7 * np.abs(x_ref).max()
    xp = product_form.solve(gen2.pf, sym, product_form.invert(gen2.pf, sym, val), ba)
    assert np.abs(xp[:neq] - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


def test_oracle_runs_the_reference_photo_on_grain_static_model(net2):
    """First model of tests/test_photo_on_grain.py:101-120 of the reference on this network: the static
    cloud with the reference's own tolerances for this (stiffer) network, shortened from 5 Myr to 1 Myr
    to keep the CPU suite short (the free-fall and hot-core stages of that test take minutes per model
    on this network).  The reference asserts `return_code == 0` and nothing else; the budgets of the
    elements every reaction conserves must also stay put."""
    from oracle.oracle import Oracle
    from uclchem_b200.params import params_from_dict
    orc = Oracle(net2)
    orc.set_deadline(120.0)
    r = orc.run_model(0, params_from_dict({"endAtFinalDensity": False, "freefall": False, "initialDens": 1e4,
                                           "initialTemp": 10.0, "finalDens": 1e5, "finalTime": 1.0e6,
                                           "abstol_min": 1e-15, "reltol": 1e-5})[:, 0])
    orc.set_deadline(0.0)
    assert r["flag"] == 0 and r["physics"][-1, 0] == pytest.approx(1.0e6, rel=1e-6)
    for e in ("C", "O", "N"):
        w = np.array([_count(n, e) for n in net2.names[:333]], float)
        a0, a1 = float(w @ r["abund"][0, :333]), float(w @ r["abund"][-1, :333])
        assert a0 > 0 and abs(a1 - a0) <= 1e-6 * a0, (e, a0, a1)


def _count(name, element):
    """Atoms of `element` (one- or two-letter symbol) in a species name like '#CH3OH', 'HCO+', '@SIC2'."""
    import re
    s = name.lstrip("#@").rstrip("+-")
    tot = 0
    for sym, num in re.findall(r"(CL|MG|SI|HE|[A-Z])(\d*)", s):
        if sym == element:
            tot += int(num) if num else 1
    return tot


def test_device_library_for_the_second_network_builds_and_loads(net2):
    """`libuclgpu_crp_photo.so` (compact shared-memory layout, dense threshold 0.95: uclchem_b200/build.py)
    compiles for sm_100a, loads without a GPU and reports this network; no compute call here."""
    import shutil
    from uclchem_b200 import build
    from uclchem_b200._capi import EXPORTED, Library
    if not (shutil.which("nvcc") or __import__("pathlib").Path("/usr/local/cuda/bin/nvcc").exists()):
        pytest.skip("nvcc not available")
    so = build.compile("crp_photo")
    L = Library("crp_photo")
    assert so.exists() and (L.nspec, L.nreac, L.naug) == (net2.nspec, net2.nreac, 338) and L.tag == "crp_photo"
    assert all(hasattr(L.lib, sym) for sym in EXPORTED)
    assert L.species == net2.names

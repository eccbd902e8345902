"""GPU parity tests proper: every call goes through the C ABI (ctypes -> libuclgpu_default.so)
and is compared with the oracle on the same inputs, or with the reference's golden tables."""
import numpy as np
import pytest
from conftest import GOLDEN, STATIC, max_dex

from uclchem_b200 import model
from uclchem_b200.params import PARAM_INDEX, params_from_dict

pytestmark = pytest.mark.gpu

DEX_TOL = 0.01     # north-star: <= 0.01 dex for every species above 1e-15
CONS_TOL = 1e-10   # north-star: element conservation to 1e-10 relative


def _states(n=4):
    gold = np.load(GOLDEN / "static_full.npz")
    rows = [0, 3, 12, 25, 46][:n]
    return np.array([np.append(np.maximum(gold["abund"][r], 1e-30), 1e4) for r in rows])


def test_rate_coefficients_match_oracle(lib, oracle):
    """calculateReactionRates (rates.f90:21-343) per cell: <= 1e-12 relative, same zero pattern."""
    ys = _states()
    for pd_ in (STATIC, dict(STATIC, initialTemp=30.0, zeta=10.0, radfield=3.0, baseAv=1.0),
                dict(STATIC, initialTemp=200.0, initialDens=1e6, zeta=100.0, radfield=10.0, baseAv=0.5),
                dict(STATIC, initialTemp=60.0, cosmicRayAttenuation=True, ionModel="H", improvedH2CRPDissociation=True)):
        p1 = params_from_dict(pd_)
        got = lib.get_rates(np.repeat(p1, len(ys), axis=1), ys)
        for k in range(len(ys)):
            ref = oracle.get_rates(p1[:, 0], ys[k, :335])
            assert np.array_equal(got[k] == 0, ref == 0)
            m = ref != 0
            assert np.abs(got[k][m] / ref[m] - 1).max() < 1e-12


def test_get_odes_matches_oracle(lib, oracle):
    """get_odes (wrap.f90:516-547) = 1e-7 s of integration followed by F."""
    ys = _states(3)[1:]
    p1 = params_from_dict(STATIC)
    got = lib.get_odes(np.repeat(p1, len(ys), axis=1), ys)
    for k in range(len(ys)):
        ref = oracle.get_odes(p1[:, 0], ys[k, :335])
        assert np.abs(got[k] - ref).max() <= 1e-5 * np.abs(ref).max()


def test_rhs_elements_conserved(lib, net):
    """reference tests/test_ode_conservation.py on the device RHS."""
    ys = _states()
    p1 = params_from_dict(STATIC)
    yd = lib.probe_rhs(np.repeat(p1, len(ys), axis=1), ys)
    elements, counts, _ = net.element_matrix()
    for k in range(len(ys)):
        for e in ("H", "N", "C", "O"):
            assert abs(counts[elements.index(e)] @ yd[k, :335]) < 1e-15


def test_newton_solve_matches_dense(lib, net):
    """Analytic Jacobian + generated sparse LU on the device against numpy dense algebra."""
    from uclchem_b200 import symbolic
    from uclchem_b200.table_emulator import TableEngine
    sym = symbolic.build(net)
    eng = TableEngine(sym)
    ys = _states(3)
    p1 = params_from_dict(STATIC)
    pp = np.repeat(p1, len(ys), axis=1)
    rates = lib.get_rates(pp, ys)
    rng = np.random.default_rng(0)
    for gamma in (1e3, 1e9):
        b = rng.standard_normal((len(ys), 336)) * np.abs(ys)
        x = lib.probe_newton(pp, ys, gamma, b)
        for k in range(len(ys)):
            y = ys[k].copy()
            y[sym.iB], y[sym.iS] = y[net.bulk_list].sum(), y[net.surface_list].sum()
            val = eng.assemble(y, rates[k], gamma)
            A = eng.to_dense(val)
            Aold = np.zeros_like(A)
            Aold[np.ix_(sym.perm, sym.perm)] = A
            ba = np.zeros(sym.naug)
            ba[:336] = b[k]
            ba[sym.iB] = b[k][sym.iB] - b[k][net.bulk_list].sum()     # constraint rows: r - sum(members)
            ba[sym.iS] = b[k][sym.iS] - b[k][net.surface_list].sum()
            ref = np.linalg.solve(Aold, ba)
            w = 1.0 / (1e-8 * np.abs(y) + 1e-25)
            err = np.sqrt(np.mean(((x[k] - ref)[:336] * w) ** 2))
            assert err <= 1e-5 * np.sqrt(np.mean((ref[:336] * w) ** 2))  # cond(P) reaches 1e25 for the late-time state


@pytest.fixture(scope="module")
def static_gpu(lib):
    return lib.run_grid(0, params_from_dict(STATIC), timepoints=500, want_physics=True, want_chem=True)


def test_static_cloud_config1_against_golden(static_gpu):
    """BASELINE config 1: the reference's own 1 Myr static cloud, every stored time."""
    gold = np.load(GOLDEN / "static_full.npz")
    out = static_gpu
    assert out["flag"][0] == 0 and out["stats"][0][7] == 46
    np.testing.assert_allclose(out["physics"][0, :47, 0], gold["physics"][:47, 0], rtol=1e-3)
    worst = max(max_dex(out["abund"][0, r], gold["abund"][r]) for r in range(1, 47))
    assert worst < 1e-4, worst  # far inside the 0.01 dex bar: limited by the 6 digits of the file


def test_element_and_charge_conservation(static_gpu, net, oracle):
    """North-star: element (and charge) conservation to 1e-10 relative over the whole 1 Myr run.
    Elements that every reaction conserves must be invariant outright.  H and charge are NOT
    invariants of the reference RHS (protonated ions freeze out as their neutral, electrons
    "freeze" at the H rate: SURVEY.md Q10); for those the drift must equal the oracle's drift."""
    out = static_gpu
    elements, counts, charge = net.element_matrix()
    ls, lr, gs, gr = net.stoichiometry()
    ref = oracle.run_model(0, params_from_dict(STATIC)[:, 0])
    a0, a1, r1 = out["abund"][0, 0], out["abund"][0, 46], ref["y_final"][:335]
    exact = []
    for e, row in list(zip(elements, counts)) + [("charge", charge)]:
        imbalance = np.zeros(net.nreac)
        np.add.at(imbalance, gr, row[gs])
        np.add.at(imbalance, lr, -row[ls])
        scale = np.abs(row * a0).sum()
        if not imbalance.any():
            exact.append(e)
            assert abs(row @ a1 - row @ a0) < CONS_TOL * scale, e
        else:
            assert abs(row @ a1 - row @ r1) < CONS_TOL * scale, e
    assert {"HE", "C", "N", "O", "S", "SI", "MG", "CL"} <= set(exact) and "H" not in exact


def test_small_grid_against_oracle(lib, oracle, net):
    """A slice of BASELINE config 2 (density x temperature x zeta), 10^4 yr, vs the oracle."""
    dens, temp, zeta = np.meshgrid([1e3, 1e5, 1e7], [10.0, 55.0], [1.0, 100.0], indexing="ij")
    p = params_from_dict({"initialDens": dens.ravel(), "initialTemp": temp.ravel(), "zeta": zeta.ravel(),
                          "radfield": 1.0, "baseAv": 2.0, "rout": 0.05, "finalTime": 1e4})
    out = lib.run_grid(0, p)
    ref, _, rflag, _ = oracle.run_grid(0, p)
    assert (out["flag"] == 0).all() and (rflag == 0).all()
    for c in range(p.shape[1]):
        assert max_dex(out["y_final"][c, :335], ref[c, :335]) < DEX_TOL, c
    np.testing.assert_allclose(out["phys_final"][:, 4], 2.0 + 0.05 * 3.086e18 * dens.ravel() / 1.6e21, rtol=1e-12)


def test_freefall_and_hot_core_chain(lib, oracle):
    """BASELINE config 3 in miniature: free-fall stage 1, then hot_core stage 2 from its output."""
    p1 = params_from_dict({"endAtFinalDensity": True, "freefall": True, "initialDens": 1e2, "finalDens": 1e5,
                           "initialTemp": 10.0, "finalTime": 5e6})
    s1 = lib.run_grid(0, p1)
    r1 = oracle.run_model(0, p1[:, 0])
    assert s1["flag"][0] == 0 and s1["stats"][0][7] == r1["stats"]["nintervals"] == 89
    assert max_dex(s1["y_final"][0, :335], r1["y_final"][:335]) < DEX_TOL
    assert abs(s1["phys_final"][0, 1] / r1["phys_final"][1] - 1) < 1e-4
    y0 = r1["y_final"][None, :]
    p2 = params_from_dict({"initialDens": 1e5, "initialTemp": 10.0, "finalTime": 2e4, "freezeFactor": 0.0,
                           "temp_indx": [1, 3, 5], "max_temperature": [100.0, 300.0, 250.0]})
    s2 = lib.run_grid(1, p2, y0=np.repeat(y0, 3, axis=0))
    assert (s2["flag"] == 0).all()
    for c in range(3):
        r2 = oracle.run_model(1, p2[:, c], y0=y0[0])
        assert abs(s2["phys_final"][c, 2] - r2["phys_final"][2]) < 1e-9
        assert max_dex(s2["y_final"][c, :335], r2["y_final"][:335]) < DEX_TOL, c


def test_cshock_against_oracle(lib, oracle):
    """BASELINE config 4 in miniature (parity pinned only via the self-validated oracle)."""
    sc = np.load(GOLDEN / "shockstart.npy")
    p = params_from_dict({"initialDens": [1e4, 1e5], "initialTemp": 10.0, "finalTime": 30.0,
                          "shock_vel": [20.0, 35.0], "reltol": 1e-6, "abstol_min": 1e-20})
    y0 = np.repeat(np.append(sc, 1e4)[None, :], 2, axis=0)
    out = lib.run_grid(2, p, y0=y0)
    assert (out["flag"] == 0).all()
    for c in range(2):
        r = oracle.run_model(2, p[:, c], y0=y0[c])
        assert abs(out["dissipation_time"][c] / r["dissipation_time"] - 1) < 1e-12
        np.testing.assert_allclose(out["phys_final"][c, 1:3], r["phys_final"][1:3], rtol=1e-6)
        assert max_dex(out["y_final"][c, :335], r["y_final"][:335]) < DEX_TOL, c


def test_edge_cases(lib):
    # empty grid
    out = lib.run_grid(0, params_from_dict({"initialDens": np.zeros(0)}, ncell=0))
    assert out["flag"].shape == (0,)
    # physics init failure is a per-cell flag, not an exception (constants.f90:23)
    p = params_from_dict({"temp_indx": [3, 9], "max_temperature": 300.0, "finalTime": 1e-6})
    assert list(lib.run_grid(1, p)["flag"]) == [0, -2]
    # ragged costs in one launch: a cheap and an expensive cell, plus more cells than SMs
    p = params_from_dict({"initialDens": np.r_[1e7, np.full(160, 1e3)], "finalTime": np.r_[1e3, np.full(160, 1e-5)]})
    o = lib.run_grid(0, p)
    assert (o["flag"] == 0).all() and o["stats"][0, 0] > 10 * o["stats"][1:, 0].max()
    # too few time points -> NOT_ENOUGH_TIMEPOINTS_ERROR (io.f90:69-72)
    o = lib.run_grid(0, params_from_dict({"finalTime": 1.0}), timepoints=3, want_chem=True, want_physics=True)
    assert o["flag"][0] == -6


def test_model_api_mirrors_reference(lib, net):
    """uclchem.model.cloud call conventions (model.py:227-316)."""
    pd_ = {"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 1e3}
    res = model.cloud(param_dict=pd_, out_species=["SO", "CO"])
    assert res[0] == 0 and len(res) == 3 and 0 < res[1] < res[2]
    phys, chem, rates, start, flag = model.cloud(param_dict=pd_, return_array=True, return_rates=True)
    assert flag == 0 and phys.shape[1:] == (1, 8) and chem.shape[1:] == (1, 335) and rates.shape[2] == 3203
    assert phys[-1, 0, 0] == pytest.approx(1e3) and np.array_equal(start, chem[-1, 0])
    df_phys, df_chem, _, _, flag = model.cloud(param_dict=pd_, return_dataframe=True)
    assert list(df_phys.columns) == model.PHYSICAL_PARAMETERS and df_chem.columns[0] == "H"
    phys2, chem2, _, _, flag2 = model.hot_core(3, 300.0, param_dict={"initialDens": 1e5, "finalTime": 1e2,
                                               "freezeFactor": 0.0}, return_array=True, starting_chemistry=start)
    assert flag2 == 0 and np.allclose(chem2[0, 0], start)
    g = model.cloud_grid({"initialDens": [1e3, 1e4, 1e5], "finalTime": 1e2}, out_species=["CO"])
    assert g["abundances"].shape == (3, 335) and (g["flag"] == 0).all()


def test_disk_mode_files(lib, net, tmp_path):
    """The reference's disk mode (io.f90 formats, uclchem_b200/datio.py): outputFile / abundSaveFile are
    written after the run, abundLoadFile seeds the next one."""
    from uclchem_b200 import datio
    pd_ = {"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 1e3}
    full, save = tmp_path / "full.dat", tmp_path / "final.dat"
    res = model.cloud(param_dict={**pd_, "outputFile": str(full), "abundSaveFile": str(save)}, out_species=["CO"])
    assert res[0] == 0 and res[1] > 0
    phys, chem, _, start, flag = model.cloud(param_dict=pd_, return_array=True)
    names, data = datio.read_output_file(full)
    assert names[8:] == net.names and data.shape == (phys.shape[0], 8 + net.nspec)
    assert np.allclose(data[:, 8:], chem[:, 0, :], rtol=1e-5, atol=0) and np.allclose(data[:, 0], phys[:, 0, 0], rtol=1e-3)
    final = datio.read_abundances(save, net.nspec)
    assert np.allclose(final, start, rtol=1e-5, atol=0)
    res2 = model.cloud(param_dict={**pd_, "finalTime": 1e2, "abundLoadFile": str(save), "outputFile": str(full)})
    _, data2 = datio.read_output_file(full)
    # row 0 of the new run is the loaded state (totals like BULK / SURFACE / E- may be re-derived by the model)
    keep = np.array([n not in ("BULK", "SURFACE", "E-") for n in net.names])
    assert res2[0] == 0 and np.allclose(data2[0, 8:][keep], final[keep], rtol=1e-4, atol=1e-29)


def test_device_rhs_matches_oracle_F_componentwise(lib, oracle, net):
    """Kernel-level parity of F (chemistry.f90:294-352 + odes.f90) at 18 states: golden-trajectory abundances
    evaluated under cold / warm / hot physics, so that both branches of the three-phase transfer (net surface
    growth S >= 0 and S < 0) occur.  Per component: 1e-12 relative, plus the rounding of the component's own
    sum -- ydot_i is a difference of gains and losses that the device adds in a different order (team
    reductions), so a few ulps of the GROSS sum is the floor for any reordering of the reference's loop."""
    from uclchem_b200 import symbolic
    from uclchem_b200.table_emulator import COV0, TableEngine
    sym = symbolic.build(net)
    eng = TableEngine(sym)
    gold = np.load(GOLDEN / "static_full.npz")
    rows = [0, 3, 12, 25, 46, 60]
    physics = [dict(STATIC), dict(STATIC, initialDens=1e5, initialTemp=50.0, zeta=10.0),
               dict(STATIC, initialDens=1e6, initialTemp=200.0, zeta=100.0, radfield=3.0, baseAv=1.0)]
    signs = set()
    u = np.finfo(float).eps
    for pd_ in physics:
        p1 = params_from_dict(pd_)
        ys = np.array([np.append(np.maximum(gold["abund"][r], 1e-30), pd_["initialDens"]) for r in rows])
        got = lib.probe_rhs(np.repeat(p1, len(ys), axis=1), ys)
        for k in range(len(ys)):
            ref = oracle.probe_rhs(p1[:, 0], ys[k, :335])
            rate = oracle.get_rates(p1[:, 0], ys[k, :335])
            y = ys[k].copy()
            y[sym.iB], y[sym.iS] = y[net.bulk_list].sum(), y[net.surface_list].sum()
            ye, _ = eng.ext_state(y, rate)
            flux = rate * np.prod(ye[sym.flux_f], axis=1)
            rows_ = np.repeat(np.arange(net.nspec), np.diff(sym.g_ptr))
            gross = np.zeros(336)
            gross[:335] = np.bincount(rows_, weights=np.abs(flux[sym.g_reac]), minlength=335)
            net_ = np.bincount(rows_, weights=flux[sym.g_reac] * sym.g_sign, minlength=335)
            S = net_[net.surface_list].sum()
            signs.add(bool(S < 0))
            gs, gb = gross[net.surface_list].sum(), gross[net.bulk_list].sum()
            q = min(1.0, max(1e-30, y[sym.iB]) / max(1e-30, y[sym.iS])) / max(1e-30, y[sym.iB]) if S < 0 else COV0
            src = net.bulk_list if S < 0 else net.surface_list
            gross[net.surface_list] += gs * q * y[src]
            gross[net.bulk_list] += gs * q * y[src]
            gross[sym.iS] = gs + gs * q * y[src].sum()
            gross[sym.iB] = gb + gs * q * y[src].sum()
            tol = 1e-12 * np.abs(ref) + 64 * u * gross
            bad = np.abs(got[k] - ref) > tol
            assert not bad.any(), (pd_["initialTemp"], rows[k], np.where(bad)[0][:5], got[k][bad][:5], ref[bad][:5])
            well = gross[:335] < 1e2 * np.abs(ref[:335])       # components without cancellation: plain relative error
            assert np.abs(got[k][:335][well] / ref[:335][well] - 1).max() < 1e-12
    assert signs == {True, False}


def test_config2_cells_at_1myr_against_oracle_fixture(lib):
    """1 Myr config-2 cells, including the regions where the integration is hard, against the oracle run to
    completion offline (tools/make_config2_fixture.py -> tests/golden/config2_1myr_cells.npz): cells 8679, 7245,
    5277 cost the oracle 56 k - 155 k steps (finite-difference Jacobian), cells 6179, 4146, 5721, 8361 are cells
    the round-1 engine needed more than 30 000 steps for.  North-star bar: 0.01 dex above 1e-15."""
    from bench import config2_params
    fx = np.load(GOLDEN / "config2_1myr_cells.npz")
    cells = fx["cells"]
    assert (fx["flag"] == 0).all() and len(cells) >= 14
    P = config2_params()
    out = lib.run_grid(0, np.ascontiguousarray(P[:, cells]), step_budget=1000000)
    assert (out["flag"] == 0).all(), dict(zip(cells.tolist(), out["flag"].tolist()))
    worst = {int(c): max_dex(out["y_final"][k, :335], fx["y_final"][k, :335]) for k, c in enumerate(cells)}
    assert max(worst.values()) <= DEX_TOL, worst


def test_collapse_model_against_oracle(lib, oracle):
    """collapse.f90 (north-star `collapse`; no golden exists: parity pinned only via the self-validated oracle).
    BE4 at full length (0.97 x 1.855e5 yr, 184 output intervals: Bonnor-Ebert profile, enclosed-mass radius search),
    filament and ambipolar over their first 2e4 yr, plus the reference's refusal of an unknown mode."""
    p = params_from_dict({"collapse_mode": [2, 3, 4, 9], "rout": [0.2, 0.2, 0.5, 0.2], "baseAv": 1.0, "initialTemp": 10.0,
                          "finalTime": 2.0e4})
    out = lib.run_grid(3, p, timepoints=400, want_physics=True, want_chem=True)
    assert list(out["flag"]) == [0, 0, 0, -2]
    for c in range(3):
        r = oracle.run_model(3, p[:, c], timepoints=400)
        n = r["physics"].shape[0]
        assert r["flag"] == 0 and out["stats"][c][7] == r["stats"]["nintervals"] == n - 1
        np.testing.assert_allclose(out["physics"][c, :n, 0], r["physics"][:, 0], rtol=1e-12)          # output times
        np.testing.assert_allclose(out["physics"][c, :n, [1, 4]], r["physics"][:, [1, 4]].T, rtol=1e-9)  # density, Av
        assert max_dex(out["y_final"][c, :335], r["y_final"][:335]) < DEX_TOL, c
        worst = max(max_dex(out["abund"][c, row], r["abund"][row]) for row in range(1, n))
        assert worst < DEX_TOL, (c, worst)
    assert out["phys_final"][0, 0] > 0.97 * 1.855e5 - 1.0       # BE4 ignores finalTime (collapse.f90:39-41)
    res = model.collapse("BE4", None, param_dict={"rout": 0.2, "baseAv": 1.0}, out_species=["CO"])
    assert res[0] == 0 and res[1] == pytest.approx(out["y_final"][0, 49], rel=1e-12)


def test_jshock_model_against_oracle(lib, oracle):
    """jshock.f90 (no golden exists: parity pinned only via the self-validated oracle): heating to 5e3 (v/10)^2 K
    inside the shock width, cooling / compression phase, sputtering with the shock velocity."""
    sc = np.load(GOLDEN / "shockstart.npy")
    p = params_from_dict({"initialDens": [1e3, 1e4], "initialTemp": 10.0, "finalTime": [1e4, 1e3], "shock_vel": [10.0, 30.0],
                          "reltol": 1e-6, "abstol_min": 1e-20})
    y0 = np.stack([np.append(sc, 1e3), np.append(sc, 1e4)])
    out = lib.run_grid(4, p, y0=y0, timepoints=1000, want_physics=True, want_chem=True)
    assert (out["flag"] == 0).all()
    for c in range(2):
        r = oracle.run_model(4, p[:, c], y0=y0[c], timepoints=1000)
        n = r["physics"].shape[0]
        assert r["flag"] == 0 and out["stats"][c][7] == n - 1
        np.testing.assert_allclose(out["physics"][c, :n, :4], r["physics"][:, :4], rtol=1e-9)   # time, density, temperatures
        assert out["physics"][c, :n, 2].max() == pytest.approx(5e3 * (p[PARAM_INDEX["shock_vel"], c] / 10) ** 2, rel=1e-3)
        assert max_dex(out["y_final"][c, :335], r["y_final"][:335]) < DEX_TOL, c
    res = model.jshock(10.0, param_dict={"initialDens": 1e3, "finalTime": 1.0, "reltol": 1e-6, "abstol_min": 1e-20},
                       return_array=True, starting_chemistry=sc)
    assert len(res) == 5 and res[-1] == 0


def test_hot_core_g3_full_length_with_a_failed_dvode_call(lib):
    """G3 at full length on the engine: hot_core(3, 300) from startcollapse, 1 Myr, all 282 stored times of the
    reference's phase2-full.dat.  Run at reltol 1e-8 (no DVODE call fails on the engine) and at 1.03e-8, where one
    call fails inside the mantle's evaporation and integrateODESystem's retry policy (chemistry.f90:246-292) takes
    over: both must reproduce the reference's file (first seen on hardware: <= 5e-5 and 1.0e-3 dex)."""
    gold = np.load(GOLDEN / "phase2_full.npz")
    sc = np.load(GOLDEN / "startcollapse.npy")
    rts = [1e-8, 1.03e-8]
    p = params_from_dict({"endAtFinalDensity": False, "freefall": False, "initialDens": 1e5, "initialTemp": 10.0,
                          "finalDens": 1e5, "finalTime": 1.0e6, "freezeFactor": 0.0, "thermdesorb": True, "temp_indx": 3,
                          "max_temperature": 300.0, "reltol": rts})
    y0 = np.repeat(np.append(sc, 1e5)[None, :], len(rts), axis=0)
    out = lib.run_grid(1, p, y0=y0, timepoints=500, want_physics=True, want_chem=True)
    assert (out["flag"] == 0).all() and (out["stats"][:, 7] == 282).all()
    from uclchem_b200._capi import STAT_FIELDS
    fails = out["stats"][:, STAT_FIELDS.index("nfailcall")]
    assert fails[1] >= 1, fails        # the policy is exercised
    for c in range(len(rts)):
        np.testing.assert_allclose(out["physics"][c, 1:283, 2], gold["physics"][1:, 2], atol=6e-3)
        worst = max(max_dex(out["abund"][c, row], gold["abund"][row]) for row in range(1, 283))
        assert worst < DEX_TOL, (rts[c], worst)


def test_cshock_through_three_dissipation_times(lib, oracle):
    """C-shock past both cadence branches of cshock.f90:149-156 (steps of timestep_factor * t_diss up to 2 t_diss,
    then x1.1) and through the peak of the ion-neutral drift velocity, at a shock speed above the 19 km/s
    vaporisation threshold of sputtering.f90 (refractory ice species are sputtered too)."""
    sc = np.load(GOLDEN / "shockstart.npy")
    p = params_from_dict({"initialDens": 1e5, "initialTemp": 10.0, "finalTime": 1.0, "shock_vel": 30.0, "reltol": 1e-6,
                          "abstol_min": 1e-20})
    y0 = np.append(sc, 1e5)[None, :]
    tdiss = lib.run_grid(2, p, y0=y0)["dissipation_time"][0]
    p[PARAM_INDEX["finaltime"]] = 3.0 * tdiss
    out = lib.run_grid(2, p, y0=y0, timepoints=600, want_physics=True, want_chem=True)
    r = oracle.run_model(2, p[:, 0], y0=y0[0], timepoints=600)
    n = r["physics"].shape[0]
    assert out["flag"][0] == 0 and r["flag"] == 0 and out["stats"][0][7] == n - 1 and n > 205
    t = out["physics"][0, :n, 0]
    assert t[-1] > 3.0 * tdiss and np.isclose(t[-1] / t[-2], 1.1, rtol=1e-3) and np.isclose(t[100] - t[99], 0.01 * tdiss, rtol=1e-6)
    np.testing.assert_allclose(out["physics"][0, :n, :4], r["physics"][:, :4], rtol=1e-6)
    assert out["physics"][0, :n, 2].max() > 1000.0          # the gas was shock heated
    worst = max(max_dex(out["abund"][0, row], r["abund"][row]) for row in range(1, n))
    assert worst < DEX_TOL and max_dex(out["y_final"][0, :335], r["y_final"][:335]) < DEX_TOL, worst


def test_rate_coefficient_overrides_against_oracle(lib, net):
    """Per-reaction alpha / beta / gamma dictionaries of the reference's parameter dictionary (wrap.f90:744-761,
    985-1019; keys are 1-based reaction indices of network.f90): the device uses an overridden copy of the three
    tables; the oracle is handed a network whose arrays were edited the same way."""
    import copy
    from oracle.oracle import Oracle
    from uclchem_b200._capi import UclgpuError
    r_cr, r_two, r_frz = net.type_ranges["CRP"][0] + 3, net.type_ranges["TWOBODY"][0] + 40, net.type_ranges["FREEZE"][0] + 5
    over = {"alpha": {r_cr + 1: 10.0 * net.alpha[r_cr], r_frz + 1: 0.1}, "beta": {r_two + 1: 0.5}, "gamma": {r_two + 1: 50.0}}
    pd_ = {"initialDens": 1e5, "initialTemp": 20.0, "finalTime": 1e4, **over}
    phys, chem, _, start, flag = model.cloud(param_dict=pd_, return_array=True)
    net2 = copy.deepcopy(net)
    net2.alpha[r_cr], net2.alpha[r_frz], net2.beta[r_two], net2.gama[r_two] = 10.0 * net.alpha[r_cr], 0.1, 0.5, 50.0
    p = params_from_dict({k: v for k, v in pd_.items() if k not in over})
    ref = Oracle(net2).run_model(0, p[:, 0])
    base = Oracle(net).run_model(0, p[:, 0])
    assert flag == 0 and ref["flag"] == 0
    assert max_dex(start, ref["y_final"][:335]) < DEX_TOL
    assert max_dex(start, base["y_final"][:335]) > 0.05                  # the overrides matter
    # rates seen by the integrator carry the overrides (row 1 of the rate trajectory)
    _, _, rates, _, _ = model.cloud(param_dict=dict(pd_, finalTime=1.0), return_array=True, return_rates=True)
    assert rates[1, 0, r_cr] == pytest.approx(10.0 * net.alpha[r_cr] * 1.0, rel=1e-12)     # CRP: alpha * zeta
    # gamma of a surface two-body reaction feeds tables fixed at generation time: refused, not half-applied
    with pytest.raises(UclgpuError):
        model.cloud(param_dict={"finalTime": 1.0, "gamma": {net.type_ranges["LH"][0] + 1: 10.0}}, return_array=True)


def test_postprocess_model_against_oracle(lib, oracle):
    """postprocess.f90 (no golden exists: parity pinned only via the self-validated oracle): two tracers with
    different histories in one call, without and with supplied shielding column densities (which also freeze the
    H2 / CO photodissociation rates inside F, chemistry.f90:310-319)."""
    n, spy = 30, 3.16e7
    t = np.linspace(0.0, 2.9e4, n) * spy
    ramp = t / t[-1]
    g = np.zeros((2, 10, n))
    # a cold tracer that is compressed tenfold, and a diffuse one that is heated to 70 K (below 3e4 cm^-3, i.e. outside
    # the band where DVODE stalls: a 3e5 cm^-3 tracer heated through 34-48 K costs either arm > 1e5 steps)
    for c, (d0, dfac, dt_) in enumerate(((1e4, 9.0, 20.0), (3e3, 2.0, 60.0))):
        g[c, 0], g[c, 1], g[c, 2] = t, d0 * (1 + dfac * ramp), 10 + dt_ * ramp
        g[c, 3], g[c, 4], g[c, 5] = g[c, 2], 1.0 + c, 1.0 + 9 * c
        g[c, 6] = 1e21 * (1 + c) * (1 + ramp)
        g[c, 7], g[c, 8], g[c, 9] = 0.4 * g[c, 6], 1e-5 * g[c, 6], 1e-6 * g[c, 6]
    p = params_from_dict({"initialDens": [1e4, 3e3], "initialTemp": 10.0})
    oracle.set_deadline(240.0)      # guard: a stalling model must fail the test, not hang it
    for use in (False, True):
        out = lib.run_grid(5, p, timepoints=n, want_physics=True, want_chem=True, pp_grid=g, pp_coldens=use, step_budget=300000)
        assert (out["flag"] == 0).all() and (out["stats"][:, 7] == n).all()
        for c in range(2):
            r = oracle.run_model(5, p[:, c], timepoints=n, pp_grid=g[c], pp_coldens=use)
            assert r["flag"] == 0 and r["physics"].shape[0] == n + 1
            np.testing.assert_allclose(out["physics"][c, : n + 1, :7], r["physics"][:, :7], rtol=1e-12)
            worst = max(max_dex(out["abund"][c, row], r["abund"][row]) for row in range(1, n + 1))
            assert worst < DEX_TOL, (use, c, worst)
    oracle.set_deadline(0.0)
    res = model.postprocess(param_dict={"initialDens": 1e4}, out_species=["CO"], time_array=g[0, 0], density_array=g[0, 1],
                            gas_temperature_array=g[0, 2], dust_temperature_array=g[0, 3], zeta_array=g[0, 5],
                            radfield_array=g[0, 4])
    assert res[0] == 0 and res[1] > 0
    grid = model.postprocess_grid({"initialDens": [1e4, 3e3]}, g[:, 0], g[:, 1], g[:, 2], g[:, 3], g[:, 5], g[:, 4], out_species=["CO"])
    assert (grid["flag"] == 0).all() and grid["out_species"][0, 0] == pytest.approx(res[1], rel=1e-12)


def test_cosmic_ray_attenuation_models_against_oracle(lib, oracle):
    """physics-core.f90:121-157 (ionizationDependency): column-density dependent H2 cosmic-ray dissociation rate for
    both ionisation models, refreshed after every output interval; free fall makes the column grow with time."""
    p = params_from_dict({"initialDens": [1e3, 1e4], "finalDens": 1e6, "freefall": True, "endAtFinalDensity": False,
                          "initialTemp": 15.0, "finalTime": 3e5, "cosmicRayAttenuation": True, "ionModel": ["L", "H"],
                          "improvedH2CRPDissociation": True, "zeta": [1.0, 5.0]})
    out = lib.run_grid(0, p, step_budget=300000)
    ref, pf, flag, _ = oracle.run_grid(0, p, nthreads=2)
    assert (out["flag"] == 0).all() and (flag == 0).all()
    for c in range(2):
        np.testing.assert_allclose(out["phys_final"][c, :7], pf[c, :7], rtol=1e-6)
        assert max_dex(out["y_final"][c, :335], ref[c, :335]) < DEX_TOL, c
    # the reference refuses improvedH2CRPDissociation without cosmicRayAttenuation (physics-core.f90:60-64)
    bad = params_from_dict({"improvedH2CRPDissociation": True, "finalTime": 1.0})
    assert lib.run_grid(0, bad)["flag"][0] == -2

"""The engine's deviations from DVODE MF=22 -- analytic Jacobian, bordered unknowns [y, S, tau], BULK / SURFACE
constraint rows, fixed-pattern sparse LU without pivoting + explicit inverse of the dense block -- run INSIDE the
oracle's DVODE on the CPU: the table emulator (the programs the kernel executes) is plugged in through the oracle's
debug hook `orc_set_linalg_hook` in place of the finite-difference Jacobian and LINPACK.  Same step controller,
same RHS, only the linear algebra differs, so this pins the generated Jacobian / LU programs dynamically, over a
whole integration, without a GPU.  (tools/study_engine_linalg.py is the same harness with diagnostics; on the B200
it reproduced the device's Newton trace digit for digit for hundreds of iterations -- DESIGN.md section 9.)"""
import ctypes as C

import numpy as np
import pytest
from conftest import ROOT, max_dex

from uclchem_b200 import symbolic
from uclchem_b200.network import Network
from uclchem_b200.params import params_from_dict
from uclchem_b200.table_emulator import TableEngine

SETUP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_int)
SOLVE = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def harness():
    from oracle.oracle import Oracle
    net = Network.from_json(ROOT / "uclchem_b200" / "networks" / "default.json")
    orc = Oracle(net)
    sym = symbolic.build(net)
    eng = TableEngine(sym)
    neq, naug = sym.neq, sym.naug
    surf, bulk = np.asarray(net.surface_list), np.asarray(net.bulk_list)
    orc.lib.orc_ctx_rate.restype = C.POINTER(C.c_double)
    orc.lib.orc_ctx_rate.argtypes = [C.c_void_p]
    st = {"nsing": 0}

    def setup(ctx, yp, gamma, fresh):
        if fresh:   # DVODE's saved Jacobian (JSV = 1): J depends on the state and rates of the last evaluation only
            st["y"] = np.ctypeslib.as_array(yp, (neq,)).copy()
            st["rate"] = np.ctypeslib.as_array(orc.lib.orc_ctx_rate(ctx), (net.nreac,)).copy()
        with np.errstate(all="ignore"):
            st["fv"] = eng.factor(eng.assemble(st["y"], st["rate"], gamma))
        if not np.isfinite(st["fv"]).all():
            st["nsing"] += 1
            return 1
        return 0

    def solve(ctx, bp):
        b = np.ctypeslib.as_array(bp, (neq,))
        ba = np.zeros(naug)
        ba[:neq] = b
        ba[sym.iS] = b[sym.iS] - b[surf].sum()   # constraint rows: r - sum(member r), engine_la.cuh newton_rhs
        ba[sym.iB] = b[sym.iB] - b[bulk].sum()
        b[:] = eng.solve(st["fv"], ba)[:neq]

    cbs = (SETUP(setup), SOLVE(solve))

    def run(pd_, hooked):
        orc.lib.orc_set_linalg_hook(cbs[0] if hooked else SETUP(), cbs[1] if hooked else SOLVE())
        try:
            return orc.run_model(0, params_from_dict(pd_)[:, 0])
        finally:
            orc.lib.orc_set_linalg_hook(SETUP(), SOLVE())
    return net, run, st


@pytest.mark.parametrize("pd_", [
    {"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 1e3},                                  # config[0] cell
    {"initialDens": 1e6, "initialTemp": 60.0, "zeta": 30.0, "radfield": 3.0, "finalTime": 1e2},   # warm, mantle evaporating
])
def test_engine_linear_algebra_inside_the_reference_integrator(harness, pd_):
    net, run, st = harness
    ref = run(pd_, hooked=False)
    got = run(pd_, hooked=True)
    assert ref["flag"] == 0 and got["flag"] == 0 and st["nsing"] == 0
    assert max_dex(got["y_final"][: net.nspec], ref["y_final"][: net.nspec]) < 1e-3
    # same algorithm, same amount of work: steps within 10 %, far fewer RHS calls (no finite differences)
    assert abs(got["stats"]["nst"] - ref["stats"]["nst"]) <= 0.1 * ref["stats"]["nst"], (got["stats"], ref["stats"])
    assert got["stats"]["nfe"] < 0.2 * ref["stats"]["nfe"]

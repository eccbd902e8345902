"""CPU study (no GPU): the oracle's DVODE with the ENGINE's linear algebra plugged in -- analytic Jacobian, bordered
unknowns, BULK / SURFACE constraint rows, unpivoted sparse LU + explicit dense inverse, all through the table
emulator (the programs the kernel runs) -- via the oracle's experiment hook (orc_set_linalg_hook).  Used to
reproduce, on the CPU, cells on which the engine and the oracle part ways.

    python tools/study_engine_linalg.py tag dens temp zeta final_time reltol abstol_min [mode]
mode: engine (default) | dense (analytic J, LAPACK solve of the same bordered system) | plain (oracle untouched)
"""
import ctypes as C, os, sys, functools, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
tag = sys.argv[1]
dens, temp, zeta, tfin, rt, am = (float(a) for a in sys.argv[2:8])
mode = sys.argv[8] if len(sys.argv) > 8 else "engine"
if not os.environ.get("STUDY_NO_TRACE"):
    os.environ["ORC_TRACE"] = f"/tmp/study_{tag}_{mode}.txt"
from oracle.oracle import Oracle
from uclchem_b200 import symbolic
from uclchem_b200.network import Network
from uclchem_b200.params import params_from_dict
from uclchem_b200.table_emulator import TableEngine

net = Network.from_json(ROOT / "uclchem_b200" / "networks" / f"{tag}.json")
orc = Oracle(net)
sym = symbolic.build(net)
eng = TableEngine(sym)
neq, naug, nreac = sym.neq, sym.naug, net.nreac
surf, bulk = np.asarray(net.surface_list), np.asarray(net.bulk_list)
orc.lib.orc_ctx_rate.restype = C.POINTER(C.c_double)
orc.lib.orc_ctx_rate.argtypes = [C.c_void_p]
SETUP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_int)
SOLVE = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double))
state = dict(nsetup=0, nsolve=0, nsing=0)


def setup(ctx, yp, gamma, fresh):
    state["nsetup"] += 1
    if fresh:
        state["jy"] = np.ctypeslib.as_array(yp, (neq,)).copy()
        state["jrate"] = np.ctypeslib.as_array(orc.lib.orc_ctx_rate(ctx), (nreac,)).copy()
    val = eng.assemble(state["jy"], state["jrate"], gamma)
    if os.environ.get("STUDY_DUMP"):   # keep the last Jacobian states; dump them when BULK first turns negative from a healthy value
        hist = state.setdefault("hist", [])
        hist.append((np.ctypeslib.as_array(yp, (neq,)).copy(), state["jrate"].copy(), gamma))
        del hist[:-400]
        yb = [h[0][sym.iB] for h in hist]
        if not state.get("dumped") and yb[-1] < 0 and max(yb) > 2e-7:
            state["dumped"] = True
            np.savez(os.environ["STUDY_DUMP"], y=np.array([h[0] for h in hist]), rate=np.array([h[1] for h in hist[-1:]]), gamma=np.array([h[2] for h in hist]))
            print("dumped", len(hist), "states; BULK", yb[-5:])
    if mode == "dense":
        state["A"] = eng.to_dense(val)
        return 0
    with np.errstate(all="ignore"):
        fv = eng.factor(val)
    state["fv"] = fv
    if mode == "check":
        state["A"] = eng.to_dense(val); state["gamma"] = gamma; state["fresh"] = fresh
    if not np.isfinite(fv).all():
        state["nsing"] += 1
        return 1
    return 0


def solve(ctx, bp):
    state["nsolve"] += 1
    b = np.ctypeslib.as_array(bp, (neq,))
    ba = np.zeros(naug)
    ba[:neq] = b
    ba[sym.iS] = b[sym.iS] - b[surf].sum()
    ba[sym.iB] = b[sym.iB] - b[bulk].sum()
    with np.errstate(all="ignore"):
        if mode == "dense":
            x = np.empty(naug); x[sym.perm] = np.linalg.solve(state["A"], ba[sym.perm])
        else:
            x = eng.solve(state["fv"], ba)
            if mode == "check":
                A = state["A"]; bp_ = ba[sym.perm]
                xr = np.linalg.solve(A, bp_)
                Al = A.astype(np.longdouble)
                for _ in range(3):   # iterative refinement with extended-precision residuals: the reference solution
                    res = (bp_.astype(np.longdouble) - Al @ xr.astype(np.longdouble)).astype(np.float64)
                    xr = xr + np.linalg.solve(A, res)
                xl = np.linalg.solve(A, bp_)
                xt = np.empty(naug); xt[sym.perm] = xr
                xp = np.empty(naug); xp[sym.perm] = xl
                yj = state["jy"]
                ewt = rt * np.abs(yj) + np.maximum(1e-10 * np.abs(yj), am)
                wr = lambda v: float(np.sqrt(np.mean((v[:neq] / ewt) ** 2)))
                e_eng, e_lap, nx = wr(x - xt), wr(xp - xt), wr(xt)
                state.setdefault("log", []).append((state["nsolve"], state["gamma"], nx, e_eng, e_lap, yj[sym.iB], yj[sym.iS]))
                if (e_eng > 0.05 or e_lap > 0.05) and yj[sym.iB] > 0 and yj[sym.iS] > 0 and state.setdefault("nbad", 0) < 60:
                    state["nbad"] += 1
                    i = int(np.argmax(np.abs((x - xt)[:neq]) / ewt))
                    print(f"solve {state['nsolve']} gamma {state['gamma']:.3e}: |x| {nx:.2e} wrms err engine {e_eng:.2e} lapack {e_lap:.2e}; worst {net.names[i] if i < net.nspec else i} "
                          f"x {x[i]:.3e} true {xt[i]:.3e} ewt {ewt[i]:.1e}  BULK {yj[sym.iB]:.3e} SURF {yj[sym.iS]:.3e}")
                xr = xt
                if os.environ.get("STUDY_USE_REF"): x = xr
    b[:] = x[:neq]


cb = (SETUP(setup), SOLVE(solve))
if mode != "plain":
    orc.lib.orc_set_linalg_hook(cb[0], cb[1])
import json
p = params_from_dict(dict({"initialDens": dens, "initialTemp": temp, "zeta": zeta, "finalTime": tfin, "reltol": rt, "abstol_min": am},
                          **json.loads(os.environ.get("STUDY_EXTRA", "{}"))))   # e.g. the config-2 grid: {"radfield": 1.0, "baseAv": 2.0, "rout": 0.05}
t0 = time.time()
orc.set_deadline(float(os.environ.get("STUDY_SECONDS", "600")))
r = orc.run_model(0, p[:, 0])
print(f"mode {mode}: flag {r['flag']} stats {r['stats']}  hook calls {state['nsetup']} setups ({state['nsing']} singular) {state['nsolve']} solves  {time.time() - t0:.0f} s")
y = r["y_final"]
print("BULK", y[sym.iB], "sum bulk", y[bulk].sum(), "SURFACE", y[sym.iS], "sum surf", y[surf].sum())
np.save(f"/tmp/study_{tag}_{mode}_y.npy", y)

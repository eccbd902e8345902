"""CPU study (no GPU): how accurate is the engine's unpivoted factorisation of P = I - gamma J along the trajectory of
a cell, as a function of gamma?  States from the oracle, matrix / factor / solve through the table emulator (the
same programs the kernel runs), reference solution from LAPACK with partial pivoting.

    python tools/study_nopivot_accuracy.py [tag] [dens temp zeta]
"""
import sys, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from oracle.oracle import Oracle
from uclchem_b200 import symbolic
from uclchem_b200.network import Network
from uclchem_b200.params import params_from_dict
from uclchem_b200.table_emulator import TableEngine

tag = sys.argv[1] if len(sys.argv) > 1 else "crp_photo"
dens, temp, zeta = (float(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (1e7, 100.0, 1e3)
rt, am = 1e-5, 1e-15
net = Network.from_json(ROOT / "uclchem_b200" / "networks" / f"{tag}.json")
orc = Oracle(net)
sym = symbolic.build(net)
eng = TableEngine(sym)
rng = np.random.default_rng(1)
for tyr in (1e-2, 1.0, 10.0, 100.0, 1e3):
    pd_ = {"initialDens": dens, "initialTemp": temp, "zeta": zeta, "finalTime": tyr, "reltol": rt, "abstol_min": am}
    p = params_from_dict(pd_)
    r = orc.run_model(0, p[:, 0])
    y = r["y_final"].copy()
    rate = orc.get_rates(p[:, 0], y)
    f, S = eng.rhs(y, rate)
    ewt = rt * np.abs(y) + np.maximum(1e-10 * np.abs(y), am)
    ewt_aug = np.concatenate([ewt, np.full(sym.naug - len(ewt), 1.0)])
    print(f"t = {tyr:g} yr: oracle flag {r['flag']} nst {r['stats']['nst']}  S = {S:.3e}")
    for gamma in 10.0 ** np.arange(2, 13):
        val = eng.assemble(y, rate, gamma)
        A = eng.to_dense(val)
        fv = eng.factor(val)
        b = np.zeros(sym.naug); b[: sym.neq] = gamma * f
        x = eng.solve(fv, b)
        xr = np.empty(sym.naug); xr[sym.perm] = np.linalg.solve(A, b[sym.perm])
        e1 = np.sqrt(np.mean(((x - xr)[: sym.neq] / ewt) ** 2)); n1 = np.sqrt(np.mean((xr[: sym.neq] / ewt) ** 2))
        b2 = np.zeros(sym.naug); b2[: sym.neq] = ewt * rng.standard_normal(sym.neq)
        x2 = eng.solve(fv, b2)
        xr2 = np.empty(sym.naug); xr2[sym.perm] = np.linalg.solve(A, b2[sym.perm])
        e2 = np.sqrt(np.mean(((x2 - xr2)[: sym.neq] / ewt) ** 2)); n2 = np.sqrt(np.mean((xr2[: sym.neq] / ewt) ** 2))
        piv = fv[sym.diag_pos[: sym.n0]]          # reciprocals of the sparse pivots
        print(f"   gamma {gamma:8.1e}: cond {np.linalg.cond(A):9.2e}  min|pivot| {1.0 / np.abs(piv).max():9.2e}  "
              f"newton rhs: |x| {n1:9.2e} err {e1:9.2e} ({e1 / max(n1, 1e-300):8.1e})   random rhs: |x| {n2:9.2e} err {e2:9.2e} ({e2 / max(n2, 1e-300):8.1e})")

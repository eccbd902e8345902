"""CPU study for the next kernel step: replace the 21 dependent levels of the
triangular solves by explicit sparse inverses of the factors of the 247 sparse pivots.

For the golden states (tests/golden/getydot_cases.npz) and a range of gamma = h*rl1 it builds
P = I - gamma*J with the table emulator, factors it on the generated pattern, extracts
L11 (unit lower), U11, U12, L21 from the factored storage, forms the explicit inverses on their
closure patterns and compares  x = P^-1 b  from the level-scheduled substitution with the
inverse-based product form:  y1 = Linv b1;  b2' = b2 - L21 y1;  x2 = Tinv b2';  x1 = Uinv (y1 - U12 x2).
Reports pattern sizes (they decide where the values can live) and the accuracy of the product form.
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
from uclchem_b200 import symbolic
from uclchem_b200.network import load_default
from uclchem_b200.table_emulator import TableEngine

net = load_default(); sym = symbolic.build(net); eng = TableEngine(sym)
n0, m, naug = sym.n0, sym.m, sym.naug
g = np.load(ROOT / "tests/golden/getydot_cases.npz")

def factors(val):
    """Dense copies of L11 (unit lower), U11 (with true diagonal), U12, L21 from factored storage."""
    L = np.eye(n0); U = np.zeros((n0, n0)); U12 = np.zeros((n0, m)); L21 = np.zeros((m, n0))
    for lv in sym.fwd_levels:
        for e, r in enumerate(lv["rows"]):
            a, b = lv["ptr"][e], lv["ptr"][e + 1]
            L[r, lv["cols"][a:b]] = val[lv["pos"][a:b]]
    for t in range(m):
        a, b = sym.tail_l_ptr[t], sym.tail_l_ptr[t + 1]
        L21[t, sym.tail_l_col[a:b]] = val[sym.tail_l_pos[a:b]]
    for i in range(n0):
        U[i, i] = 1.0 / val[sym.diag_pos[i]]      # pivots are stored as reciprocals
    for lv in sym.bwd_levels:
        for e, r in enumerate(lv["rows"]):
            a, b = lv["ptr"][e], lv["ptr"][e + 1]
            for p, c in zip(lv["pos"][a:b], lv["cols"][a:b]):
                if c < n0: U[r, c] = val[p]
                else: U12[r, c - n0] = val[p]
    return L, U, U12, L21

def closure(T, lower):
    """Structural pattern of inv(T) for a triangular pattern T (reachability)."""
    B = (T != 0)
    R = np.eye(n0, dtype=bool)
    order = range(n0) if lower else range(n0 - 1, -1, -1)
    for i in order:   # row i of inv depends on rows j with T[i, j] != 0
        for j in np.where(B[i])[0]:
            if j != i: R[i] |= R[j]
    return R

Lp = Up = None
rng = np.random.default_rng(1)
worst = 0.0
for case in range(6):
    y, rate = g[f"y_{case}"], g[f"rate_{case}"]
    y = y.copy(); y[sym.iB] = y[net.bulk_list].sum(); y[sym.iS] = y[net.surface_list].sum()
    for gamma in (1e2, 1e5, 1e8, 1e11, 1e13):
        val = eng.factor(eng.assemble(y, rate, gamma))
        L, U, U12, L21 = factors(val)
        if Lp is None:
            Lp, Up = closure(L, True), closure(U, False)
            print(f"n0 {n0}, dense block {m}; nnz offdiag L11 {int((L != 0).sum()) - n0}, U11 {int((U != 0).sum()) - n0}, "
                  f"U12 {int((U12 != 0).sum())}, L21 {int((L21 != 0).sum())}")
            print(f"closure patterns: inv(L11) offdiag {int(Lp.sum()) - n0}, inv(U11) incl. diag {int(Up.sum())}; "
                  f"L21*inv(L11) {int(((L21 != 0).astype(int) @ Lp.astype(int) != 0).sum())} of {m * n0}, "
                  f"inv(U11)*U12 {int((Up.astype(int) @ (U12 != 0).astype(int) != 0).sum())} of {m * n0}")
        Linv = np.linalg.inv(L) * Lp
        Uinv = np.linalg.inv(U) * Up
        Tinv = val[sym.off_dense: sym.off_dense + m * m].reshape(m, m)
        for trial in range(3):
            b = np.zeros(naug); b[:net.nspec + 1] = rng.standard_normal(net.nspec + 1) * (np.abs(y[:net.nspec + 1]) + 1e-20)
            b[sym.iB] = 0; b[sym.iS] = 0
            x_ref = eng.solve(val, b)
            bn = b[sym.perm]
            y1 = Linv @ bn[:n0]
            b2 = bn[n0:] - L21 @ y1
            x2 = Tinv @ b2
            x1 = Uinv @ (y1 - U12 @ x2)
            x = np.empty(naug); x[sym.perm] = np.concatenate([x1, x2])
            scale = np.abs(x_ref) + 1e-300
            big = np.abs(x_ref) > 1e-12 * np.abs(x_ref).max()
            err = float((np.abs(x - x_ref) / scale)[big].max())
            worst = max(worst, err)
        print(f"case {case} gamma {gamma:.0e}: max rel. difference substitution vs product form (entries > 1e-12 max) {err:.2e}; "
              f"max |Linv| {np.abs(Linv).max():.2e} max |Uinv| {np.abs(Uinv).max():.2e}")
print("worst", worst)

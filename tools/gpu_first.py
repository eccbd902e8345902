"""GPU bring-up helper: kernel-level parity against the oracle / one model / small grids.
usage: gpu_first.py probe | model <finalTime> | grid <ncell> <finalTime>"""
import functools, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from uclchem_b200.network import load_default
from uclchem_b200.params import params_from_dict
from uclchem_b200._capi import get_library
from oracle.oracle import Oracle
print = functools.partial(print, flush=True)
STAGE = sys.argv[1]
net = load_default(); orc = Oracle(net); lib = get_library(); lib.init()
gold = np.load(ROOT / 'tests/golden/static_full.npz')
from uclchem_b200._capi import STAT_FIELDS as STAT

if STAGE == "probe":
    p1 = params_from_dict({"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 1e6})
    rows = [0, 3, 12, 25, 46]
    ys = np.array([np.append(np.maximum(gold['abund'][r], 1e-30), 1e4) for r in rows])
    pp = np.repeat(p1, len(rows), axis=1)
    rg = lib.get_rates(pp, ys)
    for k, r in enumerate(rows):
        ro = orc.get_rates(p1[:, 0], ys[k, :335]); m = ro != 0
        rel = np.abs(rg[k][m] / ro[m] - 1)
        print('rates row', r, 'max rel err', rel.max(), 'zero-mismatch', int(((rg[k] == 0) != (ro == 0)).sum()))
    yd = lib.probe_rhs(pp, ys)
    from uclchem_b200 import symbolic
    from uclchem_b200.table_emulator import TableEngine
    sym = symbolic.build(net); eng = TableEngine(sym)
    for k, r in enumerate(rows):
        y = ys[k].copy(); y[sym.iB] = y[net.bulk_list].sum(); y[sym.iS] = y[net.surface_list].sum()
        ref, S = eng.rhs(y, rg[k])
        print('rhs row', r, 'max err/scale', np.abs(yd[k] - ref).max() / np.abs(ref).max(), 'S', S)
    # get_odes vs oracle get_odes (reference semantics: 1e-7 s pre-integration)
    go = lib.get_odes(pp, ys)
    for k, r in enumerate(rows):
        ro = orc.get_odes(p1[:, 0], ys[k, :335])
        print('get_odes row', r, 'max err/scale', np.abs(go[k] - ro).max() / np.abs(ro).max())
    rng = np.random.default_rng(0)
    for gamma in (1e3, 1e9):
        b = rng.standard_normal((len(rows), 336)) * np.abs(ys)
        x = lib.probe_newton(pp, ys, gamma, b)
        for k, r in enumerate(rows):
            y = ys[k].copy(); y[sym.iB] = y[net.bulk_list].sum(); y[sym.iS] = y[net.surface_list].sum()
            val = eng.assemble(y, rg[k], gamma)
            ba = np.zeros(sym.naug); ba[:336] = b[k]; ba[sym.iB] = 0; ba[sym.iS] = 0
            xe = eng.solve(eng.factor(val), ba)
            print('newton gamma', gamma, 'row', r, 'rel err', np.abs(x[k] - xe).max() / np.abs(xe).max())
elif STAGE == "model":
    final = float(sys.argv[2])
    p1 = params_from_dict({"initialDens": 1e4, "initialTemp": 10.0, "finalTime": final})
    t = time.time(); out = lib.run_grid(0, p1, timepoints=500, want_physics=True, want_chem=True); dt = time.time() - t
    print('1 model wall', dt, 'kernel ms', lib.last_kernel_ms(), 'flag', out['flag'], dict(zip(STAT, out['stats'][0])))
    ga = gold['abund']; n = int(out['stats'][0][7]) + 1; worst = 0
    for row in range(1, min(n, 47)):
        a = out['abund'][0, row]; bb = ga[row]; m = bb > 1e-15
        dex = np.abs(np.log10(a[m] / bb[m])); worst = max(worst, dex.max())
        print('row', row, gold['physics'][row, 0], out['physics'][0, row, 0], 'max dex', dex.max(), net.names[np.where(m)[0][dex.argmax()]])
    print('WORST DEX vs golden', worst)
elif STAGE == "grid":
    nc = int(sys.argv[2]); final = float(sys.argv[3])
    dens = 10 ** np.linspace(3, 7, nc)
    pg = params_from_dict({"initialDens": dens, "initialTemp": 10.0, "finalTime": final, "baseAv": 2.0})
    t = time.time(); o = lib.run_grid(0, pg); dt = time.time() - t
    ms, _ = lib.last_kernel_ms()
    print('grid', nc, 'wall', dt, 'kernel s', ms / 1e3, 'models/s', nc / (ms / 1e3), 'flags', np.unique(o['flag']),
          'mean nst', o['stats'][:, 0].mean(), 'max nst', o['stats'][:, 0].max(), 'mean nlu', o['stats'][:, 3].mean())

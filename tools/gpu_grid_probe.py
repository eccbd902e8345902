import sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import get_library, STAT_FIELDS
from uclchem_b200.params import PARAM_INDEX
lib = get_library(); lib.init()
P = config2_params()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
idx = np.linspace(0, P.shape[1]-1, n).astype(int)
p = np.ascontiguousarray(P[:, idx])
t = time.time(); o = lib.run_grid(0, p); dt = time.time()-t
ms,_ = lib.last_kernel_ms()
st = o['stats']; S = {k: st[:,i] for i,k in enumerate(STAT_FIELDS)}
print('cells', n, 'wall', dt, 'kernel s', ms/1e3, 'models/s', n/(ms/1e3), 'flags', np.unique(o['flag'], return_counts=True))
cyc = S['cyc_total']/1.9e9
order = np.argsort(cyc)[::-1]
print('per-cell seconds: median', np.median(cyc), 'mean', cyc.mean(), 'max', cyc.max(), 'sum/148', cyc.sum()/148)
for k in order[:12]:
    print(f"cell {idx[k]} dens {p[PARAM_INDEX['initialdens'],k]:.2e} T {p[PARAM_INDEX['initialtemp'],k]:.0f} zeta {p[PARAM_INDEX['zeta'],k]:.1f} sec {cyc[k]:.2f} nst {S['nst'][k]} nlu {S['nlu'][k]} nje {S['nje'][k]} ncfn {S['ncfn'][k]} netf {S['netf'][k]} failcalls {S['nfailcall'][k]} nsing {S['nsing'][k]}")
tot = {k: S[k].sum() for k in ('cyc_rates','cyc_rhs','cyc_jac','cyc_factor','cyc_dense','cyc_solve','cyc_total')}
print({k: round(v/tot['cyc_total'],3) for k,v in tot.items()})
np.savez('gpurun_out/grid_stats_%d.npz' % n, idx=idx, params=p, stats=st, flag=o['flag'], y_final=o['y_final'], kernel_ms=ms)
print('mean nst', S['nst'].mean(), 'mean nlu', S['nlu'].mean(), 'mean nni', S['nni'].mean(), 'mean nfe', S['nfe'].mean())

"""Run the cells of config 2 that exceeded a 30000-step budget (tools/config2_heavy_cells.npy) with a larger budget."""
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
idx = np.load(ROOT / "tools/config2_heavy_cells.npy")
budget = int(sys.argv[1])
p = np.ascontiguousarray(P[:, idx])
t = time.time(); o = lib.run_grid(0, p, step_budget=budget); dt = time.time() - t
st = o["stats"]; S = {k: st[:, i] for i, k in enumerate(STAT_FIELDS)}
print("cells", len(idx), "wall", dt, "flags", np.unique(o["flag"], return_counts=True))
cyc = S["cyc_total"] / 1.9e9
print("sec: median", np.median(cyc), "max", cyc.max(), "sum", cyc.sum(), "nst median", np.median(S["nst"]), "max", S["nst"].max())
np.savez(ROOT / "gpurun_out/heavy_stats.npz", idx=idx, stats=st, flag=o["flag"], y_final=o["y_final"])
for k in np.argsort(cyc)[::-1][:15]:
    print(f"cell {idx[k]} dens {p[PARAM_INDEX['initialdens'],k]:.2e} T {p[PARAM_INDEX['initialtemp'],k]:.0f} zeta {p[PARAM_INDEX['zeta'],k]:.1f} sec {cyc[k]:.1f} nst {S['nst'][k]} ncfn {S['ncfn'][k]} netf {S['netf'][k]} failcalls {S['nfailcall'][k]} nint {S['nintervals'][k]} flag {o['flag'][k]}")

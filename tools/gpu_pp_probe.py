"""Postprocess model on the GPU, step by step (run on the GPU box under an outer `timeout`)."""
import sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.params import params_from_dict
use = bool(int(sys.argv[1])); budget = int(sys.argv[2]); nt = int(sys.argv[3])
n, spy = 30, 3.16e7
t = np.linspace(0.0, 2.9e4, n) * spy
ramp = t / t[-1]
g = np.zeros((2, 10, n))
for c, (d0, dt_) in enumerate(((1e4, 20.0), (3e5, 60.0))):
    g[c, 0], g[c, 1], g[c, 2] = t, d0 * (1 + 9 * ramp), 10 + dt_ * ramp
    g[c, 3], g[c, 4], g[c, 5] = g[c, 2], 1.0 + c, 1.0 + 9 * c
    g[c, 6] = 1e21 * (1 + c) * (1 + ramp)
    g[c, 7], g[c, 8], g[c, 9] = 0.4 * g[c, 6], 1e-5 * g[c, 6], 1e-6 * g[c, 6]
g = np.ascontiguousarray(g[:, :, :nt])
p = params_from_dict({"initialDens": [1e4, 3e5], "initialTemp": 10.0})
L = Library("default"); L.init([0])
print("launch use", use, "budget", budget, "ntime", nt)
t0 = time.time()
out = L.run_grid(5, p, timepoints=nt, want_physics=True, want_chem=True, pp_grid=g, pp_coldens=use, step_budget=budget)
S = {k: out["stats"][:, i] for i, k in enumerate(STAT_FIELDS)}
print(f"done in {time.time() - t0:.2f} s flags {out['flag']} nst {S['nst']} intervals {S['nintervals']} failcalls {S['nfailcall']}")
print("times", out["physics"][0, : nt + 1, 0])

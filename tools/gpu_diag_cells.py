"""GPU diagnostics (run on the box): (1) the 12-cell slice of config 2 used by
test_small_grid_against_oracle with trajectories, saved for offline comparison with the oracle;
(2) a strided sample of config 2 run in chunks of one cell per SM with a step budget, stats saved."""
import sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import get_library, STAT_FIELDS
from uclchem_b200.params import params_from_dict, PARAM_INDEX
lib = get_library(); lib.init()
out_dir = ROOT / "gpurun_out"; out_dir.mkdir(exist_ok=True)
what = sys.argv[1]
if what == "small":
    dens, temp, zeta = np.meshgrid([1e3, 1e5, 1e7], [10.0, 55.0], [1.0, 100.0], indexing="ij")
    p = params_from_dict({"initialDens": dens.ravel(), "initialTemp": temp.ravel(), "zeta": zeta.ravel(),
                          "radfield": 1.0, "baseAv": 2.0, "rout": 0.05, "finalTime": 1e4})
    o = lib.run_grid(0, p, timepoints=60, want_chem=True, want_physics=True)
    np.savez(out_dir / "diag_small.npz", params=p, abund=o["abund"], physics=o["physics"], stats=o["stats"],
             flag=o["flag"], y_final=o["y_final"])
    print("small done", o["flag"], o["stats"][:, 0])
else:
    n = int(sys.argv[2]); budget = int(sys.argv[3]); chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 148
    P = config2_params()
    rng = np.random.default_rng(20261017)
    idx = np.sort(rng.permutation(P.shape[1])[:n])
    res = []
    for k in range(0, n, chunk):
        ii = idx[k:k + chunk]
        p = np.ascontiguousarray(P[:, ii])
        t = time.time(); o = lib.run_grid(0, p, step_budget=budget); dt = time.time() - t
        st = o["stats"]; S = {kk: st[:, i] for i, kk in enumerate(STAT_FIELDS)}
        cyc = S["cyc_total"] / 1.9e9
        print(f"chunk {k}: wall {dt:.1f}s flags {dict(zip(*np.unique(o['flag'], return_counts=True)))} "
              f"nst med {np.median(S['nst']):.0f} max {S['nst'].max()} sec med {np.median(cyc):.2f} max {cyc.max():.2f}")
        res.append((ii, p, st, o["flag"], o["y_final"]))
        np.savez(out_dir / f"diag_sample_{n}.npz", idx=np.concatenate([r[0] for r in res]),
                 params=np.concatenate([r[1] for r in res], axis=1), stats=np.concatenate([r[2] for r in res]),
                 flag=np.concatenate([r[3] for r in res]), y_final=np.concatenate([r[4] for r in res]))
    st = np.concatenate([r[2] for r in res]); pp = np.concatenate([r[1] for r in res], axis=1)
    S = {kk: st[:, i] for i, kk in enumerate(STAT_FIELDS)}
    order = np.argsort(S["nst"])[::-1]
    for k in order[:25]:
        print(f"dens {pp[PARAM_INDEX['initialdens'],k]:.2e} T {pp[PARAM_INDEX['initialtemp'],k]:.0f} zeta {pp[PARAM_INDEX['zeta'],k]:.1f} "
              f"nst {S['nst'][k]} nlu {S['nlu'][k]} nje {S['nje'][k]} ncfn {S['ncfn'][k]} netf {S['netf'][k]} failcalls {S['nfailcall'][k]} nint {S['nintervals'][k]}")

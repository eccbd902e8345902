"""Heavy cells of config 2 (tools/config2_heavy_cells.npy) under a step budget, for two builds.
usage: gpu_heavy_ab.py <budget> <tag> [<tag> ...]"""
import sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.params import PARAM_INDEX
P = config2_params()
idx = np.load(ROOT / "tools/config2_heavy_cells.npy")
budget = int(float(sys.argv[1]))
p = np.ascontiguousarray(P[:, idx])
for tag in sys.argv[2:]:
    lib = Library(tag); lib.init([0])
    t = time.time(); o = lib.run_grid(0, p, step_budget=budget); dt = time.time() - t
    st = o["stats"]; S = {k: st[:, i] for i, k in enumerate(STAT_FIELDS)}
    nst = S["nst"]
    print(f"[{tag}] budget {budget}: wall {dt:.1f} s flags {dict(zip(*np.unique(o['flag'], return_counts=True)))}")
    print("   nst percentiles 10/50/90/99/max:", np.percentile(nst, [10, 50, 90, 99, 100]).astype(int), "sum", nst.sum(),
          " netf/nst %.3f ncfn/nst %.3f nlu/nst %.3f nje/nst %.3f failcalls mean %.1f" % (
              S["netf"].sum() / nst.sum(), S["ncfn"].sum() / nst.sum(), S["nlu"].sum() / nst.sum(), S["nje"].sum() / nst.sum(), S["nfailcall"].mean()))
    k = np.argsort(nst)[::-1][:6]
    for j in k:
        print(f"     cell {idx[j]} dens {p[PARAM_INDEX['initialdens'], j]:.2e} T {p[PARAM_INDEX['initialtemp'], j]:.0f} zeta {p[PARAM_INDEX['zeta'], j]:.1f} "
              f"nst {nst[j]} netf {S['netf'][j]} ncfn {S['ncfn'][j]} failcalls {S['nfailcall'][j]} nint {S['nintervals'][j]} flag {o['flag'][j]} sec {S['cyc_total'][j] / 1.9e9:.1f}")
    np.savez(ROOT / f"gpurun_out/heavy_{tag}_{budget}.npz", idx=idx, stats=st, flag=o["flag"])
    lib.shutdown()

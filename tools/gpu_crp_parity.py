"""Second network on the GPU vs the oracle, cell by cell, at two tolerances (run on the GPU box)."""
import sys, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from oracle.oracle import Oracle
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.network import Network
from uclchem_b200.params import params_from_dict
tag = sys.argv[1] if len(sys.argv) > 1 else "crp_photo"
net = Network.from_json(ROOT / "uclchem_b200" / "networks" / f"{tag}.json")
L = Library(tag); L.init([0]); orc = Oracle(net)
S = {k: i for i, k in enumerate(STAT_FIELDS)}
for rt, am in ((1e-5, 1e-15), (1e-8, 1e-25)):
    p = params_from_dict({"initialDens": [1e4, 1e5, 1e6, 1e7], "initialTemp": [10.0, 20.0, 60.0, 100.0], "zeta": [1.0, 1.0, 30.0, 1e3],
                          "finalTime": 1e3, "reltol": rt, "abstol_min": am})
    out = L.run_grid(0, p, step_budget=300000)
    ref, _, flag, st = orc.run_grid(0, p, nthreads=4)
    for c in range(4):
        a, b = out["y_final"][c, :net.nspec], ref[c, :net.nspec]
        m = b > 1e-15
        d = np.abs(np.log10(a[m] / b[m])); k = np.argmax(d); i = np.where(m)[0][k]
        s = out["stats"][c]
        print(f"reltol {rt:g} cell {c}: flags {out['flag'][c]}/{flag[c]} nst gpu {s[S['nst']]} oracle {st[c, 0]}  netf {s[S['netf']]}/{st[c, 6]} "
              f"ncfn {s[S['ncfn']]}/{st[c, 5]} nsing {s[S['nsing']]} maxcor {s[S['nmaxcor']]} diverge {s[S['ndiverge']]} failcalls {s[S['nfailcall']]}  "
              f"max dex {d.max():.4f} at {net.names[i]} gpu {a[i]:.3e} oracle {b[i]:.3e}; species above 0.01 dex: {(d > 0.01).sum()}")

"""Run on the GPU box: the hot, dense, strongly irradiated cell (n = 1e7, T = 100 K, zeta = 1e3, 1e3 yr) at loose
tolerances around the reference test's reltol = 1e-5 / abstol_min = 1e-15, engine vs oracle, both networks, plus the
converged answer (oracle at reltol 1e-8).  Shows which arm has an 'accident' (step-count blow-up) at which tolerance."""
import sys, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from oracle.oracle import Oracle
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.network import Network
from uclchem_b200.params import params_from_dict
S = {k: i for i, k in enumerate(STAT_FIELDS)}
RT = [0.5e-5, 0.9e-5, 0.97e-5, 1.0e-5, 1.03e-5, 1.1e-5, 2e-5, 1e-6]
for tag in ("default", "crp_photo"):
    net = Network.from_json(ROOT / "uclchem_b200" / "networks" / f"{tag}.json")
    L = Library(tag); L.init([0]); orc = Oracle(net)
    base = {"initialDens": 1e7, "initialTemp": 100.0, "zeta": 1e3, "finalTime": 1e3}
    truth = orc.run_model(0, params_from_dict(dict(base, reltol=1e-8, abstol_min=1e-25))[:, 0])["y_final"][:net.nspec]
    p = params_from_dict(dict(base, reltol=RT, abstol_min=1e-15))
    out = L.run_grid(0, p, step_budget=100000)
    ref, _, flag, st = orc.run_grid(0, p, nthreads=8)
    def dex(a, b):
        m = b > 1e-15
        return np.abs(np.log10(np.maximum(a[m], 1e-300) / b[m])).max()
    for c, rt in enumerate(RT):
        s = out["stats"][c]
        print(f"{tag} reltol {rt:g}: flags {out['flag'][c]}/{flag[c]} nst gpu {s[S['nst']]} oracle {st[c, 0]} failcalls {s[S['nfailcall']]}  "
              f"dex gpu-oracle {dex(out['y_final'][c, :net.nspec], ref[c, :net.nspec]):.4f}  gpu-converged {dex(out['y_final'][c, :net.nspec], truth):.4f}  "
              f"oracle-converged {dex(ref[c, :net.nspec], truth):.4f}")

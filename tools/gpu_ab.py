"""A/B two builds of the library on the same cells: bitwise comparison of results, kernel time, phase shares.
usage: gpu_ab.py <tagA> <tagB> <ncells> [finalTime]"""
import sys, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.params import PARAM_INDEX
P = config2_params()
n = int(sys.argv[3]); final = float(sys.argv[4]) if len(sys.argv) > 4 else 1e6
heavy = set(np.load(ROOT / "tools/config2_heavy_cells.npy").tolist())
idx = [i for i in np.linspace(0, P.shape[1] - 1, n).astype(int) if i not in heavy]
p = np.ascontiguousarray(P[:, idx]); p[PARAM_INDEX["finaltime"]] = final
res = {}
for tag in sys.argv[1:3]:
    lib = Library(tag); lib.init([0])
    lib.run_grid(0, p[:, :148])  # warm-up
    o = lib.run_grid(0, p); ms, _ = lib.last_kernel_ms(0)
    st = o["stats"]; S = {k: st[:, i].astype(float) for i, k in enumerate(STAT_FIELDS)}
    tot = S["cyc_total"].sum()
    print(f"[{tag}] cells {len(idx)} kernel {ms/1e3:.3f} s  models/s {len(idx)/(ms/1e3):.1f}  flags!=0 {(o['flag']!=0).sum()}  "
          f"nst {S['nst'].mean():.0f}  cyc/step {tot/S['nst'].sum():.0f}")
    print("   shares", {k[4:]: round(S[k].sum() / tot, 3) for k in ("cyc_rates", "cyc_rhs", "cyc_jac", "cyc_factor", "cyc_dense", "cyc_solve")})
    print("   cyc/call rhs %.0f solve %.0f factor %.0f dense %.0f jac %.0f" % (S["cyc_rhs"].sum()/S["nfe"].sum(), S["cyc_solve"].sum()/S["nni"].sum(),
          S["cyc_factor"].sum()/S["nlu"].sum(), S["cyc_dense"].sum()/S["nlu"].sum(), S["cyc_jac"].sum()/S["nlu"].sum()))
    res[tag] = o
    lib.shutdown()
a, b = (res[t] for t in sys.argv[1:3])
same = np.array_equal(a["y_final"], b["y_final"])
m = a["y_final"][:, :335] > 1e-15
dex = np.abs(np.log10(b["y_final"][:, :335][m] / a["y_final"][:, :335][m])).max()
print("bitwise identical:", same, " max dex:", dex, " nst equal:", np.array_equal(a["stats"][:, 0], b["stats"][:, 0]))

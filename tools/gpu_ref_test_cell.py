"""Run on the GPU box: the cell of the reference's own test of the crp-photo network (tests/test_photo_on_grain.py:104-114:
n = 1e4, T = 10 K, 5 Myr) at that test's tolerances (reltol 1e-5, abstol_min 1e-15) and at tighter ones, engine and
oracle, each against the engine's converged answer (reltol 1e-8 / 1e-9 agree => converged)."""
import sys, functools, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from oracle.oracle import Oracle
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.network import Network
from uclchem_b200.params import params_from_dict
S = {k: i for i, k in enumerate(STAT_FIELDS)}
net = Network.from_json(ROOT / "uclchem_b200" / "networks" / "crp_photo.json")
L = Library("crp_photo"); L.init([0]); orc = Oracle(net)
RT = [1e-5, 0.97e-5, 1.03e-5, 3e-6, 1e-6, 1e-7, 1e-8, 1e-9]
AM = [1e-15, 1e-15, 1e-15, 1e-18, 1e-20, 1e-22, 1e-25, 1e-28]
base = {"endAtFinalDensity": False, "freefall": False, "initialDens": 1e4, "initialTemp": 10.0, "finalDens": 1e5, "finalTime": 5.0e6}
p = params_from_dict(dict(base, reltol=RT, abstol_min=AM))
t = time.time(); out = L.run_grid(0, p, step_budget=2000000); print(f"gpu {time.time() - t:.1f} s")
y = out["y_final"][:, :net.nspec]
def dex(a, b, floor):
    m = b > floor; d = np.abs(np.log10(np.maximum(a[m], 1e-300) / b[m])); i = np.where(m)[0][np.argmax(d)]
    return f"{d.max():.4f} ({net.names[i]}, {(d > 0.01).sum()}/{m.sum()} > 0.01)"
truth = y[6]
for c, rt in enumerate(RT):
    s = out["stats"][c]
    print(f"engine reltol {rt:g} abstol_min {AM[c]:g}: flag {out['flag'][c]} nst {s[S['nst']]} netf {s[S['netf']]} ncfn {s[S['ncfn']]} failcalls {s[S['nfailcall']]} | vs engine 1e-8: "
          f">1e-12 {dex(y[c], truth, 1e-12)}  >1e-8 {dex(y[c], truth, 1e-8)}  >1e-6 {dex(y[c], truth, 1e-6)}")
orc.set_deadline(150)
ref, _, flag, st = orc.run_grid(0, p[:, :3], nthreads=3)
for c in range(3):
    print(f"oracle reltol {RT[c]:g}: flag {flag[c]} nst {st[c, 0]} netf {st[c, 6]} | vs engine 1e-8: >1e-12 {dex(ref[c, :net.nspec], truth, 1e-12)}  >1e-8 {dex(ref[c, :net.nspec], truth, 1e-8)}  "
          f">1e-6 {dex(ref[c, :net.nspec], truth, 1e-6)} | vs engine same tol: >1e-8 {dex(ref[c, :net.nspec], y[c], 1e-8)}")
names = ["OH", "OCS", "CO", "CS", "CH3OH"]
print("out_species", names, "engine 1e-8", [f"{truth[net.names.index(n)]:.4e}" for n in names], "engine 1e-5", [f"{y[0][net.names.index(n)]:.4e}" for n in names],
      "oracle 1e-5", [f"{ref[0][net.names.index(n)]:.4e}" for n in names])
np.save(ROOT / "gpurun_out" / "ref_test_cell_y.npy", np.vstack([y, ref[:, :net.nspec]]))

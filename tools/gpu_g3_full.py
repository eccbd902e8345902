"""G3 at full length on the GPU engine (run on the box): hot_core(3, 300) from startcollapse, 1 Myr, every stored
time of the reference's phase2-full.dat (tests/golden/phase2_full.npz)."""
import sys, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.params import params_from_dict
gold = np.load(ROOT / "tests/golden/phase2_full.npz"); sc = np.load(ROOT / "tests/golden/startcollapse.npy")
lib = Library("default"); lib.init([0])
rts = [1e-8, 0.97e-8, 1.03e-8]
p = params_from_dict({"endAtFinalDensity": False, "freefall": False, "initialDens": 1e5, "initialTemp": 10.0, "finalDens": 1e5,
                      "finalTime": 1.0e6, "freezeFactor": 0.0, "thermdesorb": True, "temp_indx": 3, "max_temperature": 300.0,
                      "reltol": rts})
y0 = np.repeat(np.append(sc, 1e5)[None, :], len(rts), axis=0)
o = lib.run_grid(1, p, y0=y0, timepoints=500, want_physics=True, want_chem=True)
for c, rt in enumerate(rts):
    S = dict(zip(STAT_FIELDS, o["stats"][c]))
    n = int(S["nintervals"]) + 1
    dex = []
    for row in range(1, min(n, 283)):
        a, b = o["abund"][c, row], gold["abund"][row]; m = b > 1e-15
        dex.append(np.abs(np.log10(a[m] / b[m])).max())
    dex = np.array(dex)
    print(f"reltol {rt:g}: flag {o['flag'][c]} rows {n} nst {S['nst']} failcalls {S['nfailcall']} max dex {dex.max():.4f} at row {dex.argmax() + 1}; "
          f"rows >= 0.01: {(np.where(dex >= 0.01)[0] + 1).tolist()}; second worst {np.sort(dex)[-2]:.4f}")
np.savez_compressed(ROOT / "gpurun_out/g3_full_gpu.npz", abund=o["abund"], physics=o["physics"], stats=o["stats"], flag=o["flag"])

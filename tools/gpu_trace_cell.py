"""Debug: Newton-iteration trace of one config-2 style cell (run on the GPU box).
usage: gpu_trace_cell.py dens temp zeta final_time [records]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
cap = int(sys.argv[5]) if len(sys.argv) > 5 else 4000
os.environ["UCLGPU_TRACE"] = str(cap); os.environ["UCLGPU_TRACE_FILE"] = str(ROOT / "gpurun_out/trace_cell.bin")
from uclchem_b200.params import params_from_dict
from uclchem_b200._capi import get_library
lib = get_library(); lib.init()
p = params_from_dict({"initialDens": float(sys.argv[1]), "initialTemp": float(sys.argv[2]), "zeta": float(sys.argv[3]),
                      "radfield": 1.0, "baseAv": 2.0, "rout": 0.05, "finalTime": float(sys.argv[4])})
out = lib.run_grid(0, p, timepoints=60, want_chem=True, want_physics=True)
print(out["flag"], out["stats"][0][:12])
tr = np.fromfile(ROOT / "gpurun_out/trace_cell.bin").reshape(-1, 12)
tr = tr[tr[:, 1] != 0]
np.set_printoptions(linewidth=250)
print("tn h nq m del dcon rc nst.jcur ySURF yBULK S yh0SURF")
for r in tr[:cap]:
    print(" ".join(f"{v:.6e}" for v in r))

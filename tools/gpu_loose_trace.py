"""Debug (run on the GPU box): Newton-iteration traces of one cell at loose tolerances, engine and oracle side by side.
usage: gpu_loose_trace.py tag dens temp zeta final_time reltol abstol_min [records]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
tag = sys.argv[1]
dens, temp, zeta, tfin, rt, am = (float(a) for a in sys.argv[2:8])
cap = int(sys.argv[8]) if len(sys.argv) > 8 else 40000
out_dir = ROOT / "gpurun_out"; out_dir.mkdir(exist_ok=True)
os.environ["UCLGPU_TRACE"] = str(cap); os.environ["UCLGPU_TRACE_FILE"] = str(out_dir / f"trace_{tag}_gpu.bin")
os.environ["ORC_TRACE"] = str(out_dir / f"trace_{tag}_oracle.txt")
from uclchem_b200.params import params_from_dict
from uclchem_b200._capi import Library, STAT_FIELDS
from uclchem_b200.network import Network
from oracle.oracle import Oracle
net = Network.from_json(ROOT / "uclchem_b200" / "networks" / f"{tag}.json")
lib = Library(tag); lib.init([0])
p = params_from_dict({"initialDens": dens, "initialTemp": temp, "zeta": zeta, "finalTime": tfin, "reltol": rt, "abstol_min": am})
out = lib.run_grid(0, p, step_budget=300000)
print("gpu flag", out["flag"], dict(zip(STAT_FIELDS, out["stats"][0])))
r = Oracle(net).run_model(0, p[:, 0])
print("oracle flag", r["flag"], r["stats"])
a, b = out["y_final"][0, :net.nspec], r["y_final"][:net.nspec]
m = b > 1e-15
d = np.abs(np.log10(a[m] / b[m]))
print("max dex", d.max(), "species above 0.01 dex:", int((d > 0.01).sum()))
np.save(out_dir / f"trace_{tag}_yfinal.npy", np.stack([out["y_final"][0], r["y_final"]]))

import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
if len(sys.argv) > 2:
    os.environ["UCLGPU_DUMP_AT"] = sys.argv[2]; os.environ["UCLGPU_DUMP_FILE"] = str(ROOT / f"gpurun_out/dump_{sys.argv[2]}.bin")
os.environ["UCLGPU_TRACE"] = "4000"; os.environ["UCLGPU_TRACE_FILE"] = str(ROOT / "gpurun_out/trace.bin")
from uclchem_b200.params import params_from_dict
from uclchem_b200._capi import get_library
lib = get_library(); lib.init()
p1 = params_from_dict({"initialDens": 1e4, "initialTemp": 10.0, "finalTime": float(sys.argv[1])})
out = lib.run_grid(0, p1)
print(out['flag'], out['stats'][0][:12])

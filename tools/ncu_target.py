"""Short hot-path run for ncu: N cells strided out of the config-2 grid, integrated to `finalTime` years.
usage: ncu_target.py <ncells> <finalTime>   (same kernel and launch shape as bench.py, shorter models)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
from bench import config2_params
from uclchem_b200._capi import get_library
from uclchem_b200.params import PARAM_INDEX
lib = get_library(); lib.init([0])
P = config2_params()
n = int(sys.argv[1]); final = float(sys.argv[2])
heavy = set(np.load(ROOT / "tools/config2_heavy_cells.npy").tolist())
idx = [i for i in np.linspace(0, P.shape[1] - 1, n).astype(int) if i not in heavy]
p = np.ascontiguousarray(P[:, idx]); p[PARAM_INDEX["finaltime"]] = final
for rep in range(2):
    o = lib.run_grid(0, p)
    print("cells", len(idx), "kernel ms", lib.last_kernel_ms(0), "flags ok", bool((o["flag"] == 0).all()), flush=True)

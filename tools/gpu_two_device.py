"""In-process multi-device sharding of uclgpu_run_grid (run on a box with >= 2 GPUs): the same 592 config-2 cells
on one device and dealt round-robin over two; results must be bitwise identical, and the two devices' kernel
times show the balance."""
import sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import Library
P = config2_params()
p = np.ascontiguousarray(P[:, np.linspace(0, 9999, 592).astype(int)])
lib = Library("default")
lib.init([0]); lib.run_grid(0, p[:, :148], step_budget=2000)
t = time.time(); a = lib.run_grid(0, p, step_budget=100000); t1 = time.time() - t
ms1 = lib.last_kernel_ms(0)[0]
lib.init([0, 1]); lib.run_grid(0, p[:, :296], step_budget=2000)
t = time.time(); b = lib.run_grid(0, p, step_budget=100000); t2 = time.time() - t
ms2 = [lib.last_kernel_ms(d)[0] for d in (0, 1)]
print(f"1 device: {t1:.2f} s (kernel {ms1 / 1e3:.2f} s); 2 devices: {t2:.2f} s (kernels {ms2[0] / 1e3:.2f} / {ms2[1] / 1e3:.2f} s)")
print("bitwise identical:", np.array_equal(a["y_final"], b["y_final"]) and np.array_equal(a["flag"], b["flag"]),
      " flags != 0:", int((a["flag"] != 0).sum()))
# chunked launches: cap the per-launch result storage so the same grid needs several rounds
c = lib.run_grid(0, p, step_budget=100000, chunk_bytes=100 * 3500)
print("chunked (100 cells per launch) identical:", np.array_equal(a["y_final"], c["y_final"]), " launches per device:",
      [lib.last_kernel_ms(d)[1] for d in (0, 1)])

"""All 10^4 config-2 cells under a step budget: per-cell solver counters and cycles (load-balance / cost-model data).
usage: gpu_grid_full.py <budget> [tag]"""
import sys, time, functools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
print = functools.partial(print, flush=True)
from bench import config2_params
from uclchem_b200._capi import Library, STAT_FIELDS
budget = int(float(sys.argv[1])); tag = sys.argv[2] if len(sys.argv) > 2 else "default"
lib = Library(tag); lib.init([0])
P = config2_params()
lib.run_grid(0, P[:, ::68][:, :148], step_budget=2000)
t = time.time(); o = lib.run_grid(0, P, step_budget=budget); dt = time.time() - t
ms, _ = lib.last_kernel_ms(0)
st = o["stats"]; S = {k: st[:, i] for i, k in enumerate(STAT_FIELDS)}
print(f"[{tag}] budget {budget}: wall {dt:.1f} s kernel {ms/1e3:.1f} s flags {dict(zip(*np.unique(o['flag'], return_counts=True)))}")
sec = S["cyc_total"] / 1.965e9
print("   per-cell s: sum/148 %.1f max %.1f; nst pct 1/50/99/max" % (sec.sum() / 148, sec.max()), np.percentile(S["nst"], [1, 50, 99, 100]).astype(int))
np.savez_compressed(ROOT / f"gpurun_out/grid_full_{tag}_{budget}.npz", stats=st, flag=o["flag"], y_final=o["y_final"], kernel_ms=ms)
